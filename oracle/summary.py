"""TEST INFRASTRUCTURE (oracle): restatement of the reference's layer_summary (trainers/model_utils.py:4-6) -
tf.nn.zero_fraction + tf.summary.histogram - on numpy arrays.  tf.summary.histogram fills a HistogramProto through
core/lib/histogram/histogram.cc: default bucket limits +-1e-12 * 1.1^i (while < 1e20), 0 and +-DBL_MAX, a value goes
to bucket upper_bound(limits, value); min, max, num, sum, sum_squares are kept in float64.  Only tests may import this."""
import sys

import numpy as np


def default_bucket_limits():
    pos = []
    v = 1.0e-12
    while v < 1.0e20:
        pos.append(v)
        v *= 1.1
    pos.append(sys.float_info.max)
    return np.array([-x for x in reversed(pos)] + [0.0] + pos, dtype=np.float64)


def layer_summary(value):
    v = np.asarray(value, dtype=np.float32).reshape(-1)
    lim = default_bucket_limits()
    idx = np.minimum(np.searchsorted(lim, v.astype(np.float64), side="right"), lim.size - 1)     # upper_bound
    counts = np.bincount(idx, minlength=lim.size).astype(np.int64)
    nz = np.nonzero(counts)[0]
    v64 = v.astype(np.float64)
    return {"fraction_of_zero_values": float((v == 0).mean()) if v.size else 0.0,
            "activation": {"min": float(v64.min()), "max": float(v64.max()), "num": float(v.size), "sum": float(v64.sum()),
                           "sum_squares": float((v64 * v64).sum()), "bucket_limit": lim[nz].tolist(),
                           "bucket": counts[nz].astype(np.float64).tolist()}}
