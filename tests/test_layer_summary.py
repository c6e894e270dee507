"""layer_summary side outputs (reference trainers/model_utils.py:4-6): the oracle's HistogramProto restatement on CPU,
and on the GPU the summarised tensors against the oracle plus the device reduction against the oracle's on the SAME values."""
import numpy as np
import pytest

from oracle import summary as osum


def test_default_bucket_limits_follow_histogram_cc():
    lim = osum.default_bucket_limits()
    assert lim.size == 2 * 775 + 1 and lim[775] == 0.0
    assert lim[776] == 1.0e-12 and np.isclose(lim[777], 1.1e-12, rtol=1e-15)
    assert (np.diff(lim) > 0).all() and lim[-1] == np.finfo(np.float64).max and lim[0] == -lim[-1]


def test_oracle_layer_summary_known_values():
    s = osum.layer_summary(np.array([0.0, 0.0, 1.0, -2.0, 0.5], np.float32))
    assert s["fraction_of_zero_values"] == 0.4
    h = s["activation"]
    assert (h["min"], h["max"], h["num"], h["sum"], h["sum_squares"]) == (-2.0, 1.0, 5.0, -0.5, 5.25)
    assert sum(h["bucket"]) == 5.0
    # zeros land in the bucket whose limit is the first one above 0 (1e-12); every value v obeys prev_limit <= v < limit
    lim = osum.default_bucket_limits()
    for limit, cnt in zip(h["bucket_limit"], h["bucket"]):
        i = int(np.searchsorted(lim, limit))
        inside = [v for v in (0.0, 0.0, 1.0, -2.0, 0.5) if lim[i - 1] <= v < lim[i]]
        assert len(inside) == cnt
    assert 1.0e-12 in h["bucket_limit"]


@pytest.mark.gpu
@pytest.mark.parametrize("shape", ["ml100k_k4", "criteo_k16", "tc_tower_dropout"])
def test_gpu_layer_summary_matches_oracle(shape):
    from recommender_tensorflow_b200 import synth
    from recommender_tensorflow_b200.engine import DeepFMEngine
    from oracle import transforms
    from tests.util import make_pair, ml100k_columns
    rng = np.random.default_rng(70)
    if shape == "criteo_k16":
        cats, nums = synth.criteo_columns(2000, n_cat=26, n_num=13)
        eng = DeepFMEngine(cats, nums, embedding_size=16, hidden_units=(16, 16), max_batch=512)
        batches = [synth.criteo_batch(500, rng) for _ in range(3)]
    else:
        cols, dtypes = ml100k_columns()
        kw = dict(embedding_size=16, hidden_units=(64, 32), dropout=0.25, dropout_seed=9) if shape == "tc_tower_dropout" else \
            dict(embedding_size=4, hidden_units=(16, 16))
        eng = DeepFMEngine(cols, (), max_batch=512, feature_dtypes=dtypes, **kw)
        ml = synth.ML100K()
        batches = [ml.batch(300, rng) for _ in range(3)]
    ora, _ = make_pair(eng, seed=71)
    for feats, y in batches[:2]:                         # two train steps first: replayed rows, non-trivial step counter
        eng.train_step(feats, y)
        ora.train_step_raw(feats, y)
    feats, y = batches[2]
    got = eng.layer_summary(feats, train=True)
    o32 = ora.o32
    ids = transforms.transform(eng.specs, feats)
    x = np.stack([np.asarray(feats[n], dtype=np.float32) for n in o32.num], 1) if o32.dn else None
    want_t = o32.layer_tensors(ids, x, train=True)
    B = len(y)
    hid = sum(eng.hidden)
    dev = {"linear/linear": eng.summary_tensor(0, B), "mf/logits": eng.summary_tensor(1, B), "dnn/dnn/logits": eng.summary_tensor(3, B),
           "deep_fm/logits": eng.summary_tensor(4, B)}
    hidden = eng.summary_tensor(2, B * hid).reshape(B, hid)
    col = 0
    for i, hsz in enumerate(eng.hidden):
        dev["dnn/dnn/hiddenlayer_%d" % i] = hidden[:, col:col + hsz]
        col += hsz
    assert list(got) == eng.summary_names() and set(got) == set(want_t)
    for name in got:
        w = np.asarray(want_t[name], np.float64)
        d = np.asarray(dev[name], np.float64).reshape(w.shape)
        tol = 2e-5 * np.abs(w) + 2e-5 * max(1.0, float(np.abs(w).max()))
        assert (np.abs(d - w) <= tol).all(), (name, float(np.abs(d - w).max()))
        assert ((d == 0) == (w == 0)).mean() > 0.999, name         # the same units are dead / dropped
        # the device reduction, against the oracle's reduction of the very same values: exact
        ref = osum.layer_summary(dev[name])
        assert got[name]["fraction_of_zero_values"] == ref["fraction_of_zero_values"], name
        ha, hb = got[name]["activation"], ref["activation"]
        assert (ha["min"], ha["max"], ha["num"]) == (hb["min"], hb["max"], hb["num"]), name
        assert ha["bucket_limit"] == hb["bucket_limit"] and ha["bucket"] == hb["bucket"], name
        assert np.isclose(ha["sum"], hb["sum"], rtol=1e-9, atol=1e-9) and np.isclose(ha["sum_squares"], hb["sum_squares"], rtol=1e-9), name
