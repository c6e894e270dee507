"""Shared helpers of the parity tests: build the CUDA engine and the oracle from one description."""
import numpy as np
import torch

from oracle.deepfm import OracleDeepFM, default_opt, init_weights
from oracle.transforms import num_buckets


def ml100k_columns():
    from recommender_tensorflow_b200.trainers import ml_100k
    return ml_100k.get_feature_columns()["linear"], ml_100k.FEATURE_DTYPES


def oracle_cfg(engine, **over):
    """Oracle config equivalent to a DeepFMEngine (same model-order columns, same hyper-parameters)."""
    def o(d):
        r = default_opt(d["name"], d["lr"])
        for k in ("beta1", "beta2", "eps", "init_acc"):
            if k in d and (k in r or d["name"] == "RMSProp"):
                r[k] = d[k]
        return r
    cfg = dict(cat=[dict(s) for s in engine.specs], num=[c.key for c in engine.num_columns], k=engine.k,
               hidden=list(engine.hidden), use_linear=engine.use_linear, use_mf=engine.use_mf,
               use_dnn=engine.use_dnn, loss_reduction=engine.loss_reduction, opt_deep=o(engine.opt_deep),
               opt_linear=o(engine.opt_linear), dropout=getattr(engine, "dropout", 0.0),
               dropout_seed=getattr(engine, "dropout_seed", 0), activation=getattr(engine, "activation", "relu"))
    cfg.update(over)
    return cfg


def make_pair(engine, seed=0, lin_scale=0.05):
    """Same injected initial weights on both sides.  Linear weights get small random values (instead
    of TF's zeros) so that the linear path is exercised from step 1."""
    cfg = oracle_cfg(engine)
    w = init_weights(cfg, seed)
    rng = np.random.default_rng(seed + 1000)
    for name in ("lin", "num_lin", "bias"):
        if name in w:
            w[name] = (rng.standard_normal(w[name].shape) * lin_scale).astype(np.float32)
    for name in list(w):
        if name.startswith("b") and name != "bias":
            w[name] = (rng.standard_normal(w[name].shape) * 0.05).astype(np.float32)
    engine.set_weights(w)
    return OraclePair(OracleDeepFM(cfg, w), OracleDeepFM(cfg, w, dtype=torch.float64)), w


def rel_err(a, b, floor=1e-6):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    return float(np.max(np.abs(a - b) / (np.abs(b) + floor))) if a.size else 0.0


class OraclePair:
    """The float32 oracle (the reference semantics) and its float64 twin (tolerance calibration)."""

    def __init__(self, o32, o64):
        self.o32, self.o64 = o32, o64

    def train_step_raw(self, feats, y):
        self.last64 = self.o64.train_step_raw(feats, y)
        return self.o32.train_step_raw(feats, y)

    def forward(self, ids, x=None):
        return self.o32.forward(ids, x)

    def state(self):
        return self.o32.state()

    def state64(self):
        return self.o64.state()


def assert_state_close(engine_state, oracle_state, rtol=1e-5, atol=1e-7, what="", oracle_state64=None, report=None, tc_noise=None):
    """|got - ref| <= atol + rtol*|ref|  (the 1e-5 bar of BASELINE.json north_star).  When the float64 twin is given
    the reference is the float64 value and the allowance of every ROW (leading index: a table row, a weight-matrix
    row) grows by 4x the float32 oracle's own worst rounding error ON THAT ROW: summation-order noise of a hot row
    (thousands of lookups summed in an order TF leaves unspecified) is not a parity failure, but it buys nothing
    for the other rows.  `report` (dict) receives the worst achieved relative error per tensor."""
    for name, ref in oracle_state.items():
        got = engine_state[name].reshape(ref.shape).astype(np.float64)
        if oracle_state64 is not None:
            ref64 = oracle_state64[name].astype(np.float64)
            e32 = np.abs(ref.astype(np.float64) - ref64)
            if e32.ndim > 1:
                e32 = e32.reshape(e32.shape[0], -1).max(axis=1).reshape((-1,) + (1,) * (ref.ndim - 1))
            err = np.abs(got - ref64)
            tol = atol + rtol * np.abs(ref64) + 4.0 * e32
            base = ref64
        else:
            err = np.abs(got - ref)
            tol = atol + rtol * np.abs(ref)
            base = ref.astype(np.float64)
        if report is not None and ref.size:
            report[name] = float((err / (np.abs(base) + atol / rtol)).max())
        bad = err > tol
        # Adam divides by sqrt(v) + eps: where a gradient element is tiny, rounding noise in it is amplified by up to
        # lr / eps-hat ~ 3e3, for the float32 oracle exactly as for the CUDA path, and two independent draws of that
        # noise differ by more than 4x now and then.  A tail of at most 2e-5 of a tensor's elements may therefore
        # exceed the allowance, by no more than another factor 4; everything else must be inside it.
        n_tail = int(ref.size * 2e-5) if oracle_state64 is not None else 0
        worst = float((err / tol).max()) if ref.size else 0.0
        if tc_noise is not None and oracle_state64 is not None:
            # tcgen05 tower (3xTF32 with truncating TMEM accumulation: ~5x the rounding noise of an fp32 sum, measured in
            # tests/test_gpu_tc_gemm.py and tools/debug_cfg2.py).  Its noise is not correlated with the float32 oracle's
            # row by row, so the per-row calibration does not apply: every element must be inside
            # rtol + tc_noise x the oracle's worst deviation ON THE TENSOR (the fraction outside the strict per-row
            # allowance is reported).
            e32t = float(np.abs(ref.astype(np.float64) - oracle_state64[name].astype(np.float64)).max()) if ref.size else 0.0
            loose = atol + rtol * np.abs(oracle_state64[name].astype(np.float64)) + tc_noise * e32t
            if report is not None:
                report[name + " outside-strict-%"] = 100.0 * float(bad.mean())
            assert (err <= loose).all(), (
                "%s %s [tc]: %.3f %% elements outside the strict allowance, worst abs %.3e, oracle32 worst %.3e" % (
                    what, name, 100.0 * float(bad.mean()), float(err.max()), e32t))
            continue
        assert int(bad.sum()) <= n_tail and worst <= (4.0 if n_tail else 1.0), (
            "%s %s: %d / %d elements off (tail allowance %d), worst abs %.3e (ref scale %.3e), worst err/tol %.2f" % (
                what, name, int(bad.sum()), ref.size, n_tail, float(err.max()), float(np.abs(ref).max()), worst))


def assert_step_close(loss, logits, pair, rloss, rlogits, rtol=1e-5, what="", noise=4.0):
    """loss / logits of one step against the oracle pair (float64 twin calibrates cancellation noise).
    noise: multiple of the float32 oracle's own worst deviation from float64 that is allowed on top of rtol (4 for
    float32 CUDA-core arithmetic; the tcgen05 3xTF32 tower passes 8 - see tests/test_gpu_tc_gemm.py on the
    truncating TMEM accumulation)."""
    loss64, logits64 = pair.last64
    e32 = float(np.abs(rlogits.astype(np.float64) - logits64).max())
    err = np.abs(logits.astype(np.float64) - logits64)
    tol = 1e-6 + rtol * np.abs(logits64) + noise * e32
    assert (err <= tol).all(), "%s logits: worst abs err %.3e (float32 oracle's own error %.3e)" % (what, err.max(), e32)
    l32 = abs(rloss - loss64)
    assert abs(loss - loss64) <= rtol * abs(loss64) + 1e-7 + noise * l32, "%s loss %r vs %r" % (what, loss, loss64)
