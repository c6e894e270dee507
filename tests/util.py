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
               dropout_seed=getattr(engine, "dropout_seed", 0))
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


def assert_state_close(engine_state, oracle_state, rtol=1e-5, atol=1e-7, what="", oracle_state64=None):
    """|got - ref| <= atol + rtol*|ref|.  When the float64 twin is given the reference is the float64
    value and the allowance grows by 4x the float32 oracle's own worst rounding error on that tensor
    (summation-order noise of hot rows is not a parity failure: TF's own order is unspecified)."""
    for name, ref in oracle_state.items():
        got = engine_state[name].reshape(ref.shape).astype(np.float64)
        if oracle_state64 is not None:
            ref64 = oracle_state64[name].astype(np.float64)
            e32 = float(np.abs(ref.astype(np.float64) - ref64).max()) if ref.size else 0.0
            err = np.abs(got - ref64)
            tol = atol + rtol * np.abs(ref64) + 4.0 * e32
        else:
            err = np.abs(got - ref)
            tol = atol + rtol * np.abs(ref)
        bad = err > tol
        assert not bad.any(), "%s %s: %d / %d elements off, worst abs %.3e (ref scale %.3e)" % (
            what, name, int(bad.sum()), ref.size, float(err.max()), float(np.abs(ref).max()))


def assert_step_close(loss, logits, pair, rloss, rlogits, rtol=1e-5, what=""):
    """loss / logits of one step against the oracle pair (float64 twin calibrates cancellation noise)."""
    loss64, logits64 = pair.last64
    e32 = float(np.abs(rlogits.astype(np.float64) - logits64).max())
    err = np.abs(logits.astype(np.float64) - logits64)
    tol = 1e-6 + rtol * np.abs(logits64) + 4.0 * e32
    assert (err <= tol).all(), "%s logits: worst abs err %.3e (float32 oracle's own error %.3e)" % (what, err.max(), e32)
    l32 = abs(rloss - loss64)
    assert abs(loss - loss64) <= rtol * abs(loss64) + 1e-7 + 4.0 * l32, "%s loss %r vs %r" % (what, loss, loss64)
