"""Shared helpers of the parity tests: build the CUDA engine and the oracle from one description."""
import numpy as np

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
            if k in r and k in d:
                r[k] = d[k]
        return r
    cfg = dict(cat=[dict(s) for s in engine.specs], num=[c.key for c in engine.num_columns], k=engine.k,
               hidden=list(engine.hidden), use_linear=engine.use_linear, use_mf=engine.use_mf,
               use_dnn=engine.use_dnn, loss_reduction=engine.loss_reduction, opt_deep=o(engine.opt_deep),
               opt_linear=o(engine.opt_linear))
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
    return OracleDeepFM(cfg, w), w


def rel_err(a, b, floor=1e-6):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    return float(np.max(np.abs(a - b) / (np.abs(b) + floor))) if a.size else 0.0


def assert_state_close(engine_state, oracle_state, rtol=1e-5, atol=1e-7, what=""):
    for name, ref in oracle_state.items():
        got = engine_state[name].reshape(ref.shape)
        bad = np.abs(got.astype(np.float64) - ref) > atol + rtol * np.abs(ref)
        assert not bad.any(), "%s %s: %d / %d elements off, worst abs %.3e (ref scale %.3e)" % (
            what, name, int(bad.sum()), ref.size, float(np.abs(got - ref).max()), float(np.abs(ref).max()))
