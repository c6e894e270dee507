"""CPU: the oracle's closed-form backward against torch autograd (float64), the FM identity, and
the optimizer restatements against hand-computed single-element updates."""
import numpy as np
import torch

from oracle.deepfm import OracleDeepFM, default_opt, init_weights

CAT = [dict(name="a", kind="identity", dtype="int32", num_buckets=5),
       dict(name="b", kind="identity", dtype="int32", num_buckets=3),
       dict(name="c", kind="identity", dtype="int32", num_buckets=7)]


def _cfg(**kw):
    cfg = dict(cat=CAT, num=["x0", "x1"], k=4, hidden=[8, 6], use_linear=True, use_mf=True, use_dnn=True,
               loss_reduction="mean", opt_deep=default_opt(), opt_linear=default_opt())
    cfg.update(kw)
    return cfg


def _data(B=16, seed=0):
    rng = np.random.default_rng(seed)
    ids = np.stack([rng.integers(-1, 5, B), rng.integers(0, 3, B), rng.integers(0, 7, B)], 1).astype(np.int32)
    x = rng.standard_normal((B, 2)).astype(np.float32)
    y = (rng.random(B) < 0.4).astype(np.float32)
    return ids, x, y


def test_backward_matches_autograd_fp64():
    for red, act in (("mean", "relu"), ("sum", "relu"), ("mean", "tanh"), ("mean", "sigmoid"), ("sum", "identity")):
        cfg = _cfg(loss_reduction=red, activation=act)
        w = init_weights(cfg, 3)
        w["lin"] = np.random.default_rng(0).standard_normal(w["lin"].shape).astype(np.float32) * 0.1
        m = OracleDeepFM(cfg, w, dtype=torch.float64)
        ids, x, y = _data()
        loss, z, g = m.grads(ids, x, y)
        leaves = {n: v.w.clone().requires_grad_(True) for n, v in m.vars.items()}
        for n, v in m.vars.items():
            v.w = leaves[n]
        z2 = m.forward(ids, x)
        lv = m.loss_vec(z2, torch.as_tensor(y, dtype=torch.float64))
        (lv.mean() if red == "mean" else lv.sum()).backward()
        for n, gr in g.items():
            if isinstance(gr, tuple):
                dense = torch.zeros_like(leaves[n])
                dense.index_add_(0, gr[1], gr[2])
                gr = dense
            assert torch.allclose(gr.reshape(leaves[n].shape), leaves[n].grad, atol=1e-12), n


def test_fm_identity():
    rng = np.random.default_rng(0)
    E = rng.standard_normal((5, 6, 4))
    s = E.sum(1)
    fm = 0.5 * ((s * s) - (E * E).sum(1)).sum(1)
    pair = sum((E[:, i] * E[:, j]).sum(1) for i in range(6) for j in range(i + 1, 6))
    assert np.allclose(fm, pair)


def test_nonlazy_adam_moves_untouched_rows():
    cfg = _cfg(use_dnn=False, num=[])
    w = init_weights(cfg, 1)
    m = OracleDeepFM(cfg, w)
    ids = np.array([[0, 0, 0]], np.int32)
    m.train_step(ids, None, np.array([1.0], np.float32))
    w1 = m.vars["emb"].w.clone()
    ids2 = np.array([[1, 1, 1]], np.int32)
    m.train_step(ids2, None, np.array([0.0], np.float32))
    # row 0 of field a was not in batch 2 but still moved (momentum of step 1)
    assert (m.vars["emb"].w[0] != w1[0]).any()
    # a never-touched row never moves
    assert torch.equal(m.vars["emb"].w[4], torch.tensor(w["emb"][4]))


def test_adagrad_ftrl_single_element():
    cfg = _cfg(use_dnn=False, use_mf=False, num=[], opt_linear=default_opt("Ftrl", 0.005), loss_reduction="sum")
    w = init_weights(cfg, 1)
    m = OracleDeepFM(cfg, w)
    ids = np.array([[2, 1, 3]], np.int32)
    m.train_step(ids, None, np.array([1.0], np.float32))
    g = 1 / (1 + np.exp(0.0)) - 1.0    # sigmoid(0) - 1
    acc = 0.1 + g * g
    lin = g - (np.sqrt(acc) - np.sqrt(0.1)) / 0.005 * 0.0
    wn = -lin / (np.sqrt(acc) / 0.005)
    assert abs(float(m.vars["lin"].w[2]) - wn) < 1e-7
