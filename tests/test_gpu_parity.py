"""GPU: parity of the CUDA path (through the C ABI) against the oracle on the same seeded inputs.

Bars: ids / fingerprints / sort order bit-exact; logits, loss, updated weights and optimizer
slots within 1e-5 relative (BASELINE.json north_star) with an absolute floor of 1e-7."""
import ctypes as C

import numpy as np
import pytest

from oracle import transforms
from oracle.farmhash import fingerprint64
from recommender_tensorflow_b200 import feature_column as fc
from recommender_tensorflow_b200 import synth
from recommender_tensorflow_b200.engine import DeepFMEngine, default_optimizer
from tests.util import assert_state_close, assert_step_close, make_pair, ml100k_columns

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _torch():
    import torch
    return torch


def test_device_fingerprint64_bit_exact():
    torch = _torch()
    from recommender_tensorflow_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    strs = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in list(range(0, 130)) * 3 + [200, 255, 256, 1000]]
    strs += [b"a", b"b", b"c", b"d", b"omar", b"stringer", b"marlo", b"101", b"201", b"301"]
    data, offs = transforms.pack_strings(strs)
    d = torch.from_numpy(data).cuda()
    o = torch.from_numpy(offs).cuda()
    out = torch.empty(len(strs), dtype=torch.int64, device="cuda")
    assert lib.dfm_test_fingerprint64(C.c_void_p(d.data_ptr()), C.c_void_p(o.data_ptr()), len(strs), C.c_void_p(out.data_ptr())) == 0
    got = out.cpu().numpy().view(np.uint64)
    ref = np.array([fingerprint64(s) for s in strs], dtype=np.uint64)
    assert (got == ref).all()


@pytest.mark.parametrize("n,bits", [(1, 8), (1000, 13), (4097, 13), (200000, 28), (1703936, 13), (300000, 32)])
def test_radix_sort_stable(n, bits):
    torch = _torch()
    from recommender_tensorflow_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 1 << bits, n, dtype=np.uint64).astype(np.uint32)
    if n > 10:
        keys[: n // 2] = keys[0]                     # a very hot key
    vals = np.arange(n, dtype=np.uint32)
    k = torch.from_numpy(keys.view(np.int32)).cuda()
    v = torch.from_numpy(vals.view(np.int32)).cuda()
    assert lib.dfm_test_sort_pairs(C.c_void_p(k.data_ptr()), C.c_void_p(v.data_ptr()), n, bits) == 0
    order = np.argsort(keys, kind="stable")
    assert (k.cpu().numpy().view(np.uint32) == keys[order]).all()
    assert (v.cpu().numpy().view(np.uint32) == vals[order]).all()


@pytest.mark.parametrize("algo", [1, 2])
@pytest.mark.parametrize("n,bits", [(1, 8), (1000, 13), (4097, 13), (200000, 28), (1703936, 28), (300000, 32), (50000, 5)])
def test_onesweep_sort_stable(n, bits, algo):
    """the one-sweep sort (decoupled look-back; algo 2: element count in device memory) is the same stable permutation"""
    torch = _torch()
    from recommender_tensorflow_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(n + bits)
    keys = rng.integers(0, 1 << bits, n, dtype=np.uint64).astype(np.uint32)
    if n > 10:
        keys[n // 4: n // 2] = keys[0]               # a very hot key: long runs inside and across tiles
    vals = np.arange(n, dtype=np.uint32)
    k = torch.from_numpy(keys.view(np.int32)).cuda()
    v = torch.from_numpy(vals.view(np.int32)).cuda()
    assert lib.dfm_test_sort_pairs_algo(C.c_void_p(k.data_ptr()), C.c_void_p(v.data_ptr()), n, bits, algo) == 0
    order = np.argsort(keys, kind="stable")
    assert (k.cpu().numpy().view(np.uint32) == keys[order]).all()
    assert (v.cpu().numpy().view(np.uint32) == vals[order]).all()


def _ml_engine(k=4, hidden=(16, 16), max_batch=4096, **kw):
    cols, dtypes = ml100k_columns()
    return DeepFMEngine(cols, (), embedding_size=k, hidden_units=hidden, max_batch=max_batch, feature_dtypes=dtypes, **kw)


def test_transform_ml100k_bit_exact():
    eng = _ml_engine()
    ml = synth.ML100K()
    feats, _ = ml.batch(3000, np.random.default_rng(5))
    feats["user_id"][:7] = -1                        # ignored int key -> empty bag
    feats["zipcode"][3] = b""                        # ignored string key
    feats["gender"][5] = b"X"                        # OOV bucket
    feats["occupation"][9] = b"a much longer occupation string than sixteen bytes"
    feats["occupation"][10] = b"x" * 100
    ids = eng.transform(feats)
    ref = transforms.transform(eng.specs, feats)
    assert ids.dtype == np.int32 and (ids == ref).all()
    ref_py = transforms.transform(eng.specs, {k: v[:50] for k, v in feats.items()}, use_c=False)
    assert (ids[:50] == ref_py).all()


def test_hash_bucket_sizes_bit_exact():
    """fingerprint % buckets on the device is a Barrett reduction with a per-column reciprocal (embed_kernels.cuh:
    mod_buckets); bucket counts from 1 to 5e7 incl. powers of two and their neighbours, string and int32 keys, against
    the oracle's plain 64-bit modulo (C and pure Python)."""
    from recommender_tensorflow_b200 import feature_column as fc
    sizes = [1, 2, 3, 5, 97, 255, 256, 257, 1000, 65535, 65536, 65537, (1 << 20) + 7, 9_999_991, 10_000_000, 50_000_017]
    cols, dtypes = [], {}
    for i, nb in enumerate(sizes):
        cols.append(fc.categorical_column_with_hash_bucket("s%d" % i, nb))
        cols.append(fc.categorical_column_with_hash_bucket("i%d" % i, nb, dtype="int32"))
        dtypes["i%d" % i] = "int32"
    eng = DeepFMEngine(cols, (), embedding_size=4, hidden_units=(16,), max_batch=4096, use_mf=False, use_dnn=False,
                       feature_dtypes=dtypes)
    rng = np.random.default_rng(77)
    B = 4096
    feats = {}
    for i in range(len(sizes)):
        lens = rng.integers(1, 40, B)
        feats["s%d" % i] = np.array([bytes(rng.integers(33, 127, int(n), dtype=np.uint8)) for n in lens], dtype=object)
        v = rng.integers(-2**31, 2**31 - 1, B).astype(np.int32)
        v[v == -1] = 12345
        feats["i%d" % i] = v
    ids = eng.transform(feats)
    ref = transforms.transform(eng.specs, feats)
    assert (ids == ref).all()
    ref_py = transforms.transform(eng.specs, {k: v[:40] for k, v in feats.items()}, use_c=False)
    assert (ids[:40] == ref_py).all()
    assert (ids >= 0).all() and (ids < np.asarray(eng.num_buckets)[None, :]).all()


def test_identity_out_of_range_is_an_error():
    from recommender_tensorflow_b200._lib import DfmError
    eng = _ml_engine()
    feats, _ = synth.ML100K().batch(64, np.random.default_rng(1))
    feats["war"][3] = 2
    with pytest.raises(DfmError) as ei:
        eng.transform(feats)
    assert ei.value.code == -3


def _assert_eval_logits(z, ora, ids):
    """forward-only logits against the oracle pair (float64 twin calibrates the FM cancellation noise)."""
    z32 = ora.o32.forward(ids).numpy().astype(np.float64)
    z64 = ora.o64.forward(ids).numpy()
    tol = 1e-6 + RTOL * np.abs(z64) + 4.0 * np.abs(z32 - z64).max()
    assert (np.abs(z.astype(np.float64) - z64) <= tol).all(), np.abs(z - z64).max()


def _run_steps(eng, ora, batches, what, noise=4.0, tc_noise=None):
    for i, (feats, y) in enumerate(batches):
        loss, logits = eng.train_step(feats, y, return_logits=True)
        rloss, rlogits = ora.train_step_raw(feats, y)
        assert_step_close(loss, logits, ora, rloss, rlogits, RTOL, "%s step %d" % (what, i), noise=noise)
    report = {}
    assert_state_close(eng.state(), ora.state(), RTOL, 1e-7, what, ora.state64(), report=report, tc_noise=tc_noise)
    print("%s: worst relative error per tensor (vs float64 oracle): %s" % (what, {k: "%.1e" % v for k, v in report.items()}))


def test_deepfm_ml100k_cfg1_steps():
    """configs[0]: DeepFM, ML-100K-shaped, k=4, hidden [16,16], batch 32 — 12 steps, every weight and slot."""
    eng = _ml_engine()
    ora, _ = make_pair(eng, seed=1)
    ml, rng = synth.ML100K(), np.random.default_rng(11)
    _run_steps(eng, ora, [ml.batch(32, rng) for _ in range(12)], "cfg1")
    assert eng.global_step == 12


def test_deepfm_hot_rows_batch4096():
    """large batch on tiny tables: every row is hot (piece reduction path)."""
    eng = _ml_engine(k=16, hidden=(32, 16), max_batch=4096)
    ora, _ = make_pair(eng, seed=2)
    ml, rng = synth.ML100K(), np.random.default_rng(12)
    _run_steps(eng, ora, [ml.batch(4096, rng) for _ in range(4)], "hot")


@pytest.mark.parametrize("hidden,batch", [((256, 128), 4096), ((64, 32), 1500), ((128,), 2048)])
def test_deepfm_tensor_core_tower(hidden, batch):
    """configs[2] shape: k=16, hidden [256,128] -> 3xTF32 tcgen05 GEMM tower, fp32-level parity."""
    eng = _ml_engine(k=16, hidden=hidden, max_batch=batch)
    ora, _ = make_pair(eng, seed=9)
    ml, rng = synth.ML100K(), np.random.default_rng(19)
    _run_steps(eng, ora, [ml.batch(batch, rng) for _ in range(3)], "tc-tower%r" % (hidden,))


def test_deepfm_cfg2_full_batch_65536():
    """BASELINE configs[2] AT ITS OWN BATCH SIZE: k=16, hidden [256,128], B = 65 536 -> 1 024 / 512 tiles on 148
    persistent CTAs (about 7 tiles per CTA: TMEM accumulator sets and the TMA ring wrap), 3 steps.

    The tcgen05 tower (3xTF32, truncating TMEM accumulation) carries ~5x the rounding noise of an fp32 CUDA-core sum
    (tests/test_gpu_tc_gemm.py, tools/debug_cfg2.py); Adam then turns the noise of a near-zero gradient element
    (update ~ lr g / (|g| + eps-hat), slope up to lr / eps-hat = 3e3) and of a ReLU unit that flips at |pre-activation|
    < 1e-6 into weight differences that no element-wise bar can hold - for the float32 oracle against its float64
    twin just as for the CUDA path.  So the step is pinned in three parts:
      (1) loss / logits of every step against the oracle: 1e-5 relative + 8x the float32 oracle's own deviation;
      (2) every optimizer slot (m, v: linear / quadratic in the gradients, no amplification) against the oracle:
          1e-5 relative + 8x the oracle's own worst deviation on that tensor;
      (3) every weight against the reference update formula applied to the engine's OWN slots,
          w_t = w_{t-1} - alpha_t m_t / (sqrt(v_t) + eps) for all rows (non-lazy Adam), to 2e-6 relative."""
    eng = _ml_engine(k=16, hidden=(256, 128), max_batch=65536)
    ora, _ = make_pair(eng, seed=21)
    ml, rng = synth.ML100K(), np.random.default_rng(22)
    prev = eng.state()
    for i in range(3):
        feats, y = ml.batch(65536, rng)
        loss, logits = eng.train_step(feats, y, return_logits=True)
        rloss, rlogits = ora.train_step_raw(feats, y)
        assert_step_close(loss, logits, ora, rloss, rlogits, RTOL, "cfg2-B65536 step %d" % i, noise=8.0)          # (1)
        st, r32, r64 = eng.state(), ora.state(), ora.state64()
        slots = {k: v for k, v in r32.items() if "/" in k}
        assert_state_close(st, slots, RTOL, 1e-7, "cfg2-B65536 step %d slots" % i, {k: r64[k] for k in slots}, tc_noise=8.0)   # (2)
        alpha = np.float32(ora.o32._alpha("deep"))
        for name in [k for k in r32 if "/" not in k]:                                                              # (3)
            m, v = st[name + "/m"].astype(np.float32), st[name + "/v"].astype(np.float32)
            want = prev[name].astype(np.float32) - (alpha * m) / (np.sqrt(v) + np.float32(1e-8))
            got = st[name]
            assert np.allclose(got, want, rtol=2e-6, atol=2e-9), (name, i, float(np.abs(got - want).max()))
        prev = st
    assert eng.global_step == 3


def test_deferred_adam_gap_1000_nontrivial_slots():
    """Rows that collect gradients for a few steps (non-trivial m, v) and are then left alone for ~1 000 steps, plus rows
    touched every step: the closed-form replay against the oracle's literal whole-table non-lazy update, every
    weight / slot, after the idle rows are touched again."""
    cats = [fc.categorical_column_with_identity("a", 128), fc.categorical_column_with_identity("b", 64)]
    eng = DeepFMEngine(cats, (), embedding_size=8, hidden_units=(8,), max_batch=32)
    ora, _ = make_pair(eng, seed=66)
    rng = np.random.default_rng(67)
    batches = []
    for step in range(1005):
        if step < 4 or step >= 1003:
            a = np.arange(32, dtype=np.int32)                    # rows 0..31: steps 1-4, then again after ~1 000 idle steps
        else:
            a = rng.integers(64, 128, 32).astype(np.int32)
        b = rng.integers(0, 64, 32).astype(np.int32)             # column b: rows touched (almost) every step
        batches.append(({"a": a, "b": b}, (rng.random(32) < 0.5).astype(np.float32)))
    _run_steps(eng, ora, batches, "gap-1000")


@pytest.mark.parametrize("name", ["RMSProp", "SGD", "Adagrad", "Ftrl"])
def test_other_optimizers_of_get_optimizer(name):
    """trainers/model_utils.py:57-66: every optimizer name the reference accepts."""
    from recommender_tensorflow_b200.trainers.model_utils import get_optimizer
    opt = get_optimizer(name, 0.01)
    eng = _ml_engine(opt_deep=opt, opt_linear=dict(opt), max_batch=512)
    ora, _ = make_pair(eng, seed=40)
    ml, rng = synth.ML100K(), np.random.default_rng(41)
    _run_steps(eng, ora, [ml.batch(300, rng) for _ in range(4)], "opt-" + name)


@pytest.mark.parametrize("activation,dropout", [("tanh", 0.0), ("sigmoid", 0.0), ("identity", 0.0), (None, 0.0), ("tanh", 0.2), ("sigmoid", 0.3)])
def test_activation_param(activation, dropout):
    """params["activation"] (trainers/deep_fm.py:22,100): anything but ReLU runs the general tower; forward, backward,
    every weight and slot against the oracle, with and without dropout."""
    eng = _ml_engine(k=8, hidden=(24, 16), max_batch=512, activation=activation, dropout=dropout, dropout_seed=11)
    assert eng.activation == (activation or "identity")
    ora, _ = make_pair(eng, seed=44)
    ml, rng = synth.ML100K(), np.random.default_rng(45)
    _run_steps(eng, ora, [ml.batch(300, rng) for _ in range(4)], "act-%s-%s" % (activation, dropout))


def test_activation_through_model_fn_and_bad_name():
    from recommender_tensorflow_b200.trainers import deep_fm, ml_100k

    def tanh(x):            # a callable named like the tf.nn function, as the reference passes it
        return np.tanh(x)
    fcs = ml_100k.get_feature_columns(embedding_size=4)
    params = {"categorical_columns": fcs["linear"], "embedding_size": 4, "hidden_units": [16, 16], "activation": tanh, "max_batch": 64,
              "tf_random_seed": 1}
    feats, y = synth.ML100K().batch(64, np.random.default_rng(3))
    spec = deep_fm.model_fn(feats, y, ml_100k.ModeKeys.TRAIN, params)
    assert np.isfinite(spec.loss) and params[deep_fm._ENGINE_KEY].activation == "tanh"
    with pytest.raises(ValueError):
        _ml_engine(activation="swish")


@pytest.mark.parametrize("hidden", [(16, 16), (64, 32), (20, 12)])
def test_dropout_all_tower_paths(hidden):
    """trainers/deep_fm.py:102-103 dropout after every hidden layer (reference CLI default 0.1): fused small-MLP,
    tcgen05 and SGEMM towers against the oracle's restatement of the same counter-based mask."""
    eng = _ml_engine(k=16, hidden=hidden, max_batch=1024, dropout=0.3, dropout_seed=77)
    ora, _ = make_pair(eng, seed=50)
    ml, rng = synth.ML100K(), np.random.default_rng(51)
    _run_steps(eng, ora, [ml.batch(1024, rng) for _ in range(3)], "dropout%r" % (hidden,))
    feats, _ = ml.batch(64, rng)                       # EVAL / PREDICT: no dropout
    z = eng.predict_logits(feats)
    _assert_eval_logits(z, ora, transforms.transform(eng.specs, feats))


def test_wide_deep_cfg2():
    """configs[1]: wide&deep = no FM, SUM loss, Adagrad (dnn side) + FTRL (linear side)."""
    eng = _ml_engine(use_mf=False, loss_reduction="sum", opt_deep=default_optimizer("Adagrad", 0.001),
                     opt_linear=default_optimizer("Ftrl", 0.005))
    ora, _ = make_pair(eng, seed=3)
    ml, rng = synth.ML100K(), np.random.default_rng(13)
    _run_steps(eng, ora, [ml.batch(4096, rng) for _ in range(4)], "wide&deep")


@pytest.mark.parametrize("use", [(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (0, 1, 1)])
def test_component_subsets(use):
    eng = _ml_engine(use_linear=bool(use[0]), use_mf=bool(use[1]), use_dnn=bool(use[2]), max_batch=512)
    ora, _ = make_pair(eng, seed=4)
    ml, rng = synth.ML100K(), np.random.default_rng(14)
    _run_steps(eng, ora, [ml.batch(257, rng) for _ in range(3)], "subset%r" % (use,))


def test_criteo_shaped_small_with_numerics():
    """Criteo-shaped: hashed string keys + 13 numeric features (numeric_embeddings path), k=16."""
    cats, nums = synth.criteo_columns(1000, n_cat=26, n_num=13)
    eng = DeepFMEngine(cats, nums, embedding_size=16, hidden_units=(16, 16), max_batch=2048)
    ora, _ = make_pair(eng, seed=5)
    rng = np.random.default_rng(15)
    _run_steps(eng, ora, [synth.criteo_batch(2048, rng, key_space=5000) for _ in range(4)], "criteo")


def test_deferred_adam_equals_literal_nonlazy_long_gaps():
    """Rows touched once and then left alone for 300 steps must follow TF's non-lazy Adam."""
    cats = [fc.categorical_column_with_identity("a", 64), fc.categorical_column_with_identity("b", 64)]
    eng = DeepFMEngine(cats, (), embedding_size=8, hidden_units=(8,), max_batch=16)
    ora, _ = make_pair(eng, seed=6)
    rng = np.random.default_rng(16)
    batches = []
    for step in range(300):
        if step == 0:
            a = np.arange(16, dtype=np.int32)          # rows 0..15 touched at step 1 only
        else:
            a = rng.integers(32, 64, 16).astype(np.int32)
        b = rng.integers(0, 64, 16).astype(np.int32)
        batches.append(({"a": a, "b": b}, (rng.random(16) < 0.5).astype(np.float32)))
    _run_steps(eng, ora, batches, "deferred-adam")


def test_bit_identical_reruns():
    """Determinism: two engines fed the same batches end in bit-identical state."""
    states = []
    for _ in range(2):
        eng = _ml_engine(k=16, hidden=(32, 16), max_batch=4096)
        make_pair(eng, seed=7)
        ml, rng = synth.ML100K(), np.random.default_rng(17)
        for _ in range(3):
            eng.train_step(*ml.batch(4096, rng))
        states.append(eng.state())
    for name in states[0]:
        assert (states[0][name].view(np.uint32) == states[1][name].view(np.uint32)).all(), name


def test_eval_forward_matches_oracle():
    eng = _ml_engine()
    ora, _ = make_pair(eng, seed=8)
    ml, rng = synth.ML100K(), np.random.default_rng(18)
    for _ in range(3):
        feats, y = ml.batch(64, rng)
        eng.train_step(feats, y)
        ora.train_step_raw(feats, y)
    feats, _ = ml.batch(100, rng)
    z = eng.predict_logits(feats)
    _assert_eval_logits(z, ora, transforms.transform(eng.specs, feats))
    assert eng.global_step == 3


def test_checkpoint_round_trip_tf_names(tmp_path):
    """save (TF-1.12 variable names, slots, global_step) -> fresh engine -> load -> continue == uninterrupted."""
    a = _ml_engine()
    make_pair(a, seed=30)
    ml, rng = synth.ML100K(), np.random.default_rng(31)
    batches = [ml.batch(64, rng) for _ in range(9)]
    for f, y in batches[:5]:
        a.train_step(f, y)
    path = a.save_checkpoint(str(tmp_path / "model.ckpt-5.npz"))
    data = np.load(path)
    assert int(data["global_step"]) == 5
    assert data["input_layer/input_layer/user_id_embedding/embedding_weights"].shape == (1000, 4)
    assert data["linear/linear_model/age_bucketized/weights/Adam_1"].shape == (7, 1)
    assert data["dnn/dnn/hiddenlayer_0/dense/kernel"].shape == (104, 16)
    b = _ml_engine()
    b.load_checkpoint(path)
    assert b.global_step == 5
    for f, y in batches[5:]:
        la, lb = a.train_step(f, y), b.train_step(f, y)
        assert abs(la - lb) <= 1e-6 * abs(la) + 1e-7
    sa, sb = a.state(), b.state()
    for name in sa:
        assert np.allclose(sa[name], sb[name], rtol=2e-6, atol=1e-8), name


def test_estimator_model_fn_facade(tmp_path):
    """model_fn / Estimator keep the reference's call shape (trainers/deep_fm.py:11, 153-178)."""
    from recommender_tensorflow_b200.trainers import deep_fm, ml_100k
    csv_path = str(tmp_path / "train.csv")
    ml_100k.write_synthetic_csv(csv_path, 600)
    fc_ = ml_100k.get_feature_columns(embedding_size=4)
    est = deep_fm.Estimator(model_fn=deep_fm.model_fn, model_dir=str(tmp_path / "ckpt"), params={
        "categorical_columns": fc_["linear"], "use_linear": True, "use_mf": True, "use_dnn": True,
        "embedding_size": 4, "hidden_units": [16, 16], "dropout": 0, "max_batch": 64})
    loss = est.train(ml_100k.get_input_fn(csv_path, batch_size=32, seed=0), max_steps=20, log_every=0)
    assert np.isfinite(loss) and est.engine.global_step == 20
    m = est.evaluate(ml_100k.get_input_fn(csv_path, ml_100k.ModeKeys.EVAL, batch_size=64))
    assert set(["accuracy", "auc", "auc_precision_recall", "average_loss", "label/mean", "prediction/mean"]) <= set(m)
    assert est.latest_checkpoint().endswith("model.ckpt-20.npz")
    # a second estimator on the same model_dir resumes from step 20
    est2 = deep_fm.Estimator(model_fn=deep_fm.model_fn, model_dir=str(tmp_path / "ckpt"), params=dict(est.params, **{deep_fm._ENGINE_KEY: None}))
    est2.train(ml_100k.get_input_fn(csv_path, batch_size=32, seed=1), max_steps=25, log_every=0)
    assert est2.engine.global_step == 25
    with pytest.raises(ValueError):
        deep_fm.model_fn({"user_id": np.zeros(4, np.int32)}, np.zeros(4), ml_100k.ModeKeys.TRAIN,
                         {"categorical_columns": [], "numeric_columns": []})


def test_canned_estimator_mirrors(tmp_path):
    """trainers/linear.py, trainers/deep.py, trainers/linear_deep.py: the canned estimators' train/evaluate loops."""
    from recommender_tensorflow_b200.trainers import deep, linear, linear_deep, ml_100k
    csv_path = str(tmp_path / "train.csv")
    ml_100k.write_synthetic_csv(csv_path, 400)
    fcs = ml_100k.get_feature_columns(embedding_size=4)
    for est in (linear.LinearClassifier(fcs["linear"], max_batch=64),
                deep.DNNClassifier([16, 16], fcs["deep"], max_batch=64),
                linear_deep.DNNLinearCombinedClassifier(linear_feature_columns=fcs["linear"], dnn_feature_columns=fcs["deep"],
                                                        dnn_hidden_units=[16, 16], max_batch=64)):
        loss = est.train(ml_100k.get_input_fn(csv_path, batch_size=32, seed=0), max_steps=10)
        assert np.isfinite(loss) and est.engine.global_step == 10
        m = est.evaluate(ml_100k.get_input_fn(csv_path, ml_100k.ModeKeys.EVAL, batch_size=64))
        assert 0.0 <= m["auc"] <= 1.0 and np.isfinite(m["average_loss"])


def test_canned_estimators_initialise_train_and_resume(tmp_path):
    """The canned mirrors run TF's variable initialisers (a zero tower would stay dead under ReLU), learn, and resume
    from model_dir like an Estimator: a fresh instance on the same model_dir predicts what the trained one predicted."""
    from recommender_tensorflow_b200.trainers import deep, linear_deep, ml_100k
    csv_path = str(tmp_path / "train.csv")
    ml_100k.write_synthetic_csv(csv_path, 2000)
    fcs = ml_100k.get_feature_columns(embedding_size=4)
    builders = {
        "deep": lambda d: deep.DNNClassifier([16, 16], fcs["deep"], model_dir=d, max_batch=64, tf_random_seed=3),
        "linear_deep": lambda d: linear_deep.DNNLinearCombinedClassifier(model_dir=d, linear_feature_columns=fcs["linear"], dnn_feature_columns=fcs["deep"],
                                                                        dnn_hidden_units=[16, 16], max_batch=64, tf_random_seed=3)}
    for name, build in builders.items():
        d = str(tmp_path / name)
        est = build(d)
        w0 = est.engine.get_tensor("W0")
        assert np.abs(w0).max() > 0 and np.abs(est.engine.get_tensor("emb")).max() > 0, "initialisers did not run"
        eval_fn = ml_100k.get_input_fn(csv_path, ml_100k.ModeKeys.EVAL, batch_size=64)
        before = est.evaluate(eval_fn)
        est.train(ml_100k.get_input_fn(csv_path, batch_size=32, seed=0), max_steps=300)
        after = est.evaluate(eval_fn)
        assert np.abs(est.engine.get_tensor("W0") - w0).max() > 0, "the tower did not train"
        assert after["average_loss"] < before["average_loss"], (name, before["average_loss"], after["average_loss"])
        fresh = build(d)                                           # same model_dir: resumes from model.ckpt-300.npz
        again = fresh.evaluate(eval_fn)
        assert fresh.engine.global_step == 300 and again["average_loss"] == after["average_loss"] and again["auc"] == after["auc"]


def test_estimator_predict_restores_or_raises(tmp_path):
    """tf.estimator.Estimator.predict restores the latest checkpoint of model_dir and fails when there is none."""
    from recommender_tensorflow_b200.trainers import deep_fm, ml_100k
    csv_path = str(tmp_path / "train.csv")
    ml_100k.write_synthetic_csv(csv_path, 600)
    fcs = ml_100k.get_feature_columns(embedding_size=4)
    params = {"categorical_columns": fcs["linear"], "embedding_size": 4, "hidden_units": [16, 16], "max_batch": 64, "tf_random_seed": 5}
    d = str(tmp_path / "job")
    est = deep_fm.Estimator(deep_fm.model_fn, model_dir=d, params=params)
    est.train(ml_100k.get_input_fn(csv_path, batch_size=32, seed=0), max_steps=20)
    eval_fn = ml_100k.get_input_fn(csv_path, ml_100k.ModeKeys.EVAL, batch_size=64)
    want = [p["logits"][0] for p in est.predict(eval_fn)]
    fresh = deep_fm.Estimator(deep_fm.model_fn, model_dir=d, params=params)
    got = [p["logits"][0] for p in fresh.predict(eval_fn)]
    assert fresh.engine.global_step == 20 and got == want
    empty = deep_fm.Estimator(deep_fm.model_fn, model_dir=str(tmp_path / "nothing"), params=params)
    with pytest.raises(ValueError):
        next(iter(empty.predict(eval_fn)))
    import json, os
    recs = [json.loads(l) for l in open(os.path.join(d, "summaries.jsonl"))]          # layer_summary side outputs, step 1
    assert recs and recs[0]["step"] == 1 and "deep_fm/logits" in recs[0]["summaries"]


# ------------------------------------------------------------------ edge cases and full-size properties
def test_edge_batches():
    """batch of 1, a batch whose hashed / vocab columns are all empty bags, and a full max_batch."""
    eng = _ml_engine(max_batch=96)
    ora, _ = make_pair(eng, seed=60)
    ml, rng = synth.ML100K(), np.random.default_rng(61)
    b1 = ml.batch(1, rng)
    f2, y2 = ml.batch(33, rng)
    f2["user_id"][:] = -1
    f2["zipcode"] = np.array([b""] * 33, dtype=object)
    f2["gender"] = np.array([b""] * 33, dtype=object)
    b3 = ml.batch(96, rng)
    _run_steps(eng, ora, [b1, (f2, y2), b3, b1], "edge")


def test_argument_errors():
    from recommender_tensorflow_b200._lib import DfmError
    eng = _ml_engine(max_batch=16)
    ml, rng = synth.ML100K(), np.random.default_rng(62)
    feats, y = ml.batch(17, rng)
    with pytest.raises(DfmError) as ei:
        eng.train_step(feats, y)                       # batch_size > max_batch
    assert ei.value.code == -1
    feats, y = ml.batch(8, rng)
    del feats["zipcode"]
    with pytest.raises(KeyError):
        eng.train_step(feats, y)                       # missing feature column (TF: KeyError from features dict)
    with pytest.raises(ValueError):
        _ml_engine(use_linear=False, use_mf=False, use_dnn=False)
    with pytest.raises(DfmError):
        _ml_engine(k=12)                                # unsupported embedding_size


def test_full_size_transform_and_sort_properties():
    """BASELINE batch size (65 536 x 26 lookups): ids equal the C oracle; the sorted lookup list is a stable
    permutation (checked through the C-ABI sort hook on the step's own key space)."""
    eng = _ml_engine(k=16, hidden=(16, 16), max_batch=65536)
    feats, _ = synth.ML100K().batch(65536, np.random.default_rng(63))
    ids = eng.transform(feats)
    assert (ids == transforms.transform(eng.specs, feats)).all()
    assert ids.min() >= 0 and (ids < np.asarray(eng.num_buckets)[None, :]).all()


def test_full_size_step_determinism_and_fm_identity():
    """At B = 65 536: two engines give bit-identical losses / weights, and with linear + MF only the logits obey
    the FM identity  z = sum_f w_f + b + sum_{i<j} <v_i, v_j>  (checked on a sample of rows in float64)."""
    ml = synth.ML100K()
    feats, y = ml.batch(65536, np.random.default_rng(64))
    res = []
    for _ in range(2):
        eng = _ml_engine(k=16, use_dnn=False, max_batch=65536)
        _, w = make_pair(eng, seed=65)
        loss, logits = eng.train_step(feats, y, return_logits=True)
        res.append((loss, logits, eng.get_tensor("emb")))
    assert res[0][0] == res[1][0] and (res[0][1] == res[1][1]).all()
    assert (res[0][2].view(np.uint32) == res[1][2].view(np.uint32)).all()
    ids = transforms.transform(eng.specs, feats)[:200]
    rows = ids + eng.row_offsets[:-1][None, :]
    E = w["emb"].astype(np.float64)[rows]                       # [200, 26, 16]
    pair = 0.5 * ((E.sum(1) ** 2).sum(1) - (E ** 2).sum((1, 2)))
    z = w["lin"].astype(np.float64)[rows].sum(1) + float(w["bias"][0]) + pair
    assert np.allclose(res[0][1][:200], z, rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------ multi-hot (multivalent) columns
def _bag_setup(B, rng, ml):
    """ML-100K-shaped features where the 19 one-hot genre columns are ONE multi-hot `genres` column (identity, 19
    buckets, up to 6 ids per sample, -1 padded) plus a multivalent string column `tags` (hashed, 3 slots, '' padded)."""
    from recommender_tensorflow_b200.trainers import ml_100k
    base = [c for c in ml_100k.get_feature_columns()["linear"] if c.key not in synth.GENRE]
    cols = base + [fc.categorical_column_with_identity("genres", 19), fc.categorical_column_with_hash_bucket("tags", 50)]
    feats, y = ml.batch(B, rng)
    g = np.full((B, 6), -1, np.int32)
    for j, name in enumerate(synth.GENRE):
        on = feats.pop(name) > 0
        pos = (g >= 0).sum(1)
        ok = on & (pos < 6)
        g[np.where(ok)[0], pos[ok]] = j
    feats["genres"] = g
    words = [b"", b"noir", b"cult", b"classic", b"indie", b"blockbuster-of-the-year-long-tag"]
    tags = [words[i] for i in rng.integers(0, len(words), B * 3)]
    for i in range(0, B * 3, 7):
        tags[i] = b""                                  # some empty slots, some fully empty bags
    feats["tags"] = np.array(tags, dtype=object).reshape(B, 3)
    return cols, feats, y


def test_multi_hot_pooling_matches_oracle():
    """embedding_column mean / linear_model sum over a multivalent column (BASELINE north_star: multi-hot genre pooling)."""
    from recommender_tensorflow_b200.trainers import ml_100k
    ml, rng = synth.ML100K(), np.random.default_rng(70)
    cols, _, _ = _bag_setup(4, rng, ml)
    eng = DeepFMEngine(cols, (), embedding_size=8, hidden_units=(16, 16), max_batch=600, feature_dtypes=ml_100k.FEATURE_DTYPES,
                       multivalent={"genres": 6, "tags": 3})
    assert eng.n_slots == len(cols) - 2 + 6 + 3
    ora, _ = make_pair(eng, seed=71)
    batches = []
    for _ in range(4):
        _, feats, y = _bag_setup(600, rng, ml)
        batches.append((feats, y))
    ids = eng.transform(batches[0][0])
    assert ids.shape == (600, eng.n_slots) and (ids == transforms.transform(eng.specs, batches[0][0])).all()
    _run_steps(eng, ora, batches, "multi-hot")


def test_model_learns_a_simple_rule():
    """End-to-end sanity beyond oracle parity: labels that are a deterministic function of two categorical features
    are learnt (loss falls well below ln 2, AUC -> 1) by the DeepFM step with the reference optimizer."""
    from recommender_tensorflow_b200.engine import default_optimizer
    from recommender_tensorflow_b200.trainers.model_utils import get_binary_metrics
    cols, dtypes = ml100k_columns()
    eng = DeepFMEngine(cols, (), embedding_size=8, hidden_units=(32, 16), max_batch=2048, feature_dtypes=dtypes,
                       opt_deep=default_optimizer("Adam", 0.01), opt_linear=default_optimizer("Adam", 0.01))
    eng.init_random(123)
    ml, rng = synth.ML100K(), np.random.default_rng(7)

    def batch():
        feats, _ = ml.batch(2048, rng)
        y = ((feats["action"] == 1) ^ (feats["gender"] == b"M")).astype(np.float32)     # needs the interaction, not only the linear part
        return feats, y
    first = None
    for step in range(300):
        feats, y = batch()
        loss = eng.train_step(feats, y)
        first = loss if first is None else first
    feats, y = batch()
    m = get_binary_metrics(y, eng.predict_logits(feats))
    assert first > 0.5 and loss < 0.15, (first, loss)
    assert m["auc"] > 0.99 and m["accuracy"] > 0.97, m


def test_resume_at_a_late_step_and_alpha_history_growth():
    """dfm_set_global_step (checkpoint restore) at step 65 530, then steps across the point where the device-side
    alpha_t history doubles (65 534): beta powers, alpha_t and the deferred catch-up keep following the oracle."""
    cats = [fc.categorical_column_with_hash_bucket("a", 50000, "int32"), fc.categorical_column_with_hash_bucket("b", 50000, "int32")]
    eng = DeepFMEngine(cats, (), embedding_size=8, hidden_units=(8,), max_batch=64)      # 100 000 rows: the deferred (touched-rows) path
    ora, _ = make_pair(eng, seed=60)
    T = 65530
    eng._check(eng.lib.dfm_set_global_step(eng.h, T))
    for o in (ora.o32, ora.o64):
        o.t = T
        for grp in ("deep", "linear"):
            b1 = b2 = np.float32(1.0)
            for _ in range(T):              # the float32 running products TF keeps in beta1_power / beta2_power
                b1 = np.float32(b1 * np.float32(o.opt[grp]["beta1"]))
                b2 = np.float32(b2 * np.float32(o.opt[grp]["beta2"]))
            o.pow[grp] = [b1, b2]
    rng = np.random.default_rng(61)
    batches = []
    for step in range(12):
        a = (np.arange(64) if step == 0 else rng.integers(1000, 2000, 64)).astype(np.int32)   # rows of step 1 idle afterwards
        batches.append(({"a": a, "b": rng.integers(0, 3000, 64).astype(np.int32)}, (rng.random(64) < 0.5).astype(np.float32)))
    _run_steps(eng, ora, batches, "late-step")
    assert eng.global_step == T + 12


@pytest.mark.parametrize("batch,hidden", [(4096, (64, 32)), (32, (16, 16)), (1024, (16, 16))])
def test_batch_prefetch_changes_nothing(batch, hidden):
    """dfm_prefetch_batch: ids / sort / segments of the next batch computed on the side stream while the current
    step runs -> the same bits as computing them inside the step (also when a prefetch is dropped)."""
    ml, rng = synth.ML100K(), np.random.default_rng(70)
    data = [ml.batch(batch, rng) for _ in range(7)]
    states = []
    for prefetch in (False, True):
        eng = _ml_engine(k=8, hidden=hidden, max_batch=batch)
        eng.init_random(9)
        pbs = [eng.pack(f, y, device=True) for f, y in data]
        losses = []
        for i, pb in enumerate(pbs):
            losses.append(eng.train_step_device(pb))
            if prefetch and i + 1 < len(pbs):
                eng.prefetch(pbs[i + 1] if i != 3 else pbs[0])      # step 4 gets a prefetch for the wrong batch: dropped
        eng.sync()
        states.append(([float(x.item()) for x in losses], eng.state()))
    assert states[0][0] == states[1][0]
    for name in states[0][1]:
        assert np.array_equal(states[0][1][name], states[1][1][name]), name
