"""GPU: the row-sharded step (world = 2, 4 virtual ranks on one GPU) equals the single-GPU step and the
oracle: ids exact by construction, weights / slots / loss within the fp32 tolerance (only the
association of the gradient sums changes)."""
import numpy as np
import pytest

from recommender_tensorflow_b200 import synth
from recommender_tensorflow_b200.engine import DeepFMEngine
from recommender_tensorflow_b200.sharded import VirtualCluster
from tests.util import assert_state_close, assert_step_close, make_pair, ml100k_columns

pytestmark = pytest.mark.gpu


def _names(eng):
    out = []
    for v in eng.variable_names():
        out.append(v)
        out += [v + "/" + s for s in eng.slot_names(v)]
    return out


@pytest.mark.parametrize("p2p", [False, True])
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_criteo_equals_oracle(world, p2p):
    cats, nums = synth.criteo_columns(500, n_cat=26, n_num=13)
    kw = dict(embedding_size=16, hidden_units=(16, 16))
    ref = DeepFMEngine(cats, nums, max_batch=1024, **kw)
    ora, w = make_pair(ref, seed=21)
    engs = [DeepFMEngine(cats, nums, max_batch=1024 // world, rank=r, world=world, **kw) for r in range(world)]
    for e in engs:
        e.set_weights_sharded(w)
    vc = VirtualCluster(engs, p2p=p2p)
    rng = np.random.default_rng(22)
    per = 1024 // world
    for step in range(4):
        feats, y = synth.criteo_batch(1024, rng, key_space=3000)
        pbs = []
        for r, e in enumerate(engs):
            fr = {}
            for k, v in feats.items():
                if isinstance(v, tuple):
                    data, offs = v
                    o = offs[r * per:(r + 1) * per + 1]
                    fr[k] = (data[o[0]:o[-1]].copy(), (o - o[0]).astype(np.int32))
                else:
                    fr[k] = v[r * per:(r + 1) * per]
            pbs.append(e.pack(fr, y[r * per:(r + 1) * per], device=True))
        loss, logits = vc.train_step(pbs, return_logits=True)
        rloss, rlogits = ora.train_step_raw(feats, y)
        assert_step_close(loss, logits, ora, rloss, rlogits, 1e-5, "sharded w=%d step %d" % (world, step))
    assert_state_close(vc.state(_names(engs[0])), ora.state(), 1e-5, 1e-7, "sharded w=%d" % world, ora.state64())
    assert all(e.global_step == 4 for e in engs)


@pytest.mark.parametrize("p2p", [False, True])
def test_sharded_ml100k_hot_rows_world2(p2p):
    cols, dtypes = ml100k_columns()
    kw = dict(embedding_size=4, hidden_units=(16, 16), feature_dtypes=dtypes)
    ref = DeepFMEngine(cols, (), max_batch=2048, **kw)
    ora, w = make_pair(ref, seed=23)
    engs = [DeepFMEngine(cols, (), max_batch=1024, rank=r, world=2, **kw) for r in range(2)]
    for e in engs:
        e.set_weights_sharded(w)
    vc = VirtualCluster(engs, p2p=p2p)
    ml, rng = synth.ML100K(), np.random.default_rng(24)
    for step in range(3):
        feats, y = ml.batch(2048, rng)
        pbs = [e.pack({k: v[r * 1024:(r + 1) * 1024] for k, v in feats.items()}, y[r * 1024:(r + 1) * 1024], device=True)
               for r, e in enumerate(engs)]
        loss, logits = vc.train_step(pbs, return_logits=True)
        rloss, rlogits = ora.train_step_raw(feats, y)
        assert_step_close(loss, logits, ora, rloss, rlogits, 1e-5, "sharded-ml step %d" % step)
    assert_state_close(vc.state(_names(engs[0])), ora.state(), 1e-5, 1e-7, "sharded-ml", ora.state64())


@pytest.mark.parametrize("p2p", [False, True])
def test_sharded_multi_hot_world2(p2p):
    """multivalent columns through the row-sharded path (slots are looked up at their owners, pooled locally)."""
    from recommender_tensorflow_b200.trainers import ml_100k
    from tests.test_gpu_parity import _bag_setup
    ml, rng = synth.ML100K(), np.random.default_rng(80)
    cols, _, _ = _bag_setup(4, rng, ml)
    kw = dict(embedding_size=8, hidden_units=(16, 16), feature_dtypes=ml_100k.FEATURE_DTYPES, multivalent={"genres": 6, "tags": 3})
    ref = DeepFMEngine(cols, (), max_batch=512, **kw)
    ora, w = make_pair(ref, seed=81)
    engs = [DeepFMEngine(cols, (), max_batch=256, rank=r, world=2, **kw) for r in range(2)]
    for e in engs:
        e.set_weights_sharded(w)
    vc = VirtualCluster(engs, p2p=p2p)
    for step in range(3):
        _, feats, y = _bag_setup(512, rng, ml)
        pbs = [e.pack({k: v[r * 256:(r + 1) * 256] for k, v in feats.items()}, y[r * 256:(r + 1) * 256], device=True)
               for r, e in enumerate(engs)]
        loss, logits = vc.train_step(pbs, return_logits=True)
        rloss, rlogits = ora.train_step_raw(feats, y)
        assert_step_close(loss, logits, ora, rloss, rlogits, 1e-5, "sharded-bags step %d" % step)
    assert_state_close(vc.state(_names(engs[0])), ora.state(), 1e-5, 1e-7, "sharded-bags", ora.state64())


def test_p2p_exchange_bit_identical_to_collectives():
    """The fused peer-memory exchange (dfm_xchg_*: stores into the peers' regions, flags, no collective) only changes
    WHERE rows travel, never a value: after the same steps the shards of a fused-exchange cluster equal the shards of
    the all_to_all cluster bit for bit."""
    cats, nums = synth.criteo_columns(2000, n_cat=8, n_num=4)
    kw = dict(embedding_size=16, hidden_units=(32, 16))
    world, per = 4, 512
    clusters = []
    for p2p in (False, True):
        engs = [DeepFMEngine(cats, nums, max_batch=per, rank=r, world=world, **kw) for r in range(world)]
        for e in engs:
            e.init_random(5)
        clusters.append(VirtualCluster(engs, p2p=p2p))
    rng = np.random.default_rng(90)
    for step in range(5):
        feats, y = synth.criteo_batch(world * per, rng, key_space=4000)
        losses = []
        for vc in clusters:
            pbs = []
            for r, e in enumerate(vc.engs):
                fr = {}
                for k, v in feats.items():
                    if isinstance(v, tuple):
                        data, offs = v
                        o = offs[r * per:(r + 1) * per + 1]
                        fr[k] = (data[o[0]:o[-1]].copy(), (o - o[0]).astype(np.int32))
                    else:
                        fr[k] = v[r * per:(r + 1) * per]
                pbs.append(e.pack(fr, y[r * per:(r + 1) * per], device=True))
            losses.append(vc.train_step(pbs))
        assert losses[0] == losses[1], step
    for ea, eb in zip(clusters[0].engs, clusters[1].engs):
        ea.flush(); eb.flush()
        for n in _names(ea):
            assert np.array_equal(ea.get_tensor(n), eb.get_tensor(n)), n


def test_p2p_ipc_two_processes():
    """Real thing: two processes, two GPUs, exchange regions mapped through CUDA IPC, kernels meeting through flags in
    peer memory; the fused-exchange trainer must match the NCCL all_to_all trainer bit for bit
    (tests/p2p_worker.py).  Needs >= 2 GPUs."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(root, "tests", "p2p_worker.py")]
    env = dict(os.environ, NCCL_DEBUG="WARN", PYTHONPATH=root)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=root)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "P2P_OK" in r.stdout


@pytest.mark.parametrize("p2p", [False, True])
def test_sharded_forward_only_equals_single_gpu(p2p):
    """EVAL / PREDICT on a row-sharded model (dfm_shard_forward): same weights -> the logits of the unsharded model,
    bit for bit (the rows travel, the arithmetic is the same), before and after training steps; no state change."""
    cats, nums = synth.criteo_columns(3000, n_cat=10, n_num=5)
    kw = dict(embedding_size=8, hidden_units=(32, 16))
    world, per = 2, 300
    ref = DeepFMEngine(cats, nums, max_batch=world * per, **kw)
    ora, w = make_pair(ref, seed=31)
    engs = [DeepFMEngine(cats, nums, max_batch=per, rank=r, world=world, **kw) for r in range(world)]
    for e in engs:
        e.set_weights_sharded(w)
    vc = VirtualCluster(engs, p2p=p2p)
    rng = np.random.default_rng(32)

    def split(feats, y=None):
        pbs = []
        for r, e in enumerate(engs):
            fr = {}
            for k, v in feats.items():
                if isinstance(v, tuple):
                    data, offs = v
                    o = offs[r * per:(r + 1) * per + 1]
                    fr[k] = (data[o[0]:o[-1]].copy(), (o - o[0]).astype(np.int32))
                else:
                    fr[k] = v[r * per:(r + 1) * per]
            pbs.append(e.pack(fr, None if y is None else y[r * per:(r + 1) * per], device=True))
        return pbs

    feats, y = synth.criteo_batch(world * per, rng, key_space=5000)
    assert np.array_equal(vc.predict_logits(split(feats)), ref.predict_logits(feats))
    for _ in range(3):
        f2, y2 = synth.criteo_batch(world * per, rng, key_space=5000)
        vc.train_step(split(f2, y2))
        ref.train_step(f2, y2)
    before = vc.state(_names(engs[0]))
    got, want = vc.predict_logits(split(feats)), ref.predict_logits(feats)
    assert np.allclose(got, want, rtol=1e-5, atol=1e-6)        # trained separately: fp32 re-association of the gradient sums
    after = vc.state(_names(engs[0]))
    for k in before:
        assert np.array_equal(before[k], after[k]), k
    assert all(e.global_step == 3 for e in engs)


def test_init_random_is_keyed_on_global_rows():
    """dfm_init_random(seed) on a row-sharded model gives exactly the rows of the unsharded model with the same seed
    (and the same dense tower on every rank): `eng.init_random(s)` on every rank is all a multi-GPU host has to do."""
    cats, nums = synth.criteo_columns(777, n_cat=5, n_num=3)
    kw = dict(embedding_size=8, hidden_units=(16, 16), max_batch=64)
    ref = DeepFMEngine(cats, nums, **kw)
    ref.init_random(42)
    world = 3
    engs = [DeepFMEngine(cats, nums, rank=r, world=world, **kw) for r in range(world)]
    for e in engs:
        e.init_random(42)
    full = ref.get_tensor("emb")
    for r, e in enumerate(engs):
        assert np.array_equal(e.get_tensor("emb"), full[r::world])
        for n in ("W0", "W1", "Wo", "num_emb"):
            assert np.array_equal(e.get_tensor(n), ref.get_tensor(n)), n


def test_sharded_checkpoint_shards_resume_and_merge(tmp_path):
    """A row-sharded run checkpoints one shard file per rank; fresh sharded engines resume from them bit for bit, and the
    host-side merge of the shards is a TF-named checkpoint an unsharded engine loads (same logits as the sharded model)."""
    cats, nums = synth.criteo_columns(700, n_cat=26, n_num=13)
    kw = dict(embedding_size=16, hidden_units=(16, 16))
    world, per = 2, 256
    engs = [DeepFMEngine(cats, nums, max_batch=per, rank=r, world=world, **kw) for r in range(world)]
    for e in engs:
        e.init_random(9)
    vc = VirtualCluster(engs, p2p=True)
    rng = np.random.default_rng(90)

    def split(feats, y):
        pbs = []
        for r, e in enumerate(engs):
            fr = {}
            for k, v in feats.items():
                if isinstance(v, tuple):
                    data, offs = v
                    o = offs[r * per:(r + 1) * per + 1]
                    fr[k] = (data[o[0]:o[-1]].copy(), (o - o[0]).astype(np.int32))
                else:
                    fr[k] = v[r * per:(r + 1) * per]
            pbs.append((fr, y[r * per:(r + 1) * per]))
        return pbs
    batches = [synth.criteo_batch(world * per, rng, key_space=5000) for _ in range(4)]
    for feats, y in batches[:3]:
        vc.train_step([e.pack(f, yy, device=True) for e, (f, yy) in zip(engs, split(feats, y))])
    paths = [e.save_checkpoint_shard(str(tmp_path / ("model.ckpt-3.shard-%d-of-%d.npz" % (r, world)))) for r, e in enumerate(engs)]
    sums = [e.state_checksum() for e in engs]
    loss_next = vc.train_step([e.pack(f, yy, device=True) for e, (f, yy) in zip(engs, split(*batches[3]))])
    # resume: fresh sharded engines from the shard files
    engs2 = [DeepFMEngine(cats, nums, max_batch=per, rank=r, world=world, **kw) for r in range(world)]
    for e, p in zip(engs2, paths):
        e.load_checkpoint_shard(p)
    assert [e.state_checksum() for e in engs2] == sums and all(e.global_step == 3 for e in engs2)
    vc2 = VirtualCluster(engs2, p2p=True)
    loss2 = vc2.train_step([e.pack(f, yy, device=True) for e, (f, yy) in zip(engs2, split(*batches[3]))])
    assert loss2 == loss_next
    with pytest.raises(ValueError):
        engs2[0].load_checkpoint_shard(paths[1])
    # merge: one unsharded checkpoint under TF names
    engs3 = [DeepFMEngine(cats, nums, max_batch=per, rank=r, world=world, **kw) for r in range(world)]
    for e, p in zip(engs3, paths):
        e.load_checkpoint_shard(p)
    merged = engs3[0].merge_checkpoint_shards(paths, str(tmp_path / "model.ckpt-3.npz"))
    single = DeepFMEngine(cats, nums, max_batch=world * per, **kw)
    single.load_checkpoint(merged)
    assert single.global_step == 3
    feats, y = batches[3]
    z_single = single.predict_logits(feats)
    z_sharded = VirtualCluster(engs3, p2p=True).predict_logits([e.pack(f, None, device=True) for e, (f, _) in zip(engs3, split(feats, y))])
    assert np.array_equal(z_single, np.asarray(z_sharded).reshape(-1))
