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


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_criteo_equals_oracle(world):
    cats, nums = synth.criteo_columns(500, n_cat=26, n_num=13)
    kw = dict(embedding_size=16, hidden_units=(16, 16))
    ref = DeepFMEngine(cats, nums, max_batch=1024, **kw)
    ora, w = make_pair(ref, seed=21)
    engs = [DeepFMEngine(cats, nums, max_batch=1024 // world, rank=r, world=world, **kw) for r in range(world)]
    for e in engs:
        e.set_weights_sharded(w)
    vc = VirtualCluster(engs)
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


def test_sharded_ml100k_hot_rows_world2():
    cols, dtypes = ml100k_columns()
    kw = dict(embedding_size=4, hidden_units=(16, 16), feature_dtypes=dtypes)
    ref = DeepFMEngine(cols, (), max_batch=2048, **kw)
    ora, w = make_pair(ref, seed=23)
    engs = [DeepFMEngine(cols, (), max_batch=1024, rank=r, world=2, **kw) for r in range(2)]
    for e in engs:
        e.set_weights_sharded(w)
    vc = VirtualCluster(engs)
    ml, rng = synth.ML100K(), np.random.default_rng(24)
    for step in range(3):
        feats, y = ml.batch(2048, rng)
        pbs = [e.pack({k: v[r * 1024:(r + 1) * 1024] for k, v in feats.items()}, y[r * 1024:(r + 1) * 1024], device=True)
               for r, e in enumerate(engs)]
        loss, logits = vc.train_step(pbs, return_logits=True)
        rloss, rlogits = ora.train_step_raw(feats, y)
        assert_step_close(loss, logits, ora, rloss, rlogits, 1e-5, "sharded-ml step %d" % step)
    assert_state_close(vc.state(_names(engs[0])), ora.state(), 1e-5, 1e-7, "sharded-ml", ora.state64())


def test_sharded_multi_hot_world2():
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
    vc = VirtualCluster(engs)
    for step in range(3):
        _, feats, y = _bag_setup(512, rng, ml)
        pbs = [e.pack({k: v[r * 256:(r + 1) * 256] for k, v in feats.items()}, y[r * 256:(r + 1) * 256], device=True)
               for r, e in enumerate(engs)]
        loss, logits = vc.train_step(pbs, return_logits=True)
        rloss, rlogits = ora.train_step_raw(feats, y)
        assert_step_close(loss, logits, ora, rloss, rlogits, 1e-5, "sharded-bags step %d" % step)
    assert_state_close(vc.state(_names(engs[0])), ora.state(), 1e-5, 1e-7, "sharded-bags", ora.state64())
