"""GPU: the record-staged step (csrc/fused_rows.cu: TMA-staged table records, 3xTF32 layer-0 products on mma.sync, sparse
optimizer of once-only rows applied inside the kernel) against the oracle.  It is selected for embedding_size 16 with a
first hidden layer of 16 (BASELINE.json configs[3]); every case here drives it through the C ABI like any other step."""
import numpy as np
import pytest

from recommender_tensorflow_b200 import synth
from recommender_tensorflow_b200.engine import DeepFMEngine
from tests.test_gpu_parity import _ml_engine, _run_steps
from tests.util import make_pair

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("batch", [2041, 8, 5])
def test_mostly_once_only_rows_with_numerics(batch):
    """Criteo-shaped with tables large enough that nearly every looked-up row is touched once per batch (the in-kernel
    optimizer path), ragged last tile, numeric_embeddings gradients, 5 steps so that rows are replayed too."""
    cats, nums = synth.criteo_columns(50000, n_cat=26, n_num=13)
    eng = DeepFMEngine(cats, nums, embedding_size=16, hidden_units=(16, 16), max_batch=2048)
    ora, _ = make_pair(eng, seed=21)
    rng = np.random.default_rng(22)
    _run_steps(eng, ora, [synth.criteo_batch(batch, rng) for _ in range(5)], "rows-once-%d" % batch)


def test_mixed_once_and_repeated_rows():
    """Small key space: most rows are hit several times (sorted path), some once (in-kernel path), in the same step;
    the same batch is replayed so that rows switch between the two paths from step to step."""
    cats, nums = synth.criteo_columns(8000, n_cat=26, n_num=13)
    eng = DeepFMEngine(cats, nums, embedding_size=16, hidden_units=(16, 16), max_batch=1024)
    ora, _ = make_pair(eng, seed=23)
    rng = np.random.default_rng(24)
    batches = [synth.criteo_batch(1000, rng, key_space=4000) for _ in range(2)]
    batches += [synth.criteo_batch(777, rng, key_space=20000) for _ in range(2)]
    _run_steps(eng, ora, batches, "rows-mixed")


@pytest.mark.parametrize("hidden", [(16,), (16, 32, 8)])
def test_tower_depths(hidden):
    eng = _ml_engine(k=16, hidden=hidden, max_batch=512)
    ora, _ = make_pair(eng, seed=25)
    ml, rng = synth.ML100K(), np.random.default_rng(26)
    _run_steps(eng, ora, [ml.batch(300, rng) for _ in range(4)], "rows-tower%r" % (hidden,))


@pytest.mark.parametrize("name", ["Adagrad", "Ftrl", "RMSProp", "SGD"])
def test_other_optimizers_in_kernel(name):
    from recommender_tensorflow_b200.trainers.model_utils import get_optimizer
    opt = get_optimizer(name, 0.01)
    cats, nums = synth.criteo_columns(20000, n_cat=26, n_num=13)
    eng = DeepFMEngine(cats, nums, embedding_size=16, hidden_units=(16, 16), max_batch=512, opt_deep=opt, opt_linear=dict(opt))
    ora, _ = make_pair(eng, seed=27)
    rng = np.random.default_rng(28)
    _run_steps(eng, ora, [synth.criteo_batch(500, rng) for _ in range(4)], "rows-opt-" + name)


@pytest.mark.parametrize("use", [(1, 0, 1), (0, 1, 1), (0, 0, 1)])
def test_component_subsets_k16(use):
    eng = _ml_engine(k=16, use_linear=bool(use[0]), use_mf=bool(use[1]), use_dnn=bool(use[2]), max_batch=512)
    ora, _ = make_pair(eng, seed=29)
    ml, rng = synth.ML100K(), np.random.default_rng(30)
    _run_steps(eng, ora, [ml.batch(257, rng) for _ in range(3)], "rows-subset%r" % (use,))


def test_bit_identical_reruns_and_unique_rows():
    cats, nums = synth.criteo_columns(50000, n_cat=26, n_num=13)
    rng = np.random.default_rng(31)
    batches = [synth.criteo_batch(1024, rng) for _ in range(3)]
    res = []
    for _ in range(2):
        eng = DeepFMEngine(cats, nums, embedding_size=16, hidden_units=(16, 16), max_batch=1024)
        eng.init_random(32)
        losses = [eng.train_step(f, y) for f, y in batches]
        res.append((losses, eng.state_checksum()))
    assert res[0] == res[1]
