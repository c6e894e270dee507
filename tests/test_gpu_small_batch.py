"""GPU: the small-batch step (BASELINE configs[0] / [1]) - one CUDA graph per step (api.cu: train_step_graphed) and the
single-CTA sort + segment builder (embed_kernels.cuh: small_segments_kernel).  Both are shortcuts around launch latency
only: with either one switched off (DFM_NO_GRAPH / DFM_NO_SMALL_SEGMENTS, read once per process, hence the subprocesses)
the same batches must leave bit-identical state behind.  The parity tests against the oracle run through both anyway."""
import os
import subprocess
import sys

import numpy as np
import pytest

from recommender_tensorflow_b200 import synth
from tests.test_gpu_parity import _ml_engine
from tests.util import make_pair

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys
sys.path.insert(0, %r)
import numpy as np
from recommender_tensorflow_b200 import synth
from tests.test_gpu_parity import _ml_engine
from tests.util import make_pair
out = []
# (row, lookup) pairs per step: 26 x batch for k = 4 (52, 832, 2 054, 16 380: single-CTA builder; 16 406, 78 000: radix sort),
# 4 x batch for k = 16 (tiny-vocabulary columns bypass the sort)
for k, hidden, batch in ((4, (16, 16), 32), (4, (16, 16), 2), (4, (16, 16), 79), (4, (16, 16), 630), (4, (16, 16), 631),
                         (4, (16, 16), 3000), (16, (64, 32), 700)):
    eng = _ml_engine(k=k, hidden=hidden, max_batch=4096)
    make_pair(eng, seed=31)
    ml, rng = synth.ML100K(), np.random.default_rng(32)
    losses = []
    for i in range(6):
        f, y = ml.batch(batch, rng)
        if i %% 2:          # host buffers (dfm_train_step_host) and device-resident batches (dfm_train_step) alternate
            losses.append(eng.train_step(f, y))
        else:
            lo = eng.train_step_device(eng.pack(f, y, device=True))      # enqueues only, on the library's stream
            eng.sync()
            losses.append(float(lo.item()))
    out.append("%%s %%s %%d" %% (eng.state_checksum(), np.asarray(losses, np.float32).view(np.uint32).tolist(), eng.graph_steps))
print("RESULT " + " | ".join(out))
""" % ROOT


def _run(env_extra):
    env = dict(os.environ)
    env.pop("DFM_NO_GRAPH", None)
    env.pop("DFM_NO_SMALL_SEGMENTS", None)
    env.update(env_extra)
    r = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    return [part.rsplit(" ", 1) for part in line[len("RESULT "):].split(" | ")]


def test_graph_and_small_segments_change_nothing():
    base = _run({})
    assert all(int(g) == 6 for _, g in base), "every small step replays as one graph: %r" % (base,)
    for env in ({"DFM_NO_GRAPH": "1"}, {"DFM_NO_SMALL_SEGMENTS": "1"}, {"DFM_NO_GRAPH": "1", "DFM_NO_SMALL_SEGMENTS": "1"}):
        other = _run(env)
        assert [s for s, _ in other] == [s for s, _ in base], "state / losses differ with %r" % (env,)
        if "DFM_NO_GRAPH" in env:
            assert all(int(g) == 0 for _, g in other)


def test_graph_path_counts_and_threshold():
    """Every step at or below DFM_GRAPH_MAX_BATCH (default 131 072: every unsharded step) replays as one graph, another
    batch size re-instantiates instead of reusing the graph blindly, and batches above the threshold take ordinary launches."""
    eng = _ml_engine(k=4, hidden=(16, 16), max_batch=16384)
    make_pair(eng, seed=33)
    ml, rng = synth.ML100K(), np.random.default_rng(34)
    eng.train_step(*ml.batch(32, rng))
    eng.train_step_device(eng.pack(*ml.batch(4096, rng), device=True))
    eng.train_step(*ml.batch(16000, rng))
    eng.train_step(*ml.batch(100, rng))
    assert np.isfinite(eng.train_step(*ml.batch(32, rng)))
    if os.environ.get("DFM_NO_GRAPH"):
        assert eng.graph_steps == 0
    elif "DFM_GRAPH_MAX_BATCH" not in os.environ:
        assert eng.graph_steps == 5
    script = SCRIPT.split("out = []")[0] + r"""
eng = _ml_engine(k=4, hidden=(16, 16), max_batch=4096)
make_pair(eng, seed=35)
ml, rng = synth.ML100K(), np.random.default_rng(36)
eng.train_step(*ml.batch(32, rng)); a = eng.graph_steps
eng.train_step(*ml.batch(3000, rng)); b = eng.graph_steps
print("RESULT %d %d" % (a, b))
"""
    env = dict(os.environ)
    env.pop("DFM_NO_GRAPH", None)
    env["DFM_GRAPH_MAX_BATCH"] = "1000"
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1] == "RESULT 1 1"
