"""Host vs device cost of the small-batch step (BASELINE configs[0] / [1]): wall time of a loop of dfm_train_step calls
without intermediate synchronisation (what bench.py's value measures), the host time of the calls alone, and the device
time between CUDA events.  usage: small_step_time.py [workload] [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from recommender_tensorflow_b200.engine import DeepFMEngine  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "deepfm_ml100k_k4_h16x16_b32"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
w = bench.WORKLOADS[name]
cats, nums, dtypes = bench.make_columns(w, 1)
B = w["batch"]
eng = DeepFMEngine(cats, nums, embedding_size=w["k"], hidden_units=w["hidden"], max_batch=B, device=0, feature_dtypes=dtypes,
                   **bench.optimizers(w))
eng.init_random(1234)
batches = bench.make_batches(w, 8, 777)
packed = [eng.pack(f, y, device=True) for f, y in batches]
loss = torch.zeros(1, device="cuda:0")
for i in range(20):
    eng.train_step_device(packed[i % 8], loss_out=loss)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for i in range(steps):
    eng.train_step_device(packed[i % 8], loss_out=loss)
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("%s: host enqueue %.1f us/step, wall %.1f us/step, device %.1f us/step, graph steps %d, launches/step %d"
      % (name, 1e6 * (t1 - t0) / steps, 1e6 * (t2 - t0) / steps, 1e3 * e0.elapsed_time(e1) / steps, eng.graph_steps, eng.last_step_launches))
