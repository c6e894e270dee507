"""Per-kernel view of the row-sharded step at bench scale on ONE GPU: `world` engines in one process (VirtualCluster, fused
peer-memory exchange wired by raw pointers), the Criteo-shaped workload with the tables of every rank resident side by
side.  Run it under `ncu --metrics gpu__time_duration.sum` for the launch list of the four phases (multi-rank commands
must not run under ncu).  usage: shard_profile.py [world] [steps]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench  # noqa: E402
from recommender_tensorflow_b200 import synth  # noqa: E402
from recommender_tensorflow_b200.engine import DeepFMEngine  # noqa: E402
from recommender_tensorflow_b200.sharded import VirtualCluster  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
name = "deepfm_criteo_1e7_k16_h16x16_b65536"
w = bench.WORKLOADS[name]
cats, nums, dtypes = bench.make_columns(w, world)
B = w["batch"]
kw = dict(embedding_size=w["k"], hidden_units=w["hidden"], max_batch=B, device=0, feature_dtypes=dtypes, **bench.optimizers(w))
engs = []
for r in range(world):
    e = DeepFMEngine(cats, nums, rank=r, world=world, **kw)
    e.init_random(1234)
    engs.append(e)
vc = VirtualCluster(engs, p2p=True)
batches = [synth.criteo_device_batches(e, B, steps + 2, 777 + r) for r, e in enumerate(engs)]
for s in range(steps + 2):
    if s == 2:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    loss = vc.train_step([b[s] for b in batches])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", loss)
