"""Worker for test_p2p_ipc_two_processes (launched with torchrun, one rank per GPU): runs the same steps through
ShardedTrainer (NCCL all_to_all) and P2PShardedTrainer (stores over IPC-mapped peer memory) and requires equal bits."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recommender_tensorflow_b200 import synth  # noqa: E402
from recommender_tensorflow_b200.engine import DeepFMEngine  # noqa: E402
from recommender_tensorflow_b200.sharded import P2PShardedTrainer, ShardedTrainer  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cats, nums = synth.criteo_columns(50000, n_cat=26, n_num=13)
    kw = dict(embedding_size=16, hidden_units=(16, 16), device=local, rank=rank, world=world)
    per = 4096
    ea = DeepFMEngine(cats, nums, max_batch=per, **kw)
    eb = DeepFMEngine(cats, nums, max_batch=per, **kw)
    ec = DeepFMEngine(cats, nums, max_batch=per, **kw)
    ea.init_random(3)
    eb.init_random(3)
    ec.init_random(3)
    ta, tb, tc = ShardedTrainer(ea), P2PShardedTrainer(eb), P2PShardedTrainer(ec)
    rng = np.random.default_rng(100 + rank)
    data = [synth.criteo_batch(per, rng, key_space=200000) for _ in range(6)]
    pcs = [ec.pack(f, y, device=True) for f, y in data]
    for step, (feats, y) in enumerate(data):
        la = float(ta.train_step(ea.pack(feats, y, device=True), per * world).item())
        lb = float(tb.train_step(eb.pack(feats, y, device=True), per * world).item())
        lc = float(tc.train_step(pcs[step], per * world, next_pb=pcs[step + 1] if step + 1 < len(pcs) else None).item())   # with request prefetch
        assert la == lb == lc, (step, la, lb, lc)
    ea.flush()
    eb.flush()
    ec.flush()
    names = []
    for v in ea.variable_names():
        names.append(v)
        names += [v + "/" + s for s in ea.slot_names(v)]
    for n in names:
        assert np.array_equal(ea.get_tensor(n), eb.get_tensor(n)), n
        assert np.array_equal(ea.get_tensor(n), ec.get_tensor(n)), n
    dist.barrier()
    if rank == 0:
        print("P2P_OK steps=6 loss=%r" % lb)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
