"""Worker for test_p2p_ipc_two_processes (launched with torchrun, one rank per GPU): runs the same steps through
ShardedTrainer (NCCL all_to_all) and XchgTrainer (kernel stores into IPC-mapped peer memory, flag-synchronised, no
collective inside the step) and requires equal bits; then checks EVAL / PREDICT through both."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recommender_tensorflow_b200 import synth  # noqa: E402
from recommender_tensorflow_b200.engine import DeepFMEngine  # noqa: E402
from recommender_tensorflow_b200.sharded import ShardedTrainer, XchgTrainer  # noqa: E402


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
    ea.init_random(3)
    eb.init_random(3)
    ta, tb = ShardedTrainer(ea), XchgTrainer(eb)
    rng = np.random.default_rng(100 + rank)
    data = [synth.criteo_batch(per, rng, key_space=200000) for _ in range(8)]
    pbs = [eb.pack(f, y, device=True) for f, y in data]
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        for step, (feats, y) in enumerate(data[:6]):
            la = float(ta.train_step(ea.pack(feats, y, device=True), per * world).item())
            lb = float(tb.train_step(pbs[step], per * world).item())
            assert la == lb, (step, la, lb)
        # back-to-back steps without any host synchronisation in between (the flags alone order the ranks)
        for step in (6, 7):
            ta.train_step(ea.pack(*data[step], device=True), per * world)
        # ... the fused exchange with the NEXT batch's requests issued beside each apply phase (dfm_xchg_train_step_next)
        tb.train_step(pbs[6], per * world, next_pb=pbs[7])
        tb.train_step(pbs[7], per * world)
        za = ta.predict_logits(pbs[0]).cpu().numpy()
        zb = tb.predict_logits(pbs[0]).cpu().numpy()
    torch.cuda.synchronize()
    eb.sync()
    assert np.array_equal(za, zb)
    ea.flush()
    eb.flush()
    names = []
    for v in ea.variable_names():
        names.append(v)
        names += [v + "/" + s for s in ea.slot_names(v)]
    for n in names:
        assert np.array_equal(ea.get_tensor(n), eb.get_tensor(n)), n
    dist.barrier()
    if rank == 0:
        print("P2P_OK steps=8 loss=%r" % lb)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
