"""CPU (gloo, world_size 2): the host-side routing of the sharded step.  `route_all_to_all` (used by the
single-process VirtualCluster) must deliver exactly what torch.distributed.all_to_all_single delivers, and
the owner-major request layout must round-trip rows and gradient rows to the right requester slots."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from recommender_tensorflow_b200.sharded import route_all_to_all, split_sizes


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _plan(world, seed):
    """Per rank: unique global rows of its batch, sorted owner-major (owner = g % world, local = g // world)."""
    rng = np.random.default_rng(seed)
    plans = []
    for r in range(world):
        g = np.unique(rng.integers(0, 1000, 200))
        order = np.lexsort((g // world, g % world))
        g = g[order]
        counts = [int((g % world == o).sum()) for o in range(world)]
        plans.append((g, counts))
    return plans


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plans = _plan(world, 7)
    g, counts = plans[rank]
    local = torch.tensor(g // world, dtype=torch.int32)
    cs, cr = torch.tensor(counts, dtype=torch.int32), torch.empty(world, dtype=torch.int32)
    dist.all_to_all_single(cr, cs)
    recv = torch.empty(int(cr.sum()), dtype=torch.int32)
    dist.all_to_all_single(recv, local, cr.tolist(), counts)
    # owner answers with a payload derived from (owner rank, local row); width 3
    reply = torch.stack([recv.float() * world + rank, recv.float(), torch.full_like(recv, rank).float()], 1).reshape(-1)
    rowbuf = torch.empty(len(g) * 3)
    dist.all_to_all_single(rowbuf, reply, split_sizes(counts, 3), split_sizes(cr.tolist(), 3))
    out[rank] = (recv.numpy().copy(), cr.numpy().copy(), rowbuf.numpy().copy())
    dist.destroy_process_group()


def test_routing_matches_gloo_all_to_all():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    plans = _plan(world, 7)
    send = [torch.tensor(g // world, dtype=torch.int32) for g, _ in plans]
    counts = [c for _, c in plans]
    recv, recv_counts = route_all_to_all(send, counts)
    for r in range(world):
        got_recv, got_cr, got_rowbuf = out[r]
        assert (got_recv == recv[r].numpy()).all()
        assert got_cr.tolist() == recv_counts[r]
        # every requested global row came back from its owner, in the requester's own (owner-major) order
        g = plans[r][0]
        rb = got_rowbuf.reshape(-1, 3)
        assert (rb[:, 0] == g).all() and (rb[:, 1] == g // world).all() and (rb[:, 2] == g % world).all()
    # the same reply routed by route_all_to_all
    replies = [torch.stack([recv[r].float() * world + r, recv[r].float(), torch.full_like(recv[r], r).float()], 1).reshape(-1)
               for r in range(world)]
    rowbufs, _ = route_all_to_all(replies, recv_counts, 3)
    for r in range(world):
        assert (rowbufs[r].numpy() == out[r][2]).all()
