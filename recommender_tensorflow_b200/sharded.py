"""Row-sharded DeepFM step over the GPUs of one box.

Tables are row-sharded (global row g lives on rank g % world at local index g / world), the dense
tower is replicated and the batch is data-parallel.  Per step every rank runs the four C-ABI phases
(`dfm_shard_requests / serve / forward_backward / apply`); between them this module performs the
collectives with `torch.distributed` (NCCL over NVLink on the GPUs, gloo in the CPU tests of the
routing logic):

    all_to_all(counts) -> all_to_all(row ids) -> [serve] -> all_to_all(rows) -> [forward/backward]
    -> all_to_all(gradient rows) + all_reduce(dense grads, loss) -> [apply]

`VirtualCluster` runs the same phases for `world` engines inside ONE process on one GPU, routing the
buffers with plain tensor copies; it is how the sharded path is checked against the single-GPU
result without a multi-GPU box (B200_PROFILING.md: emulate ranks in one process, never as
co-running kernels that wait on each other).
"""
import numpy as np


def split_sizes(counts, width=1):
    return [int(c) * width for c in counts]


def route_all_to_all(send_bufs, send_counts, width=1):
    """Pure routing used by VirtualCluster and by the CPU tests: send_bufs[r] is rank r's send buffer laid
    out by destination (send_counts[r][dst] items of `width` elements each).  Returns (recv_bufs, recv_counts)
    where recv_bufs[dst] is the concatenation over source ranks, in rank order."""
    import torch
    W = len(send_bufs)
    recv_counts = [[int(send_counts[src][dst]) for src in range(W)] for dst in range(W)]
    offs = [np.concatenate([[0], np.cumsum(send_counts[r])]).astype(np.int64) for r in range(W)]
    recv = []
    for dst in range(W):
        parts = [send_bufs[src][int(offs[src][dst]) * width:int(offs[src][dst + 1]) * width] for src in range(W)]
        recv.append(torch.cat(parts) if parts else send_bufs[0][:0])
    return recv, recv_counts


def _on_real_stream(step):
    """The C ABI treats a NULL stream as "the handle's own stream", so a step issued while torch's current stream is
    the legacy default stream (handle 0) would run the kernels and the collectives on two unordered streams.  Run
    such a step on a private stream, ordered after / before the default stream."""
    def wrapped(self, *a, **kw):
        torch = self.torch
        cur = torch.cuda.current_stream()
        if cur.cuda_stream != 0:
            return step(self, *a, **kw)
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream()
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            out = step(self, *a, **kw)
        cur.wait_stream(self._side)
        return out
    wrapped.__doc__ = step.__doc__
    return wrapped


class ShardedTrainer:
    """One process per GPU (torchrun); `engine` was created with rank / world of the process group."""

    def __init__(self, engine, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.eng, self.group = engine, group
        self.W, self.rank = engine.world, engine.rank
        dev = "cuda:%d" % engine.device
        n = engine.max_batch * max(engine.n_slots, 1)
        rw = engine.row_width
        self.req_rows = torch.empty(n, dtype=torch.int32, device=dev)
        self.recv_rows = torch.empty(2 * n + 4096, dtype=torch.int32, device=dev)
        self.reply = torch.empty((2 * n + 4096) * rw, dtype=torch.float32, device=dev)
        self.rowbuf = torch.empty(n * rw, dtype=torch.float32, device=dev)
        self.gsum = torch.empty(n * rw, dtype=torch.float32, device=dev)
        self.grecv = torch.empty((2 * n + 4096) * rw, dtype=torch.float32, device=dev)
        self.dense = torch.zeros(max(engine.dense_size, 1) + 1, dtype=torch.float32, device=dev)   # [+1]: loss rides along
        self.cnt_send = torch.empty(self.W, dtype=torch.int32, device=dev)
        self.cnt_recv = torch.empty(self.W, dtype=torch.int32, device=dev)
        self.rw = rw

    @_on_real_stream
    def train_step(self, pb, global_batch, timings=None):
        """One sharded train step on a device-resident PackedBatch; everything is enqueued on torch's current
        stream (the only host syncs are the two count exchanges).  Returns the global loss as a 0-dim cuda
        tensor.  `timings` (dict) collects per-phase device milliseconds when given (debug)."""
        torch, dist, eng, rw = self.torch, self.dist, self.eng, self.rw
        st = torch.cuda.current_stream().cuda_stream
        marks = []

        def mark(name):
            if timings is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append((name, e))
        mark("start")
        send_counts = eng.shard_requests(pb, self.req_rows, st)
        mark("requests")
        self.cnt_send.copy_(torch.tensor(send_counts, dtype=torch.int32), non_blocking=True)
        dist.all_to_all_single(self.cnt_recv, self.cnt_send, group=self.group)
        recv_counts = self.cnt_recv.tolist()
        mark("counts")
        U, n_recv = sum(send_counts), sum(recv_counts)
        dist.all_to_all_single(self.recv_rows[:n_recv], self.req_rows[:U], recv_counts, send_counts, group=self.group)
        mark("a2a_ids")
        eng.shard_serve(self.recv_rows, n_recv, self.reply, st)
        mark("serve")
        dist.all_to_all_single(self.rowbuf[:U * rw], self.reply[:n_recv * rw], split_sizes(send_counts, rw),
                               split_sizes(recv_counts, rw), group=self.group)
        mark("a2a_rows")
        nd = eng.dense_size
        eng.shard_forward_backward(pb, self.rowbuf, global_batch, self.dense[nd:nd + 1], None, self.gsum, self.dense, st)
        mark("fwd_bwd")
        dist.all_to_all_single(self.grecv[:n_recv * rw], self.gsum[:U * rw], split_sizes(recv_counts, rw),
                               split_sizes(send_counts, rw), group=self.group)
        mark("a2a_grads")
        dist.all_reduce(self.dense, group=self.group)
        mark("allreduce")
        eng.shard_apply(self.grecv, self.dense, st)
        mark("apply")
        if timings is not None:
            torch.cuda.synchronize()
            for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
                timings[n1] = timings.get(n1, 0.0) + e0.elapsed_time(e1)
        return self.dense[nd]


    @_on_real_stream
    def predict_logits(self, pb):
        """EVAL / PREDICT: logits of this rank's batch (cuda tensor [B]); no state changes."""
        torch, dist, eng, rw = self.torch, self.dist, self.eng, self.rw
        st = torch.cuda.current_stream().cuda_stream
        send_counts = eng.shard_requests(pb, self.req_rows, st)
        self.cnt_send.copy_(torch.tensor(send_counts, dtype=torch.int32), non_blocking=True)
        dist.all_to_all_single(self.cnt_recv, self.cnt_send, group=self.group)
        recv_counts = self.cnt_recv.tolist()
        U, n_recv = sum(send_counts), sum(recv_counts)
        dist.all_to_all_single(self.recv_rows[:n_recv], self.req_rows[:U], recv_counts, send_counts, group=self.group)
        eng.shard_serve(self.recv_rows, n_recv, self.reply, st)
        dist.all_to_all_single(self.rowbuf[:U * rw], self.reply[:n_recv * rw], split_sizes(send_counts, rw),
                               split_sizes(recv_counts, rw), group=self.group)
        logits = torch.empty(pb.batch_size, dtype=torch.float32, device=self.rowbuf.device)
        eng.shard_forward(pb, self.rowbuf, logits, st)
        return logits


class P2PShardedTrainer:
    """Same step with the three payload exchanges fused into the kernels: every rank's receive buffers are
    mapped into its peers through CUDA IPC, and the requesting / serving / gradient kernels store straight
    into the destination GPU over NVLink.  NCCL carries only the W x W count matrix, two 4-byte barriers and
    the dense-gradient all_reduce (which doubles as the barrier before `apply`).

    Ordering (every call below is stream-ordered; "bar" = a collective every rank must enter):
        push_ids | bar1 | serve | bar2 | forward_backward | all_reduce(dense) | apply
    * within a step: a rank reads its id buffer only after bar1, its row buffer only after bar2 and its gradient
      buffer only after the dense all_reduce, i.e. after every peer has finished the kernel that stores into it
      (a peer enters the collective only after that kernel, and kernel completion makes its peer stores visible);
    * across steps: a peer's push_ids(t+1) follows its all_reduce(t), which completes only after I entered it, i.e.
      after my serve(t) read the ids; its serve(t+1) follows bar1(t+1), which I enter after my forward_backward(t)
      read the rows; its forward_backward(t+1) follows bar2(t+1), which I enter after my apply(t) read the
      gradient rows.  So no buffer is overwritten before its reader is done and no double buffering is needed."""

    def __init__(self, engine, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.eng, self.group = engine, group
        self.W, self.rank = engine.world, engine.rank
        dev = "cuda:%d" % engine.device
        mine = torch.frombuffer(bytearray(engine.shard_ipc_export()), dtype=torch.uint8).to(dev)
        allh = torch.empty(self.W * 192, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh, mine, group=group)
        engine.shard_ipc_import(bytes(allh.cpu().numpy().tobytes()))
        self.dense = torch.zeros(max(engine.dense_size, 1) + 1, dtype=torch.float32, device=dev)
        self.cnt_mine = torch.empty(self.W, dtype=torch.int32, device=dev)
        self.cnt_next = torch.empty(self.W, dtype=torch.int32, device=dev)
        self.cnt_all = torch.empty(self.W * self.W, dtype=torch.int32, device=dev)
        self.flag = torch.zeros(1, dtype=torch.float32, device=dev)
        self._prefetched = None        # the PackedBatch whose requests were computed ahead of time
        dist.barrier(group=group)

    @_on_real_stream
    def train_step(self, pb, global_batch, timings=None, next_pb=None):
        """next_pb: the batch of the following step, if known — its requests (transform, sort, unique rows: no model
        state) are then computed on the library's side stream while this step runs."""
        torch, dist, eng = self.torch, self.dist, self.eng
        st = torch.cuda.current_stream().cuda_stream
        marks = []

        def mark(name):
            if timings is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append((name, e))
        mark("start")
        if self._prefetched is pb:
            eng.shard_adopt_prefetch(pb, st)                       # orders this stream after the prefetch
            self.cnt_mine, self.cnt_next = self.cnt_next, self.cnt_mine
        else:
            eng.shard_requests_dev(pb, self.cnt_mine, st)
        self._prefetched = None
        mark("requests")
        dist.all_gather_into_tensor(self.cnt_all, self.cnt_mine, group=self.group)
        eng.shard_p2p_plan(self.cnt_all.cpu().numpy().reshape(self.W, self.W), st)
        mark("counts")
        if next_pb is not None:
            eng.shard_prefetch_requests(next_pb, self.cnt_next, st)
            self._prefetched = next_pb
        eng.shard_p2p_push_ids(st)
        dist.all_reduce(self.flag, group=self.group)          # barrier: every owner has all its requests
        mark("push_ids")
        eng.shard_p2p_serve(st)
        dist.all_reduce(self.flag, group=self.group)          # barrier: every requester has all its rows
        mark("serve")
        nd = eng.dense_size
        eng.shard_p2p_forward_backward(pb, global_batch, self.dense[nd:nd + 1], None, self.dense, st)
        mark("fwd_bwd")
        dist.all_reduce(self.dense, group=self.group)         # dense grads + loss; also the barrier before apply
        mark("allreduce")
        eng.shard_p2p_apply(self.dense, st)
        mark("apply")
        if timings is not None:
            torch.cuda.synchronize()
            for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
                timings[n1] = timings.get(n1, 0.0) + e0.elapsed_time(e1)
        return self.dense[nd]


    @_on_real_stream
    def predict_logits(self, pb):
        """EVAL / PREDICT over the fused exchange: logits of this rank's batch (cuda tensor [B]); no state changes."""
        torch, dist, eng = self.torch, self.dist, self.eng
        st = torch.cuda.current_stream().cuda_stream
        eng.shard_requests_dev(pb, self.cnt_mine, st)
        dist.all_gather_into_tensor(self.cnt_all, self.cnt_mine, group=self.group)
        eng.shard_p2p_plan(self.cnt_all.cpu().numpy().reshape(self.W, self.W), st)
        eng.shard_p2p_push_ids(st)
        dist.all_reduce(self.flag, group=self.group)
        eng.shard_p2p_serve(st)
        dist.all_reduce(self.flag, group=self.group)
        logits = torch.empty(pb.batch_size, dtype=torch.float32, device=self.flag.device)
        eng.shard_forward(pb, None, logits, st)
        dist.all_reduce(self.flag, group=self.group)       # nobody overwrites a row buffer that is still being read
        return logits


class VirtualCluster:
    """`world` engines in one process on one GPU; same phases, routing by tensor copies (or, with p2p=True,
    by the fused peer-memory stores — the handles are wired with raw pointers since they share a process)."""

    def __init__(self, engines, p2p=False):
        import torch
        self.torch = torch
        self.engs = engines
        self.W = len(engines)
        self.p2p = p2p
        if p2p:
            ptrs = []
            for e in engines:
                ptrs += e.shard_p2p_buffers()
            for e in engines:
                e.shard_p2p_set_peers(ptrs)

    def _train_step_p2p(self, pbs, return_logits, next_pbs=None):
        torch, W = self.torch, self.W
        dev = "cuda:%d" % self.engs[0].device
        global_batch = sum(pb.batch_size for pb in pbs)
        nd = self.engs[0].dense_size
        if getattr(self, "_prefetched", None) is not None and all(a is b for a, b in zip(self._prefetched, pbs)):
            for e, pb in zip(self.engs, pbs):
                e.shard_adopt_prefetch(pb)
                e.sync()
            counts = self._cnt_next.cpu().numpy().astype(np.int32)
        else:
            counts = np.array([e.shard_requests_counts(pb) for e, pb in zip(self.engs, pbs)], dtype=np.int32)
        self._prefetched = None
        for e in self.engs:
            e.shard_p2p_plan(counts)
        if next_pbs is not None:       # requests of the next batch, on the engines' side streams, while this step runs
            self._cnt_next = torch.empty((W, W), dtype=torch.int32, device=dev)
            for r, (e, pb) in enumerate(zip(self.engs, next_pbs)):
                e.shard_prefetch_requests(pb, self._cnt_next[r])
            self._prefetched = list(next_pbs)
        for e in self.engs:
            e.shard_p2p_push_ids()
        for e in self.engs:
            e.sync()
        for e in self.engs:
            e.shard_p2p_serve()
        for e in self.engs:
            e.sync()
        denses, logits = [], []
        for e, pb in zip(self.engs, pbs):
            dn = torch.zeros(nd + 1, dtype=torch.float32, device=dev)
            lg = torch.empty(pb.batch_size, dtype=torch.float32, device=dev)
            e.shard_p2p_forward_backward(pb, global_batch, dn[nd:nd + 1], lg, dn)
            denses.append(dn); logits.append(lg)
        for e in self.engs:
            e.sync()
        total = torch.stack(denses).sum(0)
        for e in self.engs:
            e.shard_p2p_apply(total)
            e.sync()
        loss = float(total[nd].item())
        if return_logits:
            return loss, torch.cat(logits).cpu().numpy()
        return loss

    def train_step(self, pbs, return_logits=False, next_pbs=None):
        if self.p2p:
            return self._train_step_p2p(pbs, return_logits, next_pbs)
        torch, W = self.torch, self.W
        dev = "cuda:%d" % self.engs[0].device
        rw = self.engs[0].row_width
        global_batch = sum(pb.batch_size for pb in pbs)
        req, counts = [], []
        for e, pb in zip(self.engs, pbs):
            buf = torch.empty(pb.batch_size * max(e.n_slots, 1), dtype=torch.int32, device=dev)
            counts.append(e.shard_requests(pb, buf))
            req.append(buf[:sum(counts[-1])])
        recv_rows, recv_counts = route_all_to_all(req, counts)
        replies = []
        for e, rr in zip(self.engs, recv_rows):
            rep = torch.empty(max(rr.numel(), 1) * rw, dtype=torch.float32, device=dev)
            e.shard_serve(rr.contiguous(), rr.numel(), rep)
            e.sync()
            replies.append(rep[:rr.numel() * rw])
        rowbufs, _ = route_all_to_all(replies, recv_counts, rw)
        gsums, denses, logits = [], [], []
        nd = self.engs[0].dense_size
        for e, pb, rb, c in zip(self.engs, pbs, rowbufs, counts):
            U = sum(c)
            gs = torch.empty(max(U, 1) * rw, dtype=torch.float32, device=dev)
            dn = torch.zeros(nd + 1, dtype=torch.float32, device=dev)
            lg = torch.empty(pb.batch_size, dtype=torch.float32, device=dev)
            rbc = rb.contiguous() if rb.numel() else torch.zeros(rw, dtype=torch.float32, device=dev)
            e.shard_forward_backward(pb, rbc, global_batch, dn[nd:nd + 1], lg, gs, dn)
            e.sync()
            gsums.append(gs[:U * rw]); denses.append(dn); logits.append(lg)
        grecv, _ = route_all_to_all(gsums, counts, rw)
        total = torch.stack(denses).sum(0)            # all_reduce (rank order)
        for e, gr in zip(self.engs, grecv):
            grc = gr.contiguous() if gr.numel() else torch.zeros(rw, dtype=torch.float32, device=dev)
            e.shard_apply(grc, total)
            e.sync()
        loss = float(total[nd].item())
        if return_logits:
            return loss, torch.cat(logits).cpu().numpy()
        return loss

    def predict_logits(self, pbs):
        """EVAL / PREDICT: logits of all ranks' batches, concatenated in rank order (numpy)."""
        torch = self.torch
        dev = "cuda:%d" % self.engs[0].device
        rw = self.engs[0].row_width
        outs = []
        if self.p2p:
            counts = np.array([e.shard_requests_counts(pb) for e, pb in zip(self.engs, pbs)], dtype=np.int32)
            for e in self.engs:
                e.shard_p2p_plan(counts)
            for e in self.engs:
                e.shard_p2p_push_ids()
            for e in self.engs:
                e.sync()
            for e in self.engs:
                e.shard_p2p_serve()
            for e in self.engs:
                e.sync()
            for e, pb in zip(self.engs, pbs):
                lg = torch.empty(pb.batch_size, dtype=torch.float32, device=dev)
                e.shard_forward(pb, None, lg)
                e.sync()
                outs.append(lg)
        else:
            req, counts = [], []
            for e, pb in zip(self.engs, pbs):
                buf = torch.empty(pb.batch_size * max(e.n_slots, 1), dtype=torch.int32, device=dev)
                counts.append(e.shard_requests(pb, buf))
                req.append(buf[:sum(counts[-1])])
            recv_rows, recv_counts = route_all_to_all(req, counts)
            replies = []
            for e, rr in zip(self.engs, recv_rows):
                rep = torch.empty(max(rr.numel(), 1) * rw, dtype=torch.float32, device=dev)
                e.shard_serve(rr.contiguous(), rr.numel(), rep)
                e.sync()
                replies.append(rep[:rr.numel() * rw])
            rowbufs, _ = route_all_to_all(replies, recv_counts, rw)
            for e, pb, rb in zip(self.engs, pbs, rowbufs):
                lg = torch.empty(pb.batch_size, dtype=torch.float32, device=dev)
                rbc = rb.contiguous() if rb.numel() else torch.zeros(rw, dtype=torch.float32, device=dev)
                e.shard_forward(pb, rbc, lg)
                e.sync()
                outs.append(lg)
        return torch.cat(outs).cpu().numpy()

    def state(self, names_with_slots):
        """Global arrays reassembled from the shards (dense variables from rank 0)."""
        out = {}
        for name in names_with_slots:
            base = name.split("/")[0]
            if base in ("emb", "lin"):
                parts = [e.get_tensor(name) for e in self.engs]
                R = sum(p.shape[0] for p in parts)
                full = np.empty((R,) + parts[0].shape[1:], dtype=np.float32)
                for r, p in enumerate(parts):
                    full[r::self.W] = p
                out[name] = full
            else:
                out[name] = self.engs[0].get_tensor(name)
        return out
