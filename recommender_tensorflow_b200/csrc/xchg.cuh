// Row-sharded step with the exchanges FUSED into the kernels over NVLink peer memory and
// synchronised by flags in that same memory: no collective, no host round trip, no count matrix on the host.
//
// Every rank owns one exchange region (cudaMalloc, mapped into its peers through CUDA IPC):
//     ids  [2][W][CAP]      local row ids requested from me by source s (double buffered by exchange parity)
//     hdr  [2][W]           {count, base of my rows inside s's row buffer, epoch, 0}
//     grads[W][CAP][RW]     gradient rows from source s, entry j belongs to ids[..][s][j]
//     rows [NMAX][RW]       the rows my batch needs, in my unique-row order (written by their owners)
//     dense[W][ND1]         dense gradients + loss of every rank (the tower's all-reduce: each rank adds the W
//                           vectors in rank order, so all replicas stay bit-identical)
//     flags[4][W]           epoch stamps: ids landed / rows landed / gradient rows landed / dense landed, per peer
// CAP = NMAX = max_batch * value slots: every (source, owner) segment has a fixed capacity, so nobody needs to know
// how much the others send before sending.  A producer kernel stores its payload straight into the consumer's
// region and then - last block to finish, after a system-scope fence - stamps the consumer's flag with the epoch
// (st.release.sys); the consumer's stream carries a one-warp wait kernel (ld.acquire.sys spin on its OWN memory,
// bounded by a timeout) in front of the kernel that reads the payload.  One step of one rank:
//     requests -> push ids | wait ids -> serve rows | wait rows -> forward/backward -> gradient rows, dense push
//     | (side stream, after ids: sort the received ids) | wait gradient rows -> sparse apply | wait dense -> dense apply
// Buffer reuse: ids / hdr are double buffered (an owner still sorts epoch e while a fast peer pushes e+1); rows,
// grads and dense are single buffered - their next writer sits behind a flag that the reader only releases later in
// ITS stream (analysis in sharded.py).  Tests run W ranks inside one process on one GPU phase by phase (flags are
// already set when a wait kernel starts: kernels that wait on one another must not share a GPU).
#pragma once
#include "dfm_types.cuh"
#include "embed_kernels.cuh"

constexpr int XF_IDS = 0, XF_ROWS = 1, XF_GRADS = 2, XF_DENSE = 3;

struct XchgDev {
    int32_t  W, me;
    uint32_t cap;                 // entries per (source, owner) segment = rows per row buffer
    int32_t  rw, nd1;             // floats per row payload (K + 4); floats per dense slot
    uint8_t* peer[MAX_PEERS];     // exchange regions (own at index `me`)
    size_t   off_ids[2], off_hdr[2], off_grads, off_rows, off_dense, off_flags;
};
__device__ __forceinline__ uint32_t* x_ids(const XchgDev& x, int r, int par, int s) { return reinterpret_cast<uint32_t*>(x.peer[r] + x.off_ids[par]) + (size_t)s * x.cap; }
__device__ __forceinline__ uint4* x_hdr(const XchgDev& x, int r, int par) { return reinterpret_cast<uint4*>(x.peer[r] + x.off_hdr[par]); }
__device__ __forceinline__ float* x_grads(const XchgDev& x, int r) { return reinterpret_cast<float*>(x.peer[r] + x.off_grads); }
__device__ __forceinline__ float* x_rows(const XchgDev& x, int r) { return reinterpret_cast<float*>(x.peer[r] + x.off_rows); }
__device__ __forceinline__ float* x_dense(const XchgDev& x, int r, int slot) { return reinterpret_cast<float*>(x.peer[r] + x.off_dense) + (size_t)slot * x.nd1; }
__device__ __forceinline__ uint32_t* x_flag(const XchgDev& x, int r, int kind, int from) { return reinterpret_cast<uint32_t*>(x.peer[r] + x.off_flags) + kind * MAX_PEERS + from; }

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// End of a producer kernel: every thread has fenced its stores; the LAST block to arrive stamps flag `kind` of every
// peer (and of this rank itself) with the epoch.  ticket: zero-initialised counter, left at zero.
__device__ __forceinline__ void xchg_signal(const XchgDev& x, int kind, uint32_t epoch, unsigned int* ticket) {
    __shared__ bool x_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) x_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!x_last) return;
    __threadfence_system();
    if ((int)threadIdx.x < x.W) st_release_sys(x_flag(x, threadIdx.x, kind, x.me), epoch);
    if (threadIdx.x == 0) *ticket = 0;
}

// one warp: lane s waits until peer s has stamped flag `kind` with an epoch >= `epoch` (spin on local memory)
__global__ void xchg_wait_kernel(XchgDev x, int kind, uint32_t epoch, int* err) {
    const int s = threadIdx.x;
    if (s >= x.W) return;
    const uint32_t* f = x_flag(x, x.me, kind, s);
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (uint32_t spins = 0;; ++spins) {
        if ((int32_t)(ld_acquire_sys(f) - epoch) >= 0) return;
        if ((spins & 1023u) == 1023u) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 4000000000ull) { atomicOr(err, 2); return; }      // 4 s: a peer died; fail instead of hanging the GPU
        }
        __nanosleep(64);
    }
}

// requester: unique local-row ids (owner-major list req_rows, counts[W] on the device) -> the owners' id segments,
// header {count, reply base, epoch}; also fills the routing table the gradient-row kernel uses (send offsets, where my
// segment of every owner's gradient buffer starts) and the unique-row total.
__global__ void __launch_bounds__(256) xchg_push_kernel(XchgDev x, int par, uint32_t epoch, const uint32_t* __restrict__ req_rows,
                                                        const int32_t* __restrict__ counts, PeerRoute* __restrict__ route,
                                                        unsigned int* ticket, int* err) {
    __shared__ uint32_t soff[MAX_PEERS + 1];
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int o = 0; o < x.W; ++o) { soff[o] = acc; acc += (uint32_t)counts[o]; }
        soff[x.W] = acc;
    }
    __syncthreads();
    const uint32_t U = soff[x.W];
    if (blockIdx.x == 0) {
        if ((int)threadIdx.x < x.W) {
            const int o = threadIdx.x;
            uint32_t c = soff[o + 1] - soff[o];
            if (c > x.cap) { atomicOr(err, 4); c = x.cap; }
            x_hdr(x, o, par)[x.me] = make_uint4(c, soff[o], epoch, 0u);
            route->send_off[o] = soff[o];
            route->dst_off[o] = 0;
            route->peer_grecv[o] = x_grads(x, o) + (size_t)x.me * x.cap * x.rw;
        }
        if (threadIdx.x == 0) { route->W = x.W; route->me = x.me; route->send_off[x.W] = U; }
    }
    for (uint32_t u = blockIdx.x * 256 + threadIdx.x; u < U; u += gridDim.x * 256) {
        int o = 0;
        while (o + 1 < x.W && u >= soff[o + 1]) ++o;
        const uint32_t j = u - soff[o];
        if (j < x.cap) x_ids(x, o, par, x.me)[j] = req_rows[u];
    }
    xchg_signal(x, XF_IDS, epoch, ticket);
}

// owner: for every received id the row as of the previous step (deferred Adam decay replayed in registers, nothing
// written back), stored into the requester's row buffer.  Consecutive entries of one source are contiguous at the
// destination: a warp stages a chunk of rows in shared memory and writes full 512-byte store instructions.
template <int K>
__global__ void __launch_bounds__(256) xchg_serve_kernel(XchgDev x, int par, uint32_t epoch, Table tb, bool has_emb, bool has_lin,
                                                         RowReplay rr, OptDev od, OptDev ol, unsigned int* ticket) {
    constexpr int LPR = K / 4, RPP = 32 / LPR, CH = (1024 / K < 32 ? 1024 / K : 32), PASSES = CH / RPP, RW = K + 4, RW4 = RW / 4;
    __shared__ __align__(16) float tile[8][CH * RW];
    __shared__ uint32_t roff[MAX_PEERS + 1], rbase[MAX_PEERS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane % LPR, grp = lane / LPR;
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int s = 0; s < x.W; ++s) {
            const uint4 hd = x_hdr(x, x.me, par)[s];
            roff[s] = acc; rbase[s] = hd.y;
            acc += hd.x;
        }
        roff[x.W] = acc;
    }
    __syncthreads();
    const ReplayStep rs = replay_step_load(rr.rd.closed ? rr.rd : rr.rl, rr.upto);
    float* tl = tile[warp];
    // chunks never straddle two sources: every source's entries are cut into chunks of CH on their own
    uint32_t cbase[MAX_PEERS + 1];
    {
        uint32_t acc = 0;
        for (int s = 0; s < x.W; ++s) { cbase[s] = acc; acc += (roff[s + 1] - roff[s] + CH - 1) / CH; }
        cbase[x.W] = acc;
    }
    const uint32_t n_chunks = cbase[x.W];
    for (uint32_t chunk = blockIdx.x * 8 + warp; chunk < n_chunks; chunk += gridDim.x * 8) {
        int s = 0;
        while (s + 1 < x.W && chunk >= cbase[s + 1]) ++s;
        const uint32_t j0 = (chunk - cbase[s]) * CH;
        const int cnt = (int)min((uint32_t)CH, roff[s + 1] - roff[s] - j0);
        const uint32_t* ids = x_ids(x, x.me, par, s) + j0;
        float4 e[PASSES];
        float l[PASSES];
#pragma unroll
        for (int p = 0; p < PASSES; ++p) {
            const int r = p * RPP + grp;
            e[p] = make_float4(0.f, 0.f, 0.f, 0.f); l[p] = 0.f;
            if (r < cnt) {
                const size_t row = __ldg(ids + r);
                load_row_current(tb, row, sub, has_emb, rr, rs, od, ol, e[p], l[p]);
                if (!has_lin) l[p] = 0.f;
            }
        }
#pragma unroll
        for (int p = 0; p < PASSES; ++p) {
            const int r = p * RPP + grp;
            reinterpret_cast<float4*>(tl + r * RW)[sub] = e[p];
            if (sub == 0) *reinterpret_cast<float4*>(tl + r * RW + K) = make_float4(l[p], 0.f, 0.f, 0.f);
        }
        __syncwarp();
        float4* dst = reinterpret_cast<float4*>(x_rows(x, s) + (size_t)(rbase[s] + j0) * RW);
        const float4* src4 = reinterpret_cast<const float4*>(tl);
        for (int j = lane; j < cnt * RW4; j += 32) dst[j] = src4[j];
        __syncwarp();
    }
    xchg_signal(x, XF_ROWS, epoch, ticket);
}

// a kernel whose only job is the stamp (after a producer kernel that cannot carry it itself)
__global__ void xchg_signal_kernel(XchgDev x, int kind, uint32_t epoch, unsigned int* ticket) { xchg_signal(x, kind, epoch, ticket); }

// the tower's all-reduce, part 1: my dense gradients (+ my loss share) into slot `me` of every rank
__global__ void __launch_bounds__(256) xchg_dense_push_kernel(XchgDev x, uint32_t epoch, const float* __restrict__ dg, int nd,
                                                              const float* __restrict__ loss, unsigned int* ticket) {
    for (int r = blockIdx.y; r < x.W; r += gridDim.y) {
        float* dst = x_dense(x, r, x.me);
        for (int i = blockIdx.x * 256 + threadIdx.x; i <= nd; i += gridDim.x * 256) dst[i] = i < nd ? dg[i] : *loss;
    }
    // (gridDim.x * gridDim.y blocks take part in the ticket)
    __shared__ bool last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1;
    __syncthreads();
    if (!last) return;
    __threadfence_system();
    if ((int)threadIdx.x < x.W) st_release_sys(x_flag(x, threadIdx.x, XF_DENSE, x.me), epoch);
    if (threadIdx.x == 0) *ticket = 0;
}
// part 2: sum of the W slots in rank order (identical on every rank); element nd is the global loss
__global__ void __launch_bounds__(256) xchg_dense_sum_kernel(XchgDev x, int nd, float* __restrict__ out, float* __restrict__ loss_out) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i > nd) return;
    float s = 0.f;
    for (int r = 0; r < x.W; ++r) s += x_dense(x, x.me, r)[i];
    if (i < nd) out[i] = s;
    else { out[nd] = s; if (loss_out) *loss_out = s; }
}

// owner: (row id, gradient-row index) pairs of everything the peers pushed -> the sort's input; the total stays on the device
__global__ void __launch_bounds__(256) xchg_owner_keys_kernel(XchgDev x, int par, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                              uint32_t cap_out, uint32_t* __restrict__ n_out, int* err) {
    __shared__ uint32_t roff[MAX_PEERS + 1];
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int s = 0; s < x.W; ++s) { roff[s] = acc; acc += x_hdr(x, x.me, par)[s].x; }
        roff[x.W] = acc;
    }
    __syncthreads();
    uint32_t n = roff[x.W];
    if (n > cap_out) { if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(err, 4); n = cap_out; }
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_out = n;
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        int s = 0;
        while (s + 1 < x.W && i >= roff[s + 1]) ++s;
        const uint32_t j = i - roff[s];
        keys[i] = x_ids(x, x.me, par, s)[j];
        vals[i] = (uint32_t)s * x.cap + j;
    }
}
