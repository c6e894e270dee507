#include "prims.cuh"

namespace prims {

// ------------------------------------------------------------------------------------------ scan
template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* warp_tot /*[32]*/, T& block_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    T inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        T w = lane < nwarp ? warp_tot[lane] : T(0);
        T winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            T t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        warp_tot[lane] = winc - w;            // exclusive warp offsets
        if (lane == 31) warp_tot[32] = winc;  // total
    }
    __syncthreads();
    T res = warp_tot[warp] + inc - v;
    block_total = warp_tot[32];
    __syncthreads();
    return res;
}

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const T* __restrict__ in, int64_t n,
                                                                   T* __restrict__ block_sums) {
    __shared__ T wt[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    T s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j)
        if (base + j < n) s += in[base + j];
    T tot;
    block_exclusive_scan<T>(s, wt, tot);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

template <typename T>
__global__ void __launch_bounds__(1024) scan_spine_kernel(T* __restrict__ block_sums, int64_t nb, T* __restrict__ total) {
    __shared__ T wt[33];
    T carry = 0;
    for (int64_t base = 0; base < nb; base += blockDim.x) {
        int64_t i = base + threadIdx.x;
        T v = i < nb ? block_sums[i] : T(0);
        T tot;
        T ex = block_exclusive_scan<T>(v, wt, tot);
        if (i < nb) block_sums[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) {
        block_sums[nb] = carry;
        if (total) *total = carry;
    }
}

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_down_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t n,
                                                                 const T* __restrict__ block_sums) {
    __shared__ T wt[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    T v[SCAN_ITEMS];
    T s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        v[j] = (base + j < n) ? in[base + j] : T(0);
        s += v[j];
    }
    T tot;
    T ex = block_exclusive_scan<T>(s, wt, tot) + block_sums[blockIdx.x];
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        if (base + j < n) out[base + j] = ex;
        ex += v[j];
    }
}

// small inputs (radix-digit counts of a small sort, segment flags of a small batch): one block walks the tiles with a
// running carry -> one launch instead of three dependent ones
template <typename T>
__global__ void __launch_bounds__(1024) scan_single_block_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t n, T* __restrict__ total) {
    __shared__ T wt[33];
    constexpr int ITEMS = 4;
    T carry = 0;
    for (int64_t base = 0; base < n; base += 1024 * ITEMS) {
        const int64_t i0 = base + (int64_t)threadIdx.x * ITEMS;
        T v[ITEMS];
        T s = 0;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) { v[j] = (i0 + j < n) ? in[i0 + j] : T(0); s += v[j]; }
        T tot;
        T ex = block_exclusive_scan<T>(s, wt, tot) + carry;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) { if (i0 + j < n) out[i0 + j] = ex; ex += v[j]; }
        carry += tot;
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

constexpr int64_t SCAN_SINGLE_MAX = 4096;      // one tile; measured: at 32 K elements three launches beat one block (0.054 vs 0.066 ms per sort)

template <typename T>
static void exclusive_scan_t(const T* in, T* out, int64_t n, void* temp, T* total, cudaStream_t st, int64_t* launches) {
    if (n <= 0) {
        if (total) cudaMemsetAsync(total, 0, sizeof(T), st);
        return;
    }
    if (n <= SCAN_SINGLE_MAX) {
        scan_single_block_kernel<T><<<1, 1024, 0, st>>>(in, out, n, total);
        if (launches) *launches += 1;
        return;
    }
    T* bs = reinterpret_cast<T*>(temp);
    const int64_t nb = scan_blocks(n);
    scan_reduce_kernel<T><<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, n, bs);
    scan_spine_kernel<T><<<1, 1024, 0, st>>>(bs, nb, total);
    scan_down_kernel<T><<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, out, n, bs);
    if (launches) *launches += 3;
}

void exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, void* temp, uint32_t* total, cudaStream_t st, int64_t* l) {
    exclusive_scan_t<uint32_t>(in, out, n, temp, total, st, l);
}
void exclusive_scan_u64(const uint64_t* in, uint64_t* out, int64_t n, void* temp, uint64_t* total, cudaStream_t st, int64_t* l) {
    exclusive_scan_t<unsigned long long>(reinterpret_cast<const unsigned long long*>(in),
                                         reinterpret_cast<unsigned long long*>(out), n, temp,
                                         reinterpret_cast<unsigned long long*>(total), st, l);
}

// ------------------------------------------------------------------------------------ radix sort
// Tile layout shared by the histogram and scatter kernels: warp w of a block owns the contiguous
// index range [tile + w*32*ITEMS, +32*ITEMS); item j of lane l is index base + j*32 + l, so the
// stable order inside a warp is (j, lane) lexicographic and warps / blocks follow index order.
template <int RB>
__global__ void __launch_bounds__(SORT_THREADS) radix_hist_kernel(const uint32_t* __restrict__ keys, int64_t n, int shift,
                                                                  uint32_t* __restrict__ counts, int64_t nblocks) {
    constexpr int RADIX = 1 << RB;
    __shared__ uint32_t hist[RADIX];
    for (int i = threadIdx.x; i < RADIX; i += SORT_THREADS) hist[i] = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t base = (int64_t)blockIdx.x * SORT_TILE + (int64_t)warp * 32 * SORT_ITEMS + lane;
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; ++j) {
        int64_t idx = base + j * 32;
        if (idx < n) atomicAdd(&hist[(keys[idx] >> shift) & (RADIX - 1)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < RADIX; i += SORT_THREADS) counts[(int64_t)i * nblocks + blockIdx.x] = hist[i];
}

template <int RB>
__global__ void __launch_bounds__(SORT_THREADS) radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                                     int64_t n, int shift, const uint32_t* __restrict__ offsets,
                                                                     int64_t nblocks) {
    constexpr int RADIX = 1 << RB;
    constexpr int NW = SORT_THREADS / 32;
    __shared__ uint32_t wcnt[NW][RADIX];
    for (int i = threadIdx.x; i < NW * RADIX; i += SORT_THREADS) (&wcnt[0][0])[i] = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int64_t base = (int64_t)blockIdx.x * SORT_TILE + (int64_t)warp * 32 * SORT_ITEMS + lane;
    uint32_t k[SORT_ITEMS], v[SORT_ITEMS], rank[SORT_ITEMS];
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; ++j) {
        int64_t idx = base + j * 32;
        bool valid = idx < n;
        k[j] = valid ? keys_in[idx] : 0u;
        v[j] = valid ? vals_in[idx] : 0u;
    }
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; ++j) {
        int64_t idx = base + j * 32;
        bool valid = idx < n;
        uint32_t d = valid ? ((k[j] >> shift) & (RADIX - 1)) : (uint32_t)RADIX;  // tail items form their own group
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        uint32_t r = __popc(peers & lt_mask);
        uint32_t old = 0;
        if (valid && r == 0) {
            old = wcnt[warp][d];
            wcnt[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1);
        rank[j] = old + r;
        __syncwarp();
    }
    __syncthreads();
    // exclusive prefix over the warps of this block, seeded with the global (digit, block) offset
    for (int d = threadIdx.x; d < RADIX; d += SORT_THREADS) {
        uint32_t run = offsets[(int64_t)d * nblocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            uint32_t c = wcnt[w][d];
            wcnt[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; ++j) {
        int64_t idx = base + j * 32;
        if (idx < n) {
            uint32_t d = (k[j] >> shift) & (RADIX - 1);
            uint32_t pos = wcnt[warp][d] + rank[j];
            keys_out[pos] = k[j];
            vals_out[pos] = v[j];
        }
    }
}

template <int RB>
static int radix_sort_rb(uint32_t* keys[2], uint32_t* vals[2], int64_t n, int bits, void* temp, cudaStream_t st, int64_t* launches) {
    const int64_t nb = sort_blocks(n);
    const int64_t cnt = (int64_t)(1 << RB) * nb;
    uint32_t* counts = reinterpret_cast<uint32_t*>(temp);
    uint32_t* offsets = counts + (int64_t)MAX_RADIX * nb;
    void* scan_temp = reinterpret_cast<void*>(offsets + (int64_t)MAX_RADIX * nb);
    int cur = 0;
    for (int shift = 0; shift < bits; shift += RB) {
        radix_hist_kernel<RB><<<(unsigned)nb, SORT_THREADS, 0, st>>>(keys[cur], n, shift, counts, nb);
        if (launches) *launches += 1;
        exclusive_scan_u32(counts, offsets, cnt, scan_temp, nullptr, st, launches);
        radix_scatter_kernel<RB><<<(unsigned)nb, SORT_THREADS, 0, st>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n, shift,
                                                                        offsets, nb);
        if (launches) *launches += 1;
        cur ^= 1;
    }
    return cur;
}

int radix_sort_pairs(uint32_t* keys[2], uint32_t* vals[2], int64_t n, int bits, void* temp, cudaStream_t st, int64_t* launches) {
    if (n <= 1 || bits <= 0) return 0;
    // fewest passes first, then the narrowest digit (smaller histograms)
    const int p8 = (bits + 7) / 8, p9 = (bits + 8) / 9, p10 = (bits + 9) / 10;
    if (p8 <= p9 && p8 <= p10) return radix_sort_rb<8>(keys, vals, n, bits, temp, st, launches);
    if (p9 <= p10) return radix_sort_rb<9>(keys, vals, n, bits, temp, st, launches);
    return radix_sort_rb<10>(keys, vals, n, bits, temp, st, launches);
}

}  // namespace prims

// =====================================================================================================================
// One-sweep LSD radix sort (8-bit digits) with decoupled look-back, and a single-pass segment builder.
//
// (A variant without look-back - per digit one count launch whose last block scans the tile x digit matrix, then one
//  scatter launch - measured slower: 0.22 ms against 0.16 ms for this one, 4 digits of 1.7 M pairs, profiles/r02n.)
// The multi-launch sort above costs (histogram + 3 scan launches + scatter) per digit = 15-20 dependent launches of
// small grids, 0.20 ms for the 1.7 M (row, lookup) pairs of a Criteo-shaped batch although it moves only ~110 MB.
// Here ONE kernel builds the digit histograms of all passes, and every pass is ONE kernel: a tile (4096 pairs) ranks
// its keys, publishes its per-digit counts, learns the counts of all earlier tiles by looking back along a chain of
// status words (tile i publishes "aggregate" first, then "inclusive prefix"; a reader walks back until it meets a
// prefix) and scatters.  Tiles are handed out by an atomic ticket, so a tile only ever waits for tiles that started
// before it (forward progress without a co-residency assumption).  The element count may live in DEVICE memory
// (n_dev): the row-sharded owner side sorts what its peers pushed without the host ever learning how much arrived.
// =====================================================================================================================
namespace prims {

constexpr int OS_THREADS = 256;
constexpr int OS_ITEMS = 16;
constexpr int OS_TILE = OS_THREADS * OS_ITEMS;     // 4096 pairs per tile
constexpr int OS_RADIX = 256;
constexpr int OS_MAX_PASSES = 4;
constexpr uint32_t OS_AGG = 1u << 30, OS_PREFIX = 2u << 30, OS_VALUE = (1u << 30) - 1u;

static inline int64_t os_tiles(int64_t n) { return (n + OS_TILE - 1) / OS_TILE; }

// temp layout: [hist 4 x 256][tickets 8][status 4 x tiles_max x 256]   (uint32)
size_t onesweep_temp_bytes(int64_t n_max) {
    return ((size_t)OS_MAX_PASSES * OS_RADIX + 8 + (size_t)OS_MAX_PASSES * std::max<int64_t>(os_tiles(n_max), 1) * OS_RADIX) * 4 + 256;
}

__device__ __forceinline__ int64_t os_count(int64_t n_host, const uint32_t* __restrict__ n_dev) {
    return n_dev ? (int64_t)*reinterpret_cast<const volatile uint32_t*>(n_dev) : n_host;
}

// digit histograms of all passes + clears the look-back status words of the tiles this sort will use
__global__ void __launch_bounds__(OS_THREADS) os_hist_kernel(const uint32_t* __restrict__ keys, int64_t n_host, const uint32_t* __restrict__ n_dev,
                                                             int passes, uint32_t* __restrict__ hist, uint32_t* __restrict__ status,
                                                             int64_t tiles_max) {
    __shared__ uint32_t sh[OS_MAX_PASSES][OS_RADIX];
    const int64_t n = os_count(n_host, n_dev);
    const int64_t tiles = (n + OS_TILE - 1) / OS_TILE;
    const int64_t gtid = (int64_t)blockIdx.x * OS_THREADS + threadIdx.x, gsz = (int64_t)gridDim.x * OS_THREADS;
    for (int p = 0; p < passes; ++p)
        for (int64_t i = gtid; i < tiles * OS_RADIX; i += gsz) status[(size_t)p * tiles_max * OS_RADIX + i] = 0u;
    for (int i = threadIdx.x; i < OS_MAX_PASSES * OS_RADIX; i += OS_THREADS) (&sh[0][0])[i] = 0u;
    __syncthreads();
    for (int64_t i = gtid; i < n; i += gsz) {
        const uint32_t k = keys[i];
        for (int p = 0; p < passes; ++p) atomicAdd(&sh[p][(k >> (8 * p)) & 255u], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * OS_RADIX; i += OS_THREADS) {
        const uint32_t c = (&sh[0][0])[i];
        if (c) atomicAdd(hist + i, c);
    }
}

__global__ void __launch_bounds__(OS_THREADS, 3) os_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                                uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                                int64_t n_host, const uint32_t* __restrict__ n_dev, int shift,
                                                                const uint32_t* __restrict__ hist_p, uint32_t* status_p,
                                                                uint32_t* __restrict__ ticket_p) {
    constexpr int NW = OS_THREADS / 32;
    __shared__ uint32_t wcnt[NW][OS_RADIX];
    __shared__ uint32_t skey[OS_TILE], sval[OS_TILE];
    __shared__ uint32_t dstart[OS_RADIX], gdelta[OS_RADIX];
    __shared__ uint32_t scan_tmp[33];
    __shared__ uint32_t s_tile;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) s_tile = atomicAdd(ticket_p, 1u);
    for (int i = tid; i < NW * OS_RADIX; i += OS_THREADS) (&wcnt[0][0])[i] = 0u;
    __syncthreads();
    const int64_t tile = s_tile;
    const int64_t n = os_count(n_host, n_dev);
    if (tile * OS_TILE >= n) return;                      // block-uniform
    // global base of every digit = exclusive scan of this pass's histogram (thread d owns digit d)
    uint32_t base_d;
    {
        uint32_t tot;
        base_d = block_exclusive_scan<uint32_t>(hist_p[tid], scan_tmp, tot);
    }
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int64_t base = tile * OS_TILE + (int64_t)warp * 32 * OS_ITEMS + lane;
    uint32_t k[OS_ITEMS];                                 // (payloads are read when the tile is staged: 16 registers fewer, 3 CTAs per SM)
    uint16_t rank[OS_ITEMS];
#pragma unroll
    for (int j = 0; j < OS_ITEMS; ++j) {
        const int64_t idx = base + j * 32;
        const bool valid = idx < n;
        k[j] = valid ? keys_in[idx] : 0u;
    }
#pragma unroll
    for (int j = 0; j < OS_ITEMS; ++j) {
        const int64_t idx = base + j * 32;
        const bool valid = idx < n;
        const uint32_t d = valid ? ((k[j] >> shift) & 255u) : (uint32_t)OS_RADIX;   // tail items form their own group
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t r = __popc(peers & lt_mask);
        uint32_t old = 0;
        if (valid && r == 0) {
            old = wcnt[warp][d];
            wcnt[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1);
        rank[j] = (uint16_t)(old + r);
        __syncwarp();
    }
    __syncthreads();
    // thread d: exclusive prefix over the warps of this tile, tile count, look-back over the earlier tiles
    uint32_t run = 0, goff;
    {
        const int d = tid;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const uint32_t c = wcnt[w][d];
            wcnt[w][d] = run;
            run += c;
        }
        volatile uint32_t* st = status_p;
        uint32_t excl = 0;
        if (tile == 0) {
            st[d] = OS_PREFIX | run;
        } else {
            st[(size_t)tile * OS_RADIX + d] = OS_AGG | run;
            // look-back with four predecessors in flight: the tiles of a pass are co-resident and publish their aggregates
            // at about the same time, so a late tile walks back over many AGG entries - one dependent L2 load each when
            // walked singly (54 us per pass for 416 tiles, profiles/r02n); the loads of a batch are independent
            int64_t tt = tile - 1;
            bool done = false;
            while (!done && tt >= 0) {
                uint32_t sv[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) sv[q] = tt - q >= 0 ? (uint32_t)st[(size_t)(tt - q) * OS_RADIX + d] : (2u << 30);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (done) break;
                    if (sv[q] == 0u) break;                   // not published yet: re-read from this tile on
                    excl += sv[q] & OS_VALUE;
                    --tt;
                    if (sv[q] & OS_PREFIX) done = true;
                }
            }
            st[(size_t)tile * OS_RADIX + d] = OS_PREFIX | (excl + run);
        }
        goff = base_d + excl;
    }
    // Tile-local sort through shared memory, then a coalesced scatter: in digit order the 4096 pairs of a tile form
    // <= 256 runs, each contiguous in the output, so consecutive threads write consecutive addresses.  (Writing every
    // pair straight to its final position cost one 32-byte L2 sector transaction per 4-byte element: 54 us per pass for
    // 1.7 M pairs, 10 % issue-active - profiles/r02d.)
    {
        uint32_t tot;
        const uint32_t lstart = block_exclusive_scan<uint32_t>(run, scan_tmp, tot);     // first local position of digit `tid`
#pragma unroll
        for (int w = 0; w < NW; ++w) wcnt[w][tid] += lstart;
        dstart[tid] = lstart;
        gdelta[tid] = goff - lstart;                  // global position = gdelta[digit] + local position
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < OS_ITEMS; ++j) {
        const int64_t idx = base + j * 32;
        if (idx < n) {
            const uint32_t d = (k[j] >> shift) & 255u;
            const uint32_t lp = wcnt[warp][d] + rank[j];
            skey[lp] = k[j];
            sval[lp] = __ldg(vals_in + idx);
        }
    }
    __syncthreads();
    const int nvalid = (int)min((int64_t)OS_TILE, n - tile * OS_TILE);
#pragma unroll
    for (int j = 0; j < OS_ITEMS; ++j) {
        const int lp = tid + j * OS_THREADS;
        if (lp < nvalid) {
            const uint32_t kk = skey[lp];
            const uint32_t pos = gdelta[(kk >> shift) & 255u] + (uint32_t)lp;
            keys_out[pos] = kk;
            vals_out[pos] = sval[lp];
        }
    }
}

int onesweep_sort_pairs(uint32_t* keys[2], uint32_t* vals[2], int64_t n_max, const uint32_t* n_dev, int bits, void* temp,
                        cudaStream_t st, int64_t* launches) {
    if (n_max <= 0 || bits <= 0) return 0;
    const int passes = std::min(OS_MAX_PASSES, (bits + 7) / 8);
    const int64_t tiles_max = std::max<int64_t>(os_tiles(n_max), 1);
    uint32_t* hist = reinterpret_cast<uint32_t*>(temp);
    uint32_t* tickets = hist + OS_MAX_PASSES * OS_RADIX;
    uint32_t* status = tickets + 8;
    cudaMemsetAsync(hist, 0, (OS_MAX_PASSES * OS_RADIX + 8) * 4, st);
    const unsigned hgrid = (unsigned)std::min<int64_t>(tiles_max, 592);
    os_hist_kernel<<<hgrid, OS_THREADS, 0, st>>>(keys[0], n_max, n_dev, passes, hist, status, tiles_max);
    int cur = 0;
    for (int p = 0; p < passes; ++p) {
        os_scatter_kernel<<<(unsigned)tiles_max, OS_THREADS, 0, st>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n_max, n_dev, 8 * p,
                                                                       hist + p * OS_RADIX, status + (size_t)p * tiles_max * OS_RADIX, tickets + p);
        cur ^= 1;
    }
    if (launches) *launches += 1 + passes;
    return cur;
}

}  // namespace prims
