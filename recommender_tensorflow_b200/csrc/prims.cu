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
