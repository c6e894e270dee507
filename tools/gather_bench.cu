// scratch micro-benchmark: how fast can random 208-byte table records (256-byte stride) be brought on chip?
//   mode 0: cp.async.bulk (TMA 1-D) per record, issued by every warp for its own ring slot, mbarrier completion
//   mode 1: LDG.128, 4 lanes per record (w | lin | m | v quarter), U records per lane group in flight
//   mode 2: cp.async (LDGSTS) 16 B, 13 chunks per record, 32 lanes cover ~2.5 records per instruction
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 gather_bench.cu -o gather_bench
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
template <int WARPS, int RPW, int NST>   // records per warp and stage, stages
__global__ void __launch_bounds__(WARPS * 32) k_bulk(const float* __restrict__ tab, const uint32_t* __restrict__ rows, int n, float* out) {
    extern __shared__ __align__(16) float sm[];
    __shared__ uint64_t bars[WARPS * NST];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* my = sm + (size_t)warp * NST * RPW * 52;
    uint64_t* bar = bars + warp * NST;
    if (lane == 0) for (int s = 0; s < NST; ++s) mbar_init(&bar[s], 1);
    __syncwarp();
    const int gw = blockIdx.x * WARPS + warp, nw = gridDim.x * WARPS;
    const int chunks = n / RPW;
    float acc = 0.f;
    int issued = 0, used = 0;
    auto issue = [&](int c) {
        const int s = issued % NST;
        if (lane == 0) mbar_expect_tx(&bar[s], RPW * 208);
        __syncwarp();
        for (int r = lane; r < RPW; r += 32) bulk_load(my + (s * RPW + r) * 52, tab + (size_t)rows[c * RPW + r] * 64, 208, &bar[s]);
        ++issued;
    };
    int c = gw;
    for (int p = 0; p < NST - 1 && c < chunks; ++p, c += nw) issue(c);
    for (int cc = gw; cc < chunks; cc += nw) {
        if (c < chunks) { issue(c); c += nw; }
        const int s = used % NST;
        mbar_wait(&bar[s], (used / NST) & 1);
        for (int r = lane; r < RPW * 13; r += 32) { const float4 v = reinterpret_cast<const float4*>(my + s * RPW * 52)[r]; acc += v.x + v.w; }
        ++used;
        __syncwarp();
    }
    if (acc == 12345.f) out[0] = acc;
}
template <int U>
__global__ void __launch_bounds__(256) k_ldg(const float* __restrict__ tab, const uint32_t* __restrict__ rows, int n, float* out) {
    const int lane = threadIdx.x & 31, sub = lane & 3, grp = lane >> 2;
    const int gw = (blockIdx.x * 256 + threadIdx.x) >> 5, nw = (gridDim.x * 256) >> 5;
    float acc = 0.f;
    for (int base = gw * 8 * U; base + 8 * U <= n; base += nw * 8 * U) {
        float4 a[U], b[U], c2[U], d[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float* rec = tab + (size_t)rows[base + u * 8 + grp] * 64;
            a[u] = __ldg(reinterpret_cast<const float4*>(rec) + sub);
            b[u] = __ldg(reinterpret_cast<const float4*>(rec + 16));
            c2[u] = __ldg(reinterpret_cast<const float4*>(rec + 20) + sub);
            d[u] = __ldg(reinterpret_cast<const float4*>(rec + 36) + sub);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += a[u].x + b[u].y + c2[u].z + d[u].w;
    }
    if (acc == 12345.f) out[0] = acc;
}
int main(int argc, char** argv) {
    const size_t R = 100u * 1000 * 1000;      // 25.6 GB table
    const int n = 65536 * 26;
    float* tab; uint32_t* rows; float* out;
    cudaMalloc(&tab, R * 256); cudaMalloc(&rows, n * 4); cudaMalloc(&out, 4);
    cudaMemset(tab, 0, R * 256);
    std::vector<uint32_t> h(n);
    uint64_t s = 88172645463325252ull;
    for (int i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (uint32_t)(s % R); }
    cudaMemcpy(rows, h.data(), n * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](const char* name, auto launch) {
        for (int i = 0; i < 2; ++i) launch();
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) launch();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
        printf("%-40s %.3f ms  %.2f TB/s (208 B/record)  err=%s\n", name, ms, n * 208.0 / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    };
    {
        auto k = k_bulk<8, 32, 3>; int smem = 8 * 3 * 32 * 208;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        timeit("bulk 8 warps x 32 rec x 3 stages, 1 CTA/SM", [&] { k<<<148, 256, smem>>>(tab, rows, n, out); });
    }
    {
        auto k = k_bulk<16, 16, 3>; int smem = 16 * 3 * 16 * 208;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        timeit("bulk 16 warps x 16 rec x 3 stages", [&] { k<<<148, 512, smem>>>(tab, rows, n, out); });
    }
    {
        auto k = k_bulk<1, 208, 3>; int smem = 3 * 208 * 208;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        timeit("bulk 1 warp x 208 rec x 3 stages", [&] { k<<<148, 32, smem>>>(tab, rows, n, out); });
    }
    {
        auto k = k_bulk<4, 64, 3>; int smem = 4 * 3 * 64 * 208;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        timeit("bulk 4 warps x 64 rec x 3 stages", [&] { k<<<148, 128, smem>>>(tab, rows, n, out); });
    }
    timeit("ldg U=1, 8 CTA/SM", [&] { k_ldg<1><<<148 * 8, 256>>>(tab, rows, n, out); });
    timeit("ldg U=2, 8 CTA/SM", [&] { k_ldg<2><<<148 * 8, 256>>>(tab, rows, n, out); });
    timeit("ldg U=4, 4 CTA/SM", [&] { k_ldg<4><<<148 * 4, 256>>>(tab, rows, n, out); });
    timeit("ldg U=4, 8 CTA/SM", [&] { k_ldg<4><<<148 * 8, 256>>>(tab, rows, n, out); });
    timeit("ldg U=2, 2 CTA/SM", [&] { k_ldg<2><<<148 * 2, 256>>>(tab, rows, n, out); });
    return 0;
}
