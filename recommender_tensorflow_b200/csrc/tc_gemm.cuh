// Tensor-core path of the DNN tower for hidden_units that make it a real GEMM (BASELINE.json
// configs[2]: hidden [256,128], batch 65536 -> 55 GFLOP per step).
//
// fp32 parity (1e-5 relative) rules out plain TF32 (10-bit mantissa), so every product is formed
// as a 3xTF32 split accumulated in fp32 in tensor memory:
//        a*b  ~=  a_hi*b_hi + a_lo*b_hi + a_hi*b_lo ,   x_hi = tf32_rn(x), x_lo = tf32_rn(x - x_hi)
// (relative error <= ~2^-21 per product, measured against the fp64 oracle in the tests).
//
// Blackwell mapping (one CTA per 128 x BN output tile, warp-specialised):
//   warp 0      TMA producer: cp.async.bulk.tensor tiles (SWIZZLE_128B) of A and B into a ring of
//               shared-memory stages, completion on "full" mbarriers
//   warps 2..7  splitters: rewrite each landed fp32 tile in place as x_hi and write x_lo beside it
//               (element-wise, so the swizzle pattern is preserved), fence.proxy.async, arrive on
//               "split" mbarriers
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma kind::tf32 (M=128, N=BN, K=8), three
//               per k-step, accumulating in TMEM; tcgen05.commit releases the stage ("empty") and
//               finally signals the epilogue
//   warps 4..7  (after the main loop) epilogue: tcgen05.ld the fp32 accumulator (32 TMEM lanes per warp), apply the fused
//               epilogue (bias+ReLU / ReLU mask / FM-gradient add / split-K partial) and store
// Operand layouts: K-major (row-major [rows, K]) for the forward and backward-data GEMMs;
// MN-major (row-major [K, rows]) for the weight-gradient GEMM whose reduction runs over the batch.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "mlp_kernels.cuh"

namespace tc {

constexpr int BM = 128;          // output rows per CTA (= TMEM lanes)
// k-block depth BK is a template parameter: 32 fp32 (128-byte swizzle rows) or 16 (64-byte rows); smaller
// k-blocks buy a deeper TMA ring in the same shared memory (the profile showed the ring depth, not the
// tensor pipe, limiting the main loop).
constexpr int UMMA_K = 8;        // tf32 K per tcgen05.mma
constexpr int NTHREADS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (spins > (1u << 26)) __trap();   // a protocol bug must fault, not hang the GPU
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Explicit shared-state-space accesses.  The operand stages are reached through pointers derived from the dynamic
// shared-memory base by integer arithmetic, and the compiler then emits GENERIC LD.E/ST.E for plain C++ accesses
// (ncu source page of the first persistent kernel: every splitter access was LD.E.128 / ST.E.128 with long-scoreboard
// stalls); ld.shared / st.shared on 32-bit shared addresses take the short path.
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, const float4& v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// shared-memory matrix descriptor (sm_100 "version 1"), SWIZZLE_128B
//   K-major : rows of 128 B (32 fp32 of K), 8-row atoms 1024 B apart (SBO); LBO unused
//   MN-major (32-bit operands must use SWIZZLE_128B_BASE32B, TMA mode 128B_ATOM_32B): rows of 128 B
//             (32 fp32 of M/N), row index = k; 4-k-row atoms SBO = 512 B apart, 32-wide M/N chunks LBO apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;      // descriptor version (Blackwell)
    d |= (uint64_t)layout_type << 61;   // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
    return d;
}
// instruction descriptor: D fp32, A/B tf32, dense, M=128, N
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn_major, bool b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// round-to-nearest (ties away) to tf32: hi keeps 11 significant bits, so |x - hi| <= 2^-12 |x| and the
// second split leaves a residual of ~2^-24 |x| (fp32 level); the low 13 bits are zero so the tensor core
// sees exactly representable operands whatever its own input rounding is.
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }

__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
#ifndef TC_TRUNC_HI
#define TC_TRUNC_HI 0   // 1 = rely on kind::tf32 input truncation (5% faster, but one accuracy test was marginal)
#endif

struct Params {
    int M, N, K;              // C[M,N] = A(M,K) * B(K,N); K-major mode: K % 32 == 0 not required (TMA zero-fills)
    int k_per_split;          // MN-major (weight-gradient) mode: K range per blockIdx.z
    float* C; int ldc; size_t c_split_stride;
    int epi;                  // EPI_*
    EpiArgs ep;
    int split_a, split_b;     // 1: operand is raw fp32, split in the kernel; 0: hi/lo come from two tensor maps
    int dbg;                  // experiments only: 1 = skip the split work, 2 = skip the MMAs
    unsigned long long* ts;   // experiments only: per-CTA phase timestamps (8 per CTA), see dfm_test_tc_gemm mode 2
    int bn;                   // output columns per CTA (multiple of 16, <= BN): N is cut into equal tiles so that
                              // e.g. N = 416 runs as 2 x 208 instead of 256 + 160 (37% of the second tile wasted)
};

// Fused epilogue of one 128 x bn output tile.  Warp (q, half): TMEM lanes [32q, 32q+32), 32-column chunks half, half+2, ...
// Each 32x32 block is transposed through the warp's shared-memory staging area `tb` (TMEM yields one row per lane) so
// that every global access is a coalesced 16-byte-per-lane access, with all loads of a block issued up front.
__device__ __forceinline__ void epilogue_tile(const Params& p, float* __restrict__ Cz, float* tb, uint32_t tmem_main, uint32_t tmem_cross,
                                              int m0, int n0, int bn, int q, int half, int lane) {
    const int row0 = m0 + q * 32;
    const int rsub = lane >> 3, cg = lane & 7;
    const uint32_t tbs = smem_u32(tb);
    for (int c0 = half * 32; c0 < bn; c0 += 64) {
        if (n0 + c0 >= p.N) break;
        {
            float v[32], vx[32];
            tmem_ld32(tmem_main + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            tmem_ld32(tmem_cross + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, vx);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                sts128(tbs + (lane * 36 + 4 * j) * 4,
                       make_float4(v[4 * j] + vx[4 * j], v[4 * j + 1] + vx[4 * j + 1], v[4 * j + 2] + vx[4 * j + 2], v[4 * j + 3] + vx[4 * j + 3]));
        }
        __syncwarp();
        const int n = n0 + c0 + cg * 4;
        const bool nok = n < p.N && (c0 + cg * 4) < bn;   // N % 4 == 0 and bn % 16 == 0 on this path
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.epi == EPI_BIAS_RELU && nok) bias4 = __ldg(reinterpret_cast<const float4*>(p.ep.bias + n));
        const int nk = (p.epi == EPI_DE) ? (n % p.ep.K) : 0;
        float4 xa[8], xs[8];
        float xd[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int gm = row0 + it * 4 + rsub;
            xa[it] = make_float4(0.f, 0.f, 0.f, 0.f); xs[it] = xa[it]; xd[it] = 0.f;
            if (nok && gm < p.M) {
                if (p.epi == EPI_MASK) xa[it] = __ldg(reinterpret_cast<const float4*>(p.ep.act + (size_t)gm * p.ep.ld_act + n));
                else if (p.epi == EPI_DE && p.ep.s) {
                    xa[it] = __ldg(reinterpret_cast<const float4*>(p.ep.act + (size_t)gm * p.ep.ld_act + n));
                    xs[it] = __ldg(reinterpret_cast<const float4*>(p.ep.s + (size_t)gm * p.ep.K + nk));
                    xd[it] = __ldg(p.ep.dz + gm);
                }
            }
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + rsub, gm = row0 + r;
            float4 x = lds128(tbs + (r * 36 + cg * 4) * 4);
            if (p.epi == EPI_BIAS_RELU) {
                x.x = fmaxf(x.x + bias4.x, 0.f); x.y = fmaxf(x.y + bias4.y, 0.f);
                x.z = fmaxf(x.z + bias4.z, 0.f); x.w = fmaxf(x.w + bias4.w, 0.f);
                if (p.ep.drop_keep > 0.f) {
                    const uint64_t e0 = (uint64_t)(p.ep.drop_row0 + gm) * p.N + n;
                    x.x *= dfm_drop(p.ep.drop_key, e0, p.ep.drop_keep, p.ep.drop_inv);
                    x.y *= dfm_drop(p.ep.drop_key, e0 + 1, p.ep.drop_keep, p.ep.drop_inv);
                    x.z *= dfm_drop(p.ep.drop_key, e0 + 2, p.ep.drop_keep, p.ep.drop_inv);
                    x.w *= dfm_drop(p.ep.drop_key, e0 + 3, p.ep.drop_keep, p.ep.drop_inv);
                }
            } else if (p.epi == EPI_MASK) {
                const float sc = p.ep.bwd_scale > 0.f ? p.ep.bwd_scale : 1.f;
                x.x = xa[it].x > 0.f ? x.x * sc : 0.f; x.y = xa[it].y > 0.f ? x.y * sc : 0.f;
                x.z = xa[it].z > 0.f ? x.z * sc : 0.f; x.w = xa[it].w > 0.f ? x.w * sc : 0.f;
            } else if (p.epi == EPI_DE && p.ep.s) {
                x.x += xd[it] * (xs[it].x - xa[it].x); x.y += xd[it] * (xs[it].y - xa[it].y);
                x.z += xd[it] * (xs[it].z - xa[it].z); x.w += xd[it] * (xs[it].w - xa[it].w);
            }
            if (nok && gm < p.M) *reinterpret_cast<float4*>(Cz + (size_t)gm * p.ldc + n) = x;
        }
        __syncwarp();
    }
}

template <int BN, int BK>
struct Smem {
    static constexpr int A_BYTES = BM * BK * 4;
    static constexpr int B_BYTES = BN * BK * 4;
    static constexpr int STAGE = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int STAGES = (192 * 1024) / STAGE > 8 ? 8 : (192 * 1024) / STAGE;
    static_assert(STAGES >= 2, "operand ring needs two stages");
    static_assert(STAGES * STAGE >= 8 * 32 * 36 * 4, "epilogue transposes through the operand stages");
    static constexpr int TOTAL = STAGES * STAGE + 1024 /*align*/ + 512 /*barriers*/;
};

// MODE 0: A K-major [M,K], B K-major [N,K].   MODE 1: A MN-major [K,M], B MN-major [K,N] (3-D maps, see host).
template <int BN, int MODE, int BK>
__global__ void __launch_bounds__(NTHREADS, 1) gemm_kernel(const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,
                                                           const __grid_constant__ CUtensorMap mapB_hi, const __grid_constant__ CUtensorMap mapB_lo,
                                                           Params p) {
    using S = Smem<BN, BK>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::STAGES * S::STAGE);
    uint64_t* full = bars;                       // [STAGES] TMA landed
    uint64_t* split = bars + S::STAGES;          // [STAGES] hi/lo ready
    uint64_t* empty = bars + 2 * S::STAGES;      // [STAGES] MMAs that read the stage retired
    uint64_t* accum = bars + 3 * S::STAGES;      // accumulator complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S::STAGES + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bn = p.bn;      // runtime tile width; BN is the capacity (stage / TMEM sizing)
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * bn;
    const int kbeg = MODE == 1 ? blockIdx.z * p.k_per_split : 0;
    const int kend = MODE == 1 ? min(p.K, kbeg + p.k_per_split) : p.K;
    const int nkb = (kend - kbeg + BK - 1) / BK;
    constexpr uint32_t TMEM_COLS = 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;
    constexpr int n_split_threads = 192;   // warps 2..7 split; warps 4..7 then run the epilogue

    if (threadIdx.x == 0) {
        for (int s = 0; s < S::STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&split[s], n_split_threads);
            mbar_init(&empty[s], 1);
        }
        mbar_init(accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    auto stamp = [&](int slot) {
        if (p.ts) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            p.ts[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 + slot] = t;
        }
    };
    if (threadIdx.x == 0) {
        stamp(0);
        if (p.ts) { uint32_t smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); p.ts[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 + 7] = smid; }
    }

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            const uint32_t b_bytes = (uint32_t)bn * BK * 4;
            const uint32_t bytes = (p.split_a ? S::A_BYTES : 2 * S::A_BYTES) + (p.split_b ? b_bytes : 2 * b_bytes);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % S::STAGES;
                if (kb >= S::STAGES) mbar_wait(&empty[s], ((kb / S::STAGES) - 1) & 1);
                uint8_t* st = smem + s * S::STAGE;
                uint8_t *a_hi = st, *a_lo = st + S::A_BYTES, *b_hi = st + 2 * S::A_BYTES, *b_lo = st + 2 * S::A_BYTES + S::B_BYTES;
                mbar_expect_tx(&full[s], bytes);
                const int k0 = kbeg + kb * BK;
                if (MODE == 0) {
                    tma_load_2d(a_hi, &mapA_hi, &full[s], k0, m0);
                    if (!p.split_a) tma_load_2d(a_lo, &mapA_lo, &full[s], k0, m0);
                    tma_load_2d(b_hi, &mapB_hi, &full[s], k0, n0);
                    if (!p.split_b) tma_load_2d(b_lo, &mapB_lo, &full[s], k0, n0);
                } else {
                    tma_load_3d(a_hi, &mapA_hi, &full[s], 0, k0, m0 / 32);
                    tma_load_3d(b_hi, &mapB_hi, &full[s], 0, k0, n0 / 32);
                }
            }
            stamp(1);      // last TMA issued
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc(bn, MODE == 1, MODE == 1);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % S::STAGES;
                mbar_wait(&split[s], (kb / S::STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = smem_u32(smem + s * S::STAGE);
                const uint32_t a_hi = st, a_lo = st + S::A_BYTES, b_hi = st + 2 * S::A_BYTES, b_lo = st + 2 * S::A_BYTES + S::B_BYTES;
#pragma unroll
                for (int ks = 0; ks < ((p.dbg & 2) ? 0 : BK / UMMA_K); ++ks) {
                    // K-major: +32 bytes inside the 128-byte swizzle row; MN-major: next 8-row (1024 B) atom
                    const uint32_t adv = MODE == 0 ? ks * UMMA_K * 4 : ks * 1024;
                    // K-major: 8-row atoms of BK*4-byte rows (SWIZZLE_128B for BK=32, SWIZZLE_64B for BK=16)
                    const uint32_t lbo = MODE == 0 ? 16 : BK * 128, sbo = MODE == 0 ? 8 * BK * 4 : 512;
                    const uint32_t lt = MODE == 0 ? (BK == 32 ? 2 : 4) : 1;
                    const uint64_t dah = make_desc(a_hi + adv, lbo, sbo, lt), dal = make_desc(a_lo + adv, lbo, sbo, lt);
                    const uint64_t dbh = make_desc(b_hi + adv, lbo, sbo, lt), dbl = make_desc(b_lo + adv, lbo, sbo, lt);
                    // The tensor core truncates when it folds a product group into the fp32 accumulator, so the
                    // large hi*hi sum and the small cross terms live in separate accumulators (columns [0,BN) and
                    // [BN,2BN)) and are added with round-to-nearest in the epilogue.
                    umma_tf32(tmem_base, dah, dbh, idesc, (kb | ks) != 0);
                    umma_tf32(tmem_base + BN, dal, dbh, idesc, (kb | ks) != 0);
                    umma_tf32(tmem_base + BN, dah, dbl, idesc, 1);
                }
                umma_commit(&empty[s]);
            }
            umma_commit(accum);
            stamp(2);      // last MMA issued
        }
    } else {
        // ------------------------------------------------------------------ splitters (warps 2..7, 192 threads)
        const int t = threadIdx.x - 64;
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % S::STAGES;
            mbar_wait(&full[s], (kb / S::STAGES) & 1);
            uint8_t* st = smem + s * S::STAGE;
            if (p.split_a && !(p.dbg & 1)) {
                const uint32_t hi = smem_u32(st), lo = smem_u32(st + S::A_BYTES);
#pragma unroll 4
                for (int i = t; i < S::A_BYTES / 16; i += n_split_threads) {
                    float4 x = lds128(hi + i * 16);
#if TC_TRUNC_HI
                    // kind::tf32 ignores the low 13 mantissa bits of its fp32 operands (verified by the accuracy
                    // tests: with any other input rounding this split loses 2^-11), so the landed tile IS x_hi
                    // and only x_lo = tf32_rn(x - trunc(x)) is written -> one third less shared-memory traffic.
                    float4 l = make_float4(tf32_hi(x.x - tf32_trunc(x.x)), tf32_hi(x.y - tf32_trunc(x.y)),
                                           tf32_hi(x.z - tf32_trunc(x.z)), tf32_hi(x.w - tf32_trunc(x.w)));
                    sts128(lo + i * 16, l);
#else
                    float4 h = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
                    float4 l = make_float4(tf32_hi(x.x - h.x), tf32_hi(x.y - h.y), tf32_hi(x.z - h.z), tf32_hi(x.w - h.w));
                    sts128(hi + i * 16, h);
                    sts128(lo + i * 16, l);
#endif
                }
            }
            if (p.split_b && !(p.dbg & 1)) {
                const uint32_t hi = smem_u32(st + 2 * S::A_BYTES), lo = smem_u32(st + 2 * S::A_BYTES + S::B_BYTES);
#pragma unroll 4
                for (int i = t; i < bn * BK * 4 / 16; i += n_split_threads) {
                    float4 x = lds128(hi + i * 16);
#if TC_TRUNC_HI
                    // kind::tf32 ignores the low 13 mantissa bits of its fp32 operands (verified by the accuracy
                    // tests: with any other input rounding this split loses 2^-11), so the landed tile IS x_hi
                    // and only x_lo = tf32_rn(x - trunc(x)) is written -> one third less shared-memory traffic.
                    float4 l = make_float4(tf32_hi(x.x - tf32_trunc(x.x)), tf32_hi(x.y - tf32_trunc(x.y)),
                                           tf32_hi(x.z - tf32_trunc(x.z)), tf32_hi(x.w - tf32_trunc(x.w)));
                    sts128(lo + i * 16, l);
#else
                    float4 h = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
                    float4 l = make_float4(tf32_hi(x.x - h.x), tf32_hi(x.y - h.y), tf32_hi(x.z - h.z), tf32_hi(x.w - h.w));
                    sts128(hi + i * 16, h);
                    sts128(lo + i * 16, l);
#endif
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05.mma
            mbar_arrive(&split[s]);
            if (t == 0 && kb == 0) stamp(3);          // first stage landed and split
        }
        if (t == 0) stamp(4);                         // last split done
    }
    {
        // ------------------------------------------------------------------ epilogue (all 8 warps)
        // Warp w may read TMEM lanes [32*(w%4), +32); warps w and w+4 share a lane quarter and take alternate
        // 32-column chunks.  All MMAs have retired, so the operand stages are free: each 32x32 block is
        // transposed through shared memory (TMEM yields one row per lane) so that every global access of the
        // fused epilogue is a coalesced 16-byte-per-lane access, with all loads of a block issued up front.
        __syncwarp();
        const int q = warp & 3, half = warp >> 2;
        mbar_wait(accum, 0);
        if (threadIdx.x == 64) stamp(5);              // accumulator complete
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float* Cz = p.C + (size_t)blockIdx.z * p.c_split_stride;
        float* tb = reinterpret_cast<float*>(smem) + warp * (32 * 36);
        epilogue_tile(p, Cz, tb, tmem_base, tmem_base + BN, m0, n0, bn, q, half, lane);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) stamp(6);                   // epilogue done
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
}


// ------------------------------------------------------------------------------------------------------------
// Persistent variant for the K-major GEMMs (forward and backward-data of the tower).  The per-CTA timeline of the
// kernel above (globaltimer stamps, M=65536 N=256 K=160) is: 1.1 us launch gap, 1.8 us until the first stage is
// usable, 6 us main loop, 5.8 us epilogue, strictly one after the other and in lock-step on all SMs, so HBM sees a
// read burst followed by a write burst.  Here one CTA per SM walks over the output tiles and the phases overlap:
//   * tiles are 128 x bn with bn <= 128, so TMEM holds TWO accumulator sets (main + cross, 2 x 256 columns): the
//     MMA warp fills set (j+1)&1 while the epilogue warps drain set j&1 ("acc_full" / "acc_empty" mbarriers);
//   * the operand ring keeps running across tile boundaries (TMA, split and MMA never drain between tiles);
//   * the epilogue has its own 8 warps and its own staging memory.
// Warps: 0 TMA, 1 MMA, 2..5 splitters (128 threads), 6..13 epilogue (warp w: TMEM lane quarter w % 4).
constexpr int P_BN = 128;
constexpr int P_BK = 16;        // 32 KB stages -> a 5-deep operand ring (the 2-deep ring of BK = 32 left every role waiting)
constexpr int P_THREADS = 448;
constexpr int P_SPLIT_THREADS = 128;
constexpr int P_SPLIT_GROUPS = 1;     // >1: groups of splitter warps take alternate stages (measured: no gain, the split is
constexpr int P_SPLIT_GROUP = P_SPLIT_THREADS / P_SPLIT_GROUPS;   // throughput-bound on shared memory, not a latency chain)
constexpr int P_EPI_WARPS = 8;

// In-place hi/lo split of one landed operand tile by the P_SPLIT_THREADS splitter threads.  All loads of a thread are
// issued before its first store: hi[] is read and written, so the compiler would otherwise order every load behind the
// previous store and the loop would run at one shared-memory round trip per element group.
template <int CAP4>
__device__ __forceinline__ void split_tile(uint8_t* hi_bytes, uint8_t* lo_bytes, int n4, int tt) {
    constexpr int PER = (CAP4 + P_SPLIT_GROUP - 1) / P_SPLIT_GROUP;
    const uint32_t hi = smem_u32(hi_bytes), lo = smem_u32(lo_bytes);
    float4 x[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const int i = tt + u * P_SPLIT_GROUP;
        if (i < n4) x[u] = lds128(hi + i * 16);
    }
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const int i = tt + u * P_SPLIT_GROUP;
        if (i < n4) {
            const float4 h = make_float4(tf32_hi(x[u].x), tf32_hi(x[u].y), tf32_hi(x[u].z), tf32_hi(x[u].w));
            sts128(hi + i * 16, h);
            sts128(lo + i * 16, make_float4(tf32_hi(x[u].x - h.x), tf32_hi(x[u].y - h.y), tf32_hi(x[u].z - h.z), tf32_hi(x[u].w - h.w)));
        }
    }
}

template <int BK>
struct PSmem {
    static constexpr int A_BYTES = BM * BK * 4;
    static constexpr int B_BYTES = P_BN * BK * 4;
    static constexpr int STAGE = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int STAGING = P_EPI_WARPS * 32 * 36 * 4;
    static constexpr int STAGES = (224 * 1024 - STAGING - 2048) / STAGE;
    static_assert(STAGES >= 2, "operand ring needs two stages");
    static constexpr int TOTAL = STAGES * STAGE + STAGING + 1024 /*align*/ + 512 /*barriers*/;
};

template <int BK>
__global__ void __launch_bounds__(P_THREADS, 1) gemm_persist_kernel(const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,
                                                                    const __grid_constant__ CUtensorMap mapB_hi, const __grid_constant__ CUtensorMap mapB_lo,
                                                                    Params p) {
    using S = PSmem<BK>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* staging = reinterpret_cast<float*>(smem + S::STAGES * S::STAGE);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::STAGES * S::STAGE + S::STAGING);
    uint64_t* full = bars;                          // [STAGES] TMA landed
    uint64_t* split = bars + S::STAGES;             // [STAGES] hi/lo ready
    uint64_t* empty = bars + 2 * S::STAGES;         // [STAGES] MMAs that read the stage retired
    uint64_t* acc_full = bars + 3 * S::STAGES;      // [2] accumulator set complete
    uint64_t* acc_empty = bars + 3 * S::STAGES + 2; // [2] accumulator set drained by the epilogue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S::STAGES + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bn = p.bn;
    const int ntile = (p.N + bn - 1) / bn;
    const int total = ((p.M + BM - 1) / BM) * ntile;
    const int nkb = (p.K + BK - 1) / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S::STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&split[s], P_SPLIT_GROUP);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], P_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    // experiments only (p.ts != nullptr): cycles each role spends waiting, per CTA
    unsigned long long* ts = p.ts ? p.ts + (size_t)blockIdx.x * 8 : nullptr;
    long long w0 = 0, w1 = 0, c_begin = clock64();
#define TS_WAIT(acc, stmt) do { if (ts) { long long c0_ = clock64(); stmt; acc += clock64() - c0_; } else { stmt; } } while (0)

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            const uint32_t b_bytes = (uint32_t)bn * BK * 4;
            const uint32_t bytes = (p.split_a ? S::A_BYTES : 2 * S::A_BYTES) + (p.split_b ? b_bytes : 2 * b_bytes);
            uint32_t it = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const int m0 = (t / ntile) * BM, n0 = (t % ntile) * bn;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const uint32_t s = it % S::STAGES;
                    TS_WAIT(w0, mbar_wait(&empty[s], ((it / S::STAGES) & 1) ^ 1));     // passes at once on the first lap
                    uint8_t* st = smem + s * S::STAGE;
                    uint8_t *a_hi = st, *a_lo = st + S::A_BYTES, *b_hi = st + 2 * S::A_BYTES, *b_lo = b_hi + b_bytes;   // [B_hi ; B_lo] = one 2bn-row operand
                    mbar_expect_tx(&full[s], bytes);
                    const int k0 = kb * BK;
                    tma_load_2d(a_hi, &mapA_hi, &full[s], k0, m0);
                    if (!p.split_a) tma_load_2d(a_lo, &mapA_lo, &full[s], k0, m0);
                    tma_load_2d(b_hi, &mapB_hi, &full[s], k0, n0);
                    if (!p.split_b) tma_load_2d(b_lo, &mapB_lo, &full[s], k0, n0);
                }
            }
            if (ts) { ts[1] = w0; }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc(bn, false, false), idesc2 = make_idesc(2 * bn, false, false);
            uint32_t it = 0, j = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x, ++j) {
                const uint32_t buf = j & 1;
                TS_WAIT(w1, mbar_wait(&acc_empty[buf], ((j >> 1) & 1) ^ 1));           // the epilogue drained this set (first use: free)
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_main = tmem_base + buf * (2 * P_BN), d_cross = d_main + bn;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const uint32_t s = it % S::STAGES;
                    TS_WAIT(w0, mbar_wait(&split[s], (it / S::STAGES) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = smem_u32(smem + s * S::STAGE);
                    const uint32_t a_hi = st, a_lo = st + S::A_BYTES, b_hi = st + 2 * S::A_BYTES;
#pragma unroll
                    for (int ks = 0; ks < ((p.dbg & 2) ? 0 : BK / UMMA_K); ++ks) {
                        const uint32_t adv = ks * UMMA_K * 4;
                        const uint32_t lbo = 16, sbo = 8 * BK * 4, lt = BK == 32 ? 2 : 4;
                        const uint64_t dah = make_desc(a_hi + adv, lbo, sbo, lt), dal = make_desc(a_lo + adv, lbo, sbo, lt);
                        const uint64_t dbh = make_desc(b_hi + adv, lbo, sbo, lt);
                        // B_lo sits right behind the bn rows of B_hi, so hi*hi (columns [0,bn)) and hi*lo (columns
                        // [bn,2bn), the cross accumulator) are ONE N = 2bn instruction that reads A_hi once: five
                        // operand-tile reads per k-step instead of six (shared-memory bandwidth is the bound)
                        umma_tf32(d_main, dah, dbh, idesc2, (kb | ks) != 0);
                        umma_tf32(d_cross, dal, dbh, idesc, 1);
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(&acc_full[buf]);
            }
            if (ts) { ts[2] = w0; ts[3] = w1; }
        }
    } else if (warp < 2 + P_SPLIT_THREADS / 32) {
        // ------------------------------------------------------------------ splitters
        const int grp = (threadIdx.x - 64) / P_SPLIT_GROUP, tt = (threadIdx.x - 64) % P_SPLIT_GROUP;
        uint32_t it = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                if ((int)(it % P_SPLIT_GROUPS) != grp) continue;
                const uint32_t s = it % S::STAGES;
                TS_WAIT(w0, mbar_wait(&full[s], (it / S::STAGES) & 1));
                uint8_t* st = smem + s * S::STAGE;
                if (p.split_a && !(p.dbg & 1)) split_tile<S::A_BYTES / 16>(st, st + S::A_BYTES, S::A_BYTES / 16, tt);
                if (p.split_b && !(p.dbg & 1)) split_tile<S::B_BYTES / 16>(st + 2 * S::A_BYTES, st + 2 * S::A_BYTES + bn * BK * 4, bn * BK * 4 / 16, tt);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05.mma
                mbar_arrive(&split[s]);
            }
        }
        if (ts && tt == 0 && grp == 0) { ts[4] = w0; ts[6] = clock64() - c_begin - w0; }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 6..13)
        const int ew = warp - (2 + P_SPLIT_THREADS / 32);
        const int q = warp & 3, half = ew >> 2;      // TMEM lane quarter is fixed by warp % 4
        float* tb = staging + ew * (32 * 36);
        uint32_t j = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++j) {
            const int m0 = (t / ntile) * BM, n0 = (t % ntile) * bn;
            const uint32_t buf = j & 1;
            TS_WAIT(w0, mbar_wait(&acc_full[buf], (j >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d_main = tmem_base + buf * (2 * P_BN);
            epilogue_tile(p, p.C, tb, d_main, d_main + bn, m0, n0, bn, q, half, lane);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        if (ts && ew == 0 && lane == 0) { ts[5] = w0; ts[7] = clock64() - c_begin - w0; }
    }
#undef TS_WAIT
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (ts && threadIdx.x == 0) ts[0] = clock64() - c_begin;
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}

// x -> (hi, lo) element-wise over a weight tensor (weights are static within a step)
static __global__ void split_tf32_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ hi, float* __restrict__ lo) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        float v = x[i], h = tf32_hi(v);
        hi[i] = h;
        lo[i] = tf32_hi(v - h);
    }
}
// W [rows, cols] -> W^T hi / lo [cols, rows]
static __global__ void split_tf32_transpose_kernel(const float* __restrict__ x, int rows, int cols, float* __restrict__ hi, float* __restrict__ lo) {
    __shared__ float tile[32][33];
    int c = blockIdx.x * 32 + threadIdx.x, r0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int r = r0 + j;
        tile[j][threadIdx.x] = (r < rows && c < cols) ? x[(size_t)r * cols + c] : 0.f;
    }
    __syncthreads();
    int r = r0 + threadIdx.x;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int cc = blockIdx.x * 32 + j;
        if (cc < cols && r < rows) {
            float v = tile[threadIdx.x][j];
            if (lo) {
                float h = tf32_hi(v);
                hi[(size_t)cc * rows + r] = h;
                lo[(size_t)cc * rows + r] = tf32_hi(v - h);
            } else {
                hi[(size_t)cc * rows + r] = v;     // plain transpose
            }
        }
    }
}

// ---------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// row-major [rows, cols] fp32 (cols contiguous, leading dimension ld): box = bk cols x box_rows, swizzle = row bytes
inline bool make_map_2d(CUtensorMap* m, const float* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t bk) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 4};
    cuuint32_t box[2] = {bk, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               bk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// row-major [krows, cols] viewed as [cols/32][krows][32]: box = 32 x box_k x chunks  (MN-major operand tiles)
inline bool make_map_3d(CUtensorMap* m, const float* base, uint64_t krows, uint64_t cols, uint64_t ld, uint32_t box_k, uint32_t chunks) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[3] = {32, krows, cols / 32};
    cuuint64_t strides[2] = {ld * 4, 32 * 4};
    cuuint32_t box[3] = {32, box_k, chunks};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tc
