// Closed-form replay of TF's NON-LAZY sparse Adam (python/training/adam.py::_apply_sparse_shared, SURVEY.md §7
// hard part 1): every step the WHOLE variable is decayed and moved,
//     m <- b1 m;   v <- b2 v;   w <- w - alpha_tau * m / (sqrt(v) + eps)          for every row, touched or not.
// A row whose record was last materialised at step `last` is brought to step `upto` without walking the G = upto-last
// skipped steps.  With a = sqrt(v_last), q = sqrt(b2), A = a / (a + eps):
//     w_upto = w_last - m_last / (a + eps) * sum_{i=1..G} alpha_{last+i} b1^i / (1 - A (1 - q^i))
// and since 1 - q^i <= 1 - q^G stays below ~0.1 wherever b1^i still matters, 1/(1-x) = 1 + x + x^2 + x^3 (+ O(x^4),
// < 2e-8 relative for the TF defaults) gives a cubic in A whose four coefficients depend on (last, upto) only:
//     c_n = sum_{i=1..G} alpha_{last+i} b1^i (1 - q^i)^n
//         = T_n(last) - b1^G * sum_k C(n,k) (1 - q^G)^k U_{n,k}(upto),
//     U_{n,k}(s) = sum_{i>=1} alpha_{s+i} b1^i (1 - q^i)^(n-k) q^(ik),   T_n = U_{n,0}
// (all terms positive: no cancellation).  alpha_t is a deterministic function of t (TF's float32 running products of
// beta1 / beta2), so the U tables are computed once, ahead of time, in float64 (replay_tables_kernel).  Against the
// literal float32 step-by-step sequence the closed form is the MORE accurate of the two (measured: 3e-8 vs 3e-7
// absolute against float64 on displacements of 0.1; tests/test_gpu_replay.py) and it is O(1) per row instead of O(G).
// Hyper-parameters outside the series' range fall back to the step-by-step replay (adam_replay4).
#pragma once
#include "dfm_types.cuh"

struct ReplayTab {            // one per optimizer group (deep / linear); device pointers
    const float4* T4;         // [cap]      {T_0, T_1, T_2, T_3}(s)
    const float*  U;          // [cap][12]  U_{n,k}(s) at index n(n+1)/2 + k (10 used)
    const float*  alpha;      // [cap + H + 1] alpha_t (fallback loop)
    const float4* PQ;         // [REPLAY_PQ_N] {beta1^G, beta2^G, 1 - sqrt(beta2)^G, 0}: the per-gap factors (float64-computed)
    float l2b1, l2b2, lnq;    // log2(beta1), log2(beta2), ln(sqrt(beta2))   (gaps beyond the PQ table)
    int   closed;             // 0: step-by-step fallback
};
constexpr int REPLAY_PQ_N = 4096;

struct ReplayStep {           // per-kernel constants: the row of U at the target step
    float u[10];
    int   upto;
};

struct ReplayCoef { float c0, c1, c2, c3, d1, d2; };

__device__ __forceinline__ ReplayStep replay_step_load(const ReplayTab& rt, int upto) {
    ReplayStep rs;
    rs.upto = upto;
#pragma unroll
    for (int i = 0; i < 10; ++i) rs.u[i] = (rt.closed && upto >= 0) ? __ldg(rt.U + (size_t)upto * 12 + i) : 0.f;
    return rs;
}

__device__ __forceinline__ ReplayCoef replay_coef(const ReplayTab& rt, const ReplayStep& rs, int last) {
    const int G = rs.upto - last;
    float d1, d2, x;
    if (G < REPLAY_PQ_N) {                                // two 16-byte table reads instead of exp2 / expm1 per row
        const float4 pq = __ldg(rt.PQ + G);
        d1 = pq.x; d2 = pq.y; x = pq.z;
    } else {
        const float g = (float)G;
        d1 = exp2f(g * rt.l2b1); d2 = exp2f(g * rt.l2b2);
        x = -expm1f(g * rt.lnq);                          // 1 - q^G
    }
    const float4 T = __ldg(rt.T4 + last);
    ReplayCoef c;
    c.c0 = T.x - d1 * rs.u[0];
    c.c1 = T.y - d1 * fmaf(x, rs.u[2], rs.u[1]);
    c.c2 = T.z - d1 * fmaf(x, fmaf(x, rs.u[5], 2.f * rs.u[4]), rs.u[3]);
    c.c3 = T.w - d1 * fmaf(x, fmaf(x, fmaf(x, rs.u[9], 3.f * rs.u[8]), 3.f * rs.u[7]), rs.u[6]);
    c.d1 = d1; c.d2 = d2;
    return c;
}

__device__ __forceinline__ void replay_elem(float& w, float& m, float& v, const ReplayCoef& c, float eps) {
    float a, r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(v));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a + eps));
    const float A = a * r;
    const float p = fmaf(A, fmaf(A, fmaf(A, c.c3, c.c2), c.c1), c.c0);
    w = fmaf(-(m * r), p, w);
    m *= c.d1;
    v *= c.d2;
}

// T_n / U_{n,k} for every step s < cap from the alpha sequence (float64 accumulation; alpha holds cap + H + 1 entries)
static __global__ void replay_tables_kernel(const float* __restrict__ alpha, int cap, int H, double b1, double q,
                                     float4* __restrict__ T4, float* __restrict__ U) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= cap) return;
    double acc[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) acc[i] = 0.0;
    double pb = 1.0, pq = 1.0;
    for (int i = 1; i <= H; ++i) {
        pb *= b1; pq *= q;
        const double a = (double)alpha[s + i] * pb;
        const double x = 1.0 - pq, x2 = x * x, q2 = pq * pq;
        acc[0] += a;
        acc[1] += a * x;        acc[2] += a * pq;
        acc[3] += a * x2;       acc[4] += a * x * pq;   acc[5] += a * q2;
        acc[6] += a * x2 * x;   acc[7] += a * x2 * pq;  acc[8] += a * x * q2;   acc[9] += a * q2 * pq;
    }
    T4[s] = make_float4((float)acc[0], (float)acc[1], (float)acc[3], (float)acc[6]);
#pragma unroll
    for (int i = 0; i < 10; ++i) U[(size_t)s * 12 + i] = (float)acc[i];
    U[(size_t)s * 12 + 10] = 0.f; U[(size_t)s * 12 + 11] = 0.f;
}
