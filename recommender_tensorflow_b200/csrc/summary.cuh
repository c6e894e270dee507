// layer_summary side outputs (trainers/model_utils.py:4-6, call sites trainers/deep_fm.py:43,89,105,110,115): the
// fraction of zero values and the histogram of the linear / MF logits, every hidden layer's output (after dropout in
// TRAIN mode), the DNN logit and the final logits.  The fused train kernels never hold these tensors outside an SM, so a
// summary step recomputes them with a plain forward pass (the reference writes summaries every save_summary_steps = 100
// steps): the table gather is the library's own gather kernel, the tower is the straightforward kernel below, and one
// reduction kernel turns a tensor into the fields of TensorFlow's HistogramProto over its default bucket limits
// (core/lib/histogram/histogram.cc: +-1e-12 * 1.1^i, 0, +-DBL_MAX; bucket = upper_bound(limits, value)).
#pragma once
#include "dfm_types.cuh"
#include "mlp_kernels.cuh"

struct SummaryTowerArgs {
    const float* h0; int dK;                 // input layer [B, dK]
    const float* dw;                         // packed dense parameters
    int L; int H[DFM_MAX_HIDDEN]; int off_W[DFM_MAX_HIDDEN]; int off_b[DFM_MAX_HIDDEN]; int off_Wo, off_bo;
    int B, maxdim, hid_stride;               // hid_stride = sum of H
    float drop_keep, drop_inv; uint64_t drop_seed, drop_step; int64_t drop_row0;     // keep == 0: no dropout
    int act_kind;                            // DFM_ACT_*
    float* hidden_out;                       // [B, hid_stride]: layer i at column offset sum_{j<i} H[j]
    float* dnn_logit;                        // [B]
};

// one warp per sample; activations ping-pong through shared memory, lanes own output units
static __global__ void __launch_bounds__(128) summary_tower_kernel(SummaryTowerArgs a) {
    extern __shared__ float sm_sum[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* cur = sm_sum + (size_t)warp * 2 * a.maxdim;
    float* nxt = cur + a.maxdim;
    for (int b = blockIdx.x * 4 + warp; b < a.B; b += gridDim.x * 4) {
        for (int j = lane; j < a.dK; j += 32) cur[j] = a.h0[(size_t)b * a.dK + j];
        __syncwarp();
        int in = a.dK, col = 0;
        float* x = cur; float* y = nxt;
        for (int l = 0; l < a.L; ++l) {
            const int out = a.H[l];
            const float* W = a.dw + a.off_W[l];
            const float* bb = a.dw + a.off_b[l];
            const uint64_t key = a.drop_keep > 0.f ? dfm_drop_key(a.drop_seed, a.drop_step, (uint64_t)l) : 0;
            for (int o = lane; o < out; o += 32) {
                float v = 0.f;
                for (int j = 0; j < in; ++j) v = fmaf(x[j], __ldg(W + (size_t)j * out + o), v);
                v = act_fwd(a.act_kind, v + bb[o]);
                if (a.drop_keep > 0.f) v *= dfm_drop(key, (uint64_t)(a.drop_row0 + b) * out + o, a.drop_keep, a.drop_inv);
                y[o] = v;
                a.hidden_out[(size_t)b * a.hid_stride + col + o] = v;
            }
            __syncwarp();
            float* t = x; x = y; y = t;
            in = out; col += out;
        }
        float z = 0.f;
        for (int j = lane; j < in; j += 32) z = fmaf(x[j], a.dw[a.off_Wo + j], z);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
        if (lane == 0) a.dnn_logit[b] = z + a.dw[a.off_bo];
        __syncwarp();
    }
}

struct SummaryStats {            // device-side accumulator of one tensor
    double sum, sum_sq;
    unsigned long long num, zeros;
    float min, max;
};

// strided view of a tensor: element (r, c) at base[r * stride + c], r < rows, c < cols
static __global__ void __launch_bounds__(256) summary_reduce_kernel(const float* __restrict__ base, int64_t rows, int cols, int64_t stride,
                                                                     const double* __restrict__ limits, int n_limits,
                                                                     SummaryStats* __restrict__ st, unsigned long long* __restrict__ buckets) {
    double s = 0.0, s2 = 0.0;
    unsigned long long zeros = 0;
    float mn = INFINITY, mx = -INFINITY;
    const int64_t n = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const int64_t r = i / cols;
        const float v = base[r * stride + (i - r * cols)];
        s += (double)v; s2 += (double)v * (double)v;
        zeros += v == 0.f;
        mn = fminf(mn, v); mx = fmaxf(mx, v);
        int lo = 0, hi = n_limits;                 // upper_bound: first limit > v
        const double dv = (double)v;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (limits[mid] > dv) hi = mid; else lo = mid + 1; }
        atomicAdd(buckets + min(lo, n_limits - 1), 1ull);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        zeros += __shfl_xor_sync(0xffffffffu, zeros, o);
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&st->sum, s); atomicAdd(&st->sum_sq, s2);
        atomicAdd(&st->zeros, zeros);
        if (blockIdx.x == 0 && threadIdx.x == 0) st->num = (unsigned long long)n;
        // float min / max through the ordered-int trick
        auto enc = [](float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; };
        atomicMin(reinterpret_cast<int*>(&st->min), enc(mn));
        atomicMax(reinterpret_cast<int*>(&st->max), enc(mx));
    }
}
