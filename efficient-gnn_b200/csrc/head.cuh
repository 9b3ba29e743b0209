// Fused temperature head of WATS.forward (calibration/WATS.py:122-130), inference only:
//   t_i    = w2 . relu(W1 h_i + b1) + b2          (net = Linear(F,H) - ReLU - Linear(H,1))
//   T_i    = log(exp(t_i) + 1.1)
//   out_i  = log_softmax(logits_i / T_i)
// One warp per node: lanes split the hidden units for the MLP and the classes for the
// softmax (shuffle reductions).  HBM-bound: reads logits [N,C] + features [N,F], writes [N,C].
#pragma once

#include "common.cuh"

namespace egnn {

constexpr int kHeadMaxHidden = 64;
constexpr int kHeadMaxFeat = 64;

__global__ void __launch_bounds__(256)
temperature_head_kernel(const float* __restrict__ feats, const float* __restrict__ w1, const float* __restrict__ b1,
                        const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ logits,
                        float* __restrict__ out, float* __restrict__ temps_out, int64_t n, int F, int H, int C) {
    __shared__ float s_w1[kHeadMaxHidden * kHeadMaxFeat];
    __shared__ float s_b1[kHeadMaxHidden], s_w2[kHeadMaxHidden];
    for (int i = threadIdx.x; i < H * F; i += blockDim.x) s_w1[i] = w1[i];
    for (int i = threadIdx.x; i < H; i += blockDim.x) { s_b1[i] = b1[i]; s_w2[i] = w2[i]; }
    __syncthreads();
    const float bias2 = __ldg(b2);
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += nwarps) {
        const float* h = feats + i * F;
        float part = 0.f;
        for (int u = lane; u < H; u += 32) {
            float a = s_b1[u];
            for (int f = 0; f < F; ++f) a = fmaf(s_w1[u * F + f], __ldg(h + f), a);
            part = fmaf(s_w2[u], fmaxf(a, 0.f), part);
        }
        const float t = warp_sum(part) + bias2;
        const float temp = logf(expf(t) + 1.1f);
        if (temps_out && lane == 0) temps_out[i] = temp;
        const float* z = logits + i * C;
        float m = -INFINITY;
        for (int c = lane; c < C; c += 32) m = fmaxf(m, z[c] / temp);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float se = 0.f;
        for (int c = lane; c < C; c += 32) se += expf(z[c] / temp - m);
        se = warp_sum(se);
        const float lse = m + logf(se);
        for (int c = lane; c < C; c += 32) out[i * C + c] = z[c] / temp - lse;
    }
}

}  // namespace egnn
