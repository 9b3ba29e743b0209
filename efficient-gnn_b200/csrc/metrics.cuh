// Calibration metrics on the device: accuracy, mean max-probability and the
// class-wise expected calibration error of utils/ece.py:8-89 as evaluated by
// benchmark_calibration_methods.py:100-127 - per class a 10-bin one-vs-rest ECE with
// right-closed bins (numpy.digitize(p, linspace(0,1,n_bins+1), right=True) - 1), bins holding
// fewer than 4 samples skipped, then the mean over classes.
// Pass 1: one warp per selected sample bins its C probabilities (float64 atomics: the sums do
// not depend on arrival order beyond the last bit).  Pass 2: one thread per class folds the bins.
#pragma once

#include "common.cuh"

namespace egnn {

constexpr int kEceMaxBins = 32;

struct EceWs {                    // laid out in the caller's workspace
    double* sum_p;                // [C, n_bins]
    unsigned long long* count;    // [C, n_bins]
    unsigned long long* hits;     // [C, n_bins]
    double* scalars;              // [4]: selected samples, correct, sum of max prob, (unused)
};

__device__ __forceinline__ int ece_bin(double p, int n_bins) {
    // number of edges strictly below p, minus one; edge i = i * (1 / n_bins), last edge exactly 1
    const double step = 1.0 / (double)n_bins;
    int below = 0;
    for (int i = 0; i <= n_bins; ++i) {
        const double edge = (i == n_bins) ? 1.0 : (double)i * step;
        below += (edge < p);
    }
    return below - 1;
}

__global__ void __launch_bounds__(256)
ece_accumulate_kernel(const float* __restrict__ x, int is_log, const int64_t* __restrict__ labels,
                      const uint8_t* __restrict__ mask, int64_t n, int C, int n_bins, EceWs ws) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += nwarps) {
        if (mask && !mask[i]) continue;
        const int y = (int)labels[i];
        float best = -INFINITY;
        int arg = 0;
        for (int c = lane; c < C; c += 32) {
            const float v = x[i * C + c];
            const float p = is_log ? expf(v) : v;
            if (p > best) { best = p; arg = c; }
            const int b = ece_bin((double)p, n_bins);
            if (b >= 0 && b < n_bins) {
                atomicAdd(ws.sum_p + c * n_bins + b, (double)p);
                atomicAdd(ws.count + c * n_bins + b, 1ull);
                if (c == y) atomicAdd(ws.hits + c * n_bins + b, 1ull);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {            // arg max, first index on ties (numpy argmax)
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
        }
        if (lane == 0) {
            atomicAdd(ws.scalars + 0, 1.0);
            if (arg == y) atomicAdd(ws.scalars + 1, 1.0);
            atomicAdd(ws.scalars + 2, (double)best);
        }
    }
}

__global__ void ece_finalize_kernel(int C, int n_bins, EceWs ws, double* __restrict__ out3) {
    __shared__ double per_class[1024];
    const double m = ws.scalars[0];
    double acc = 0.0;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double e = 0.0;
        for (int b = 0; b < n_bins; ++b) {
            const unsigned long long cnt = ws.count[c * n_bins + b];
            if (cnt < 4ull) continue;
            const double conf = ws.sum_p[c * n_bins + b] / (double)cnt;
            const double hit = (double)ws.hits[c * n_bins + b] / (double)cnt;
            e += fabs(conf - hit) * ((double)cnt / m);
        }
        acc += e;
    }
    per_class[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double total = 0.0;
        for (int t = 0; t < (int)blockDim.x; ++t) total += per_class[t];     // fixed order
        out3[0] = m > 0 ? ws.scalars[1] / m : 0.0;
        out3[1] = m > 0 ? ws.scalars[2] / m : 0.0;
        out3[2] = m > 0 ? total / (double)C : 0.0;
    }
}

}  // namespace egnn
