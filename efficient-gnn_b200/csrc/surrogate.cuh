// Sparse structure-gradient surrogate.
//
// The UGCA attack (calib_attack/calib_fga.py:864-890) runs the calibrated
// surrogate on a DENSE [N,N] adjacency and back-propagates to it, although it
// only reads row and column `target_node` of the gradient (:881).  The base
// model is the two-layer row-normalised GCN of src/gnn/model.py:43-52:
//     A_n = D^-1 A (deg 0 -> 1),  Z1 = A_n X W1^T + b1,  H1 = relu(Z1),
//     logits = (A_n H1) W2^T + b2.
// With XW = X W1^T computed once (X and W1 do not change during an attack),
// everything the attack needs about node t follows from ONE sparse product
// Z1 = A_n XW + b1 (CSR, hidden width H <= 128) and O(N H) dot products:
//     v      = u W2                      (u = dLoss/dlogits[t], from torch on a [1,C] tensor)
//     q_i    = A_n[t,i] (v . relu'(Z1[i]))            for i in row t
//     dL/dA[t,m] = (v . relu(Z1[m]) + q_t . XW[m] - s_t) / deg_t,
//                  s_t = v . (A_n H1)[t] + q_t . (Z1[t] - b1)
//     dL/dA[i,t] = (q_i . XW[t] - q_i . (Z1[i] - b1)) / deg_i      for i in row t, 0 elsewhere
// (rows whose degree was clamped to 1 take no gradient through the degree, as
// the in-place `deg[deg == 0] = 1` of the reference does).  Edge flips of the
// running attack ride on top of the CSR as a delta list, like the wavelet path.
#pragma once

#include "common.cuh"
#include "prep.cuh"

namespace egnn {

constexpr int kGcnMaxHidden = 128;
constexpr int kGcnCtxFloats = kGcnMaxHidden + 8;   // h2[H], then deg_t, a_tt (raw, before normalisation), clamped flag

// Y[i,:] = (1/deg_i) (sum_j a_ij M[j,:] + sum_{flips in row i} dv M[dc,:]) + bias, deg_i = rowsum_i (+ flips; 0 -> 1).
// One warp per row; a lane owns float4 chunks lane, lane + 32, ... of the H-wide row (H % 4 == 0, H <= 128:
// one chunk per lane), entries of the row are taken kGcnUnroll at a time.
constexpr int kGcnUnroll = 4;

template <bool HAS_VALS>
__global__ void __launch_bounds__(256)
gcn_propagate_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                     const float* __restrict__ vals, const float* __restrict__ m, const float* __restrict__ bias,
                     float* __restrict__ y, float* __restrict__ deg_out, int64_t n, int32_t h, const DeltaList delta) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    const int chunks = h >> 2;
    const bool mine = lane < chunks;
    const int start = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float deg = 0.f;
    for (int b0 = start; b0 < end; b0 += 32) {
        const int cnt = min(32, end - b0);
        int cj = 0;
        float wj = 0.f;
        if (lane < cnt) {
            cj = ld_stream_i32(colidx + b0 + lane);
            wj = HAS_VALS ? ld_stream_f32(vals + b0 + lane) : 1.f;
        }
        deg += wj;
        for (int e0 = 0; e0 < cnt; e0 += kGcnUnroll) {
            float4 x[kGcnUnroll];
            float w[kGcnUnroll];
#pragma unroll
            for (int u = 0; u < kGcnUnroll; ++u) {
                const int c = __shfl_sync(0xffffffffu, cj, (e0 + u) & 31);
                w[u] = (e0 + u < cnt) ? __shfl_sync(0xffffffffu, wj, (e0 + u) & 31) : 0.f;
                x[u] = (mine && e0 + u < cnt) ? __ldg(reinterpret_cast<const float4*>(m + (int64_t)c * h) + lane)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < kGcnUnroll; ++u) {
                acc.x = fmaf(w[u], x[u].x, acc.x); acc.y = fmaf(w[u], x[u].y, acc.y);
                acc.z = fmaf(w[u], x[u].z, acc.z); acc.w = fmaf(w[u], x[u].w, acc.w);
            }
        }
    }
    deg = warp_sum(deg);
    for (int d = 0; d < delta.n; ++d) {
        if (delta.row[d] == row) {
            const float w = delta.val[d];
            deg += w;
            if (mine) {
                const float4 x = __ldg(reinterpret_cast<const float4*>(m + (int64_t)delta.col[d] * h) + lane);
                acc.x = fmaf(w, x.x, acc.x); acc.y = fmaf(w, x.y, acc.y);
                acc.z = fmaf(w, x.z, acc.z); acc.w = fmaf(w, x.w, acc.w);
            }
        }
    }
    const float dclamp = deg == 0.f ? 1.f : deg;
    if (lane == 0 && deg_out) deg_out[row] = deg;             // raw row sum; consumers clamp
    if (mine) {
        const float inv = 1.f / dclamp;
        const float4 b = bias ? __ldg(reinterpret_cast<const float4*>(bias) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        reinterpret_cast<float4*>(y + row * h)[lane] =
            make_float4(fmaf(acc.x, inv, b.x), fmaf(acc.y, inv, b.y), fmaf(acc.z, inv, b.z), fmaf(acc.w, inv, b.w));
    }
}

// logits[t,:] = ((A_n relu(Z1))[t,:]) W2^T + b2 for ONE node t: one CTA, warps stride the entries of row t
// (+ flips), lanes the hidden columns.  ctx: h2[H] = (A_n relu(Z1))[t,:], deg_t (raw), a_tt (raw), for the backward.
template <bool HAS_VALS>
__global__ void __launch_bounds__(256)
gcn_target_logits_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                         const float* __restrict__ vals, const float* __restrict__ z1, const float* __restrict__ w2,
                         const float* __restrict__ b2, int32_t t, int32_t h, int32_t n_classes,
                         float* __restrict__ logits, float* __restrict__ ctx, const DeltaList delta) {
    __shared__ float part[8][kGcnMaxHidden];
    __shared__ float sdeg[8], satt[8];
    __shared__ float h2[kGcnMaxHidden];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int start = __ldg(rowptr + t), end = __ldg(rowptr + t + 1);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};                     // columns lane, lane + 32, lane + 64, lane + 96
    float deg = 0.f, att = 0.f;
    auto add_entry = [&](int c, float w) {
        deg += w;
        if (c == t) att += w;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int col = lane + 32 * v;
            if (col < h) acc[v] = fmaf(w, fmaxf(__ldg(z1 + (int64_t)c * h + col), 0.f), acc[v]);
        }
    };
    for (int q = start + wid; q < end; q += 8) add_entry(__ldg(colidx + q), HAS_VALS ? __ldg(vals + q) : 1.f);
    if (wid == 0)
        for (int d = 0; d < delta.n; ++d)
            if (delta.row[d] == t) add_entry(delta.col[d], delta.val[d]);
#pragma unroll
    for (int v = 0; v < 4; ++v) part[wid][lane + 32 * v] = acc[v];
    if (lane == 0) { sdeg[wid] = deg; satt[wid] = att; }
    __syncthreads();
    float degt = 0.f, attt = 0.f;
    for (int w = 0; w < 8; ++w) { degt += sdeg[w]; attt += satt[w]; }      // every lane adds the same values (uniform per warp)
    const float inv = 1.f / (degt == 0.f ? 1.f : degt);
    for (int col = threadIdx.x; col < h; col += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += part[w][col];
        h2[col] = s * inv;
        ctx[col] = s * inv;
    }
    if (threadIdx.x == 0) { ctx[kGcnMaxHidden] = degt; ctx[kGcnMaxHidden + 1] = attt; }
    __syncthreads();
    for (int c = wid; c < n_classes; c += 8) {
        float s = 0.f;
        for (int col = lane; col < h; col += 32) s = fmaf(__ldg(w2 + (int64_t)c * h + col), h2[col], s);
        s = warp_sum(s);
        if (lane == 0) logits[c] = s + __ldg(b2 + c);
    }
}

// Shared prologue of the two gradient kernels: v = u W2, q_t, s_t in shared memory.
struct GcnGradShared {
    float v[kGcnMaxHidden];      // dLoss/d(A_n H1)[t,:]
    float qt[kGcnMaxHidden];     // a_tt (v . relu'(Z1[t]))
    float st, inv_deg_t, clamped_t;
};

__device__ __forceinline__ void gcn_grad_prologue(GcnGradShared& sh, const float* __restrict__ u, const float* __restrict__ w2,
                                                  const float* __restrict__ z1, const float* __restrict__ b1,
                                                  const float* __restrict__ ctx, int32_t t, int32_t h, int32_t n_classes) {
    const float deg_raw = ctx[kGcnMaxHidden], att_raw = ctx[kGcnMaxHidden + 1];
    const bool clamped = deg_raw == 0.f;
    const float inv = 1.f / (clamped ? 1.f : deg_raw);
    for (int col = threadIdx.x; col < h; col += blockDim.x) {
        float s = 0.f;
        for (int c = 0; c < n_classes; ++c) s = fmaf(__ldg(u + c), __ldg(w2 + (int64_t)c * h + col), s);
        sh.v[col] = s;
        sh.qt[col] = att_raw * inv * (__ldg(z1 + (int64_t)t * h + col) > 0.f ? s : 0.f);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        float s = 0.f;
        for (int col = threadIdx.x; col < h; col += 32)
            s += sh.v[col] * ctx[col] + sh.qt[col] * (__ldg(z1 + (int64_t)t * h + col) - __ldg(b1 + col));
        s = warp_sum(s);
        if (threadIdx.x == 0) { sh.st = clamped ? 0.f : s; sh.inv_deg_t = inv; sh.clamped_t = clamped ? 1.f : 0.f; }
    }
    __syncthreads();
}

// grad_row[m] = dLoss/dA[t,m] for every node m: 8 lanes per node, two H-wide dot products each.
__global__ void __launch_bounds__(256)
gcn_grad_row_kernel(const float* __restrict__ u, const float* __restrict__ w2, const float* __restrict__ z1,
                    const float* __restrict__ xw, const float* __restrict__ b1, const float* __restrict__ ctx,
                    int32_t t, int64_t n, int32_t h, int32_t n_classes, float* __restrict__ grad_row) {
    __shared__ GcnGradShared sh;
    gcn_grad_prologue(sh, u, w2, z1, b1, ctx, t, h, n_classes);
    const int sub = threadIdx.x & 7;
    const int chunks = h >> 2;
    for (int64_t m = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; m < n;
         m += ((int64_t)gridDim.x * blockDim.x) >> 3) {
        float s = 0.f;
        for (int ch = sub; ch < chunks; ch += 8) {
            const float4 z = __ldg(reinterpret_cast<const float4*>(z1 + m * h) + ch);
            const float4 x = __ldg(reinterpret_cast<const float4*>(xw + m * h) + ch);
            const float* v = sh.v + 4 * ch;
            const float* q = sh.qt + 4 * ch;
            s += v[0] * fmaxf(z.x, 0.f) + v[1] * fmaxf(z.y, 0.f) + v[2] * fmaxf(z.z, 0.f) + v[3] * fmaxf(z.w, 0.f);
            s += q[0] * x.x + q[1] * x.y + q[2] * x.z + q[3] * x.w;
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (sub == 0) grad_row[m] = (s - sh.st) * sh.inv_deg_t;
    }
}

// grad_col[i] = dLoss/dA[i,t]: non-zero only for i in row t of the (flipped) adjacency and i = t.  The caller
// zero-fills grad_col; one CTA walks row t (+ flips), one warp per entry; contributions of an entry are linear in
// its weight, so CSR entries and flips of the same column simply add (atomicAdd: a column can occur in both).
template <bool HAS_VALS>
__global__ void __launch_bounds__(256)
gcn_grad_col_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                    const float* __restrict__ vals, const float* __restrict__ u, const float* __restrict__ w2,
                    const float* __restrict__ z1, const float* __restrict__ xw, const float* __restrict__ b1,
                    const float* __restrict__ deg, const float* __restrict__ ctx, const float* __restrict__ grad_row,
                    int32_t t, int32_t h, int32_t n_classes, float* __restrict__ grad_col, const DeltaList delta) {
    __shared__ GcnGradShared sh;
    gcn_grad_prologue(sh, u, w2, z1, b1, ctx, t, h, n_classes);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    auto entry = [&](int i, float w) {
        if (i == t) return;                                   // the diagonal entry is the row formula at m = t
        const float a_ti = w * sh.inv_deg_t;
        float g = 0.f;
        for (int col = lane; col < h; col += 32) {
            const float z = __ldg(z1 + (int64_t)i * h + col);
            const float q = z > 0.f ? a_ti * sh.v[col] : 0.f;
            const float di = __ldg(deg + i);
            const float own = di == 0.f ? 0.f : (z - __ldg(b1 + col));        // clamped row: no gradient through its degree
            g += q * (__ldg(xw + (int64_t)t * h + col) - own);
        }
        g = warp_sum(g);
        if (lane == 0) {
            const float di = __ldg(deg + i);
            atomicAdd(grad_col + i, g / (di == 0.f ? 1.f : di));
        }
    };
    const int start = __ldg(rowptr + t), end = __ldg(rowptr + t + 1);
    for (int q = start + wid; q < end; q += 8) entry(__ldg(colidx + q), HAS_VALS ? __ldg(vals + q) : 1.f);
    if (wid == 0)
        for (int d = 0; d < delta.n; ++d)
            if (delta.row[d] == t) entry(delta.col[d], delta.val[d]);
    if (threadIdx.x == 0) grad_col[t] = grad_row[t];          // written by the row kernel launched before this one
}

}  // namespace egnn
