// Graph preparation kernels: dense adjacency -> CSR, and the degree /
// normaliser vectors with scipy's csgraph.laplacian(normed=True) semantics
// (reference: calibration/WATS.py:24-27, :58-59, :99).  HBM-bound streaming
// kernels; no Laplacian is ever materialised.
#pragma once

#include "common.cuh"

namespace egnn {

// ---------------------------------------------------------------------------
// dense [n, n] float32 -> CSR.  One warp per row, 128-bit loads when the row
// stride allows; pass 1 counts, a single-CTA scan builds rowptr, pass 2 fills
// in column order with ballot compaction.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dense_count_kernel(const float* __restrict__ adj, int64_t n, int64_t ld,
                   int32_t* __restrict__ counts /* rowptr + 1 */, int32_t* __restrict__ nonbinary) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const bool vec_ok = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(adj) & 15) == 0);
    for (int64_t row = warp; row < n; row += nwarps) {
        const float* r = adj + row * ld;
        int cnt = 0;
        int nb = 0;
        if (vec_ok) {
            const int64_t n4 = n >> 2;
            const float4* r4 = reinterpret_cast<const float4*>(r);
            for (int64_t i = lane; i < n4; i += 32) {
                float4 v = __ldg(r4 + i);
                cnt += (v.x != 0.f) + (v.y != 0.f) + (v.z != 0.f) + (v.w != 0.f);
                nb |= (v.x != 0.f && v.x != 1.f) | (v.y != 0.f && v.y != 1.f) |
                      (v.z != 0.f && v.z != 1.f) | (v.w != 0.f && v.w != 1.f);
            }
            for (int64_t i = (n4 << 2) + lane; i < n; i += 32) {
                float v = __ldg(r + i);
                cnt += (v != 0.f);
                nb |= (v != 0.f && v != 1.f);
            }
        } else {
            for (int64_t i = lane; i < n; i += 32) {
                float v = __ldg(r + i);
                cnt += (v != 0.f);
                nb |= (v != 0.f && v != 1.f);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            nb |= __shfl_xor_sync(0xffffffffu, nb, o);
        }
        if (lane == 0) {
            counts[row] = cnt;
            if (nb) atomicOr(nonbinary, 1);
        }
    }
}

// in-place exclusive scan of rowptr[1..n] (holding counts) -> rowptr[0..n].
// One CTA of 1024 threads; n is at most a few 1e5 on this path.
__global__ void __launch_bounds__(1024)
rowptr_scan_kernel(int32_t* __restrict__ rowptr, int64_t n) {
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) { carry_s = 0; rowptr[0] = 0; }
    __syncthreads();
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t i = base + tid;
        int v = (i < n) ? rowptr[i + 1] : 0;
        int s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane == 31) warp_tot[wid] = s;
        __syncthreads();
        if (wid == 0) {
            int t = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int u = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t += u;
            }
            warp_tot[lane] = t;
        }
        __syncthreads();
        const int carry = carry_s;
        const int incl = s + (wid ? warp_tot[wid - 1] : 0) + carry;
        if (i < n) rowptr[i + 1] = incl;
        __syncthreads();
        if (tid == 1023) carry_s = incl;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
dense_fill_kernel(const float* __restrict__ adj, int64_t n, int64_t ld,
                  const int32_t* __restrict__ rowptr, int32_t* __restrict__ colidx,
                  float* __restrict__ vals) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const unsigned below = (1u << lane) - 1u;
    for (int64_t row = warp; row < n; row += nwarps) {
        const float* r = adj + row * ld;
        int64_t out = rowptr[row];
        // 4 independent 128-byte row segments in flight per iteration
        for (int64_t base = 0; base < n; base += 128) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t c = base + u * 32 + lane;
                v[u] = (c < n) ? __ldg(r + c) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned m = __ballot_sync(0xffffffffu, v[u] != 0.f);
                if (v[u] != 0.f) {
                    const int64_t pos = out + __popc(m & below);
                    colidx[pos] = (int32_t)(base + u * 32 + lane);
                    if (vals) vals[pos] = v[u];
                }
                out += __popc(m);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// degree pass: rowsum (self loops included), diagonal, and the in-degree
// (column sum) scattered with fp64 atomics so the result does not depend on
// the order the atomics retire in (binary graphs: exact integers).
// ---------------------------------------------------------------------------
template <bool HAS_VALS>
__global__ void __launch_bounds__(256)
degree_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
              const float* __restrict__ vals, int64_t n,
              float* __restrict__ rowsum, float* __restrict__ diag, double* __restrict__ colsum,
              int32_t* __restrict__ unsorted_flag, int64_t row0) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp; row < n; row += nwarps) {
        const int s = rowptr[row], e = rowptr[row + 1];
        double rs = 0.0;
        float dg = 0.f;
        int bad = 0;
        for (int p = s + lane; p < e; p += 32) {
            const int c = ld_stream_i32(colidx + p);
            if (unsorted_flag && p + 1 < e) bad |= (__ldg(colidx + p + 1) <= c);   // strictly increasing columns?
            const float v = HAS_VALS ? ld_stream_f32(vals + p) : 1.f;
            rs += (double)v;
            if (c == row + row0) dg += v;
            atomicAdd(colsum + c, (double)v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            rs += __shfl_xor_sync(0xffffffffu, rs, o);
            dg += __shfl_xor_sync(0xffffffffu, dg, o);
        }
        if (bad) atomicOr(unsorted_flag, 1);
        if (lane == 0) {
            rowsum[row] = (float)rs;
            diag[row] = dg;
        }
    }
}

__device__ __forceinline__ void normaliser_from_w(float w, float& dinv, uint8_t& iso) {
    iso = (w == 0.f) ? 1 : 0;
    dinv = iso ? 1.f : (float)(1.0 / sqrt((double)w));
}

__global__ void __launch_bounds__(256)
normaliser_kernel(const double* __restrict__ colsum, const float* __restrict__ diag,
                  const float* __restrict__ rowsum, int64_t n, float* __restrict__ dinv,
                  uint8_t* __restrict__ iso, float* __restrict__ x0, float* __restrict__ w_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // scipy: w = colsum - diag in the adjacency dtype (float32)
    const float w = (float)colsum[i] - diag[i];
    float d;
    uint8_t is;
    normaliser_from_w(w, d, is);
    dinv[i] = d;
    iso[i] = is;
    if (x0 && rowsum) x0[i] = (float)log1p((double)rowsum[i]);
    if (w_out) w_out[i] = w;
}

__global__ void __launch_bounds__(256)
logdeg_kernel(const float* __restrict__ rowsum, int64_t n, float* __restrict__ x0) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x0[i] = (float)log1p((double)rowsum[i]);
}

struct DeltaList {
    int32_t n;
    int32_t row[EGNN_MAX_DELTA];
    int32_t col[EGNN_MAX_DELTA];
    float val[EGNN_MAX_DELTA];
};

// UGCA recompute: re-derive dinv/iso/x0 of the <= 2*budget nodes an edge flip
// touches (the *_out vectors already hold a copy of the base graph's).
// Row-sharded: w/dinv/iso are full-length (replicated), rowsum/x0 hold the rows
// [row_begin, row_begin + n_rows) only.
__global__ void patch_degrees_kernel(const float* __restrict__ w_base,
                                     const float* __restrict__ rowsum_base, DeltaList d,
                                     float* __restrict__ dinv_out, uint8_t* __restrict__ iso_out,
                                     float* __restrict__ x0_out, int64_t row_begin, int64_t n_rows) {
    const int e = threadIdx.x;
    if (e >= d.n) return;
    // in-degree of node col[e]: handled by the first delta naming that column
    {
        const int u = d.col[e];
        bool first = true;
        for (int j = 0; j < e; ++j) first &= (d.col[j] != u);
        if (first) {
            float dw = 0.f;
            for (int j = 0; j < d.n; ++j)
                if (d.col[j] == u && d.row[j] != u) dw += d.val[j];
            float dv;
            uint8_t is;
            normaliser_from_w(w_base[u] + dw, dv, is);
            dinv_out[u] = dv;
            iso_out[u] = is;
        }
    }
    // row sum of node row[e] (self loops count)
    {
        const int u = d.row[e];
        bool first = true;
        for (int j = 0; j < e; ++j) first &= (d.row[j] != u);
        if (first && u >= row_begin && u < row_begin + n_rows) {
            float dr = 0.f;
            for (int j = 0; j < d.n; ++j)
                if (d.row[j] == u) dr += d.val[j];
            x0_out[u - row_begin] = (float)log1p((double)(rowsum_base[u - row_begin] + dr));
        }
    }
}

// UGCA recompute without the three vector copies: the scratch vectors dinv_io / iso_io / x0_io / y0_io hold
// copies of the base graph's (made once per graph); mode 0 writes the patched values of every node a flip
// touches (y0 = dinv * x0 included, so the step kernel needs no prologue phase), mode 1 puts the base
// values back after the pass has been queued.  One thread per candidate node (rows, then columns of the
// flips); the first occurrence of a node does the work.
__global__ void __launch_bounds__(2 * EGNN_MAX_DELTA)
patch_nodes_kernel(const float* __restrict__ w_base, const float* __restrict__ rowsum_base,
                   const float* __restrict__ dinv_base, const uint8_t* __restrict__ iso_base,
                   const float* __restrict__ x0_base, const float* __restrict__ y0_base, DeltaList d,
                   float* __restrict__ dinv_io, uint8_t* __restrict__ iso_io, float* __restrict__ x0_io,
                   float* __restrict__ y0_io, int mode) {
    const int t = threadIdx.x;
    if (t >= 2 * d.n) return;
    const int u = t < d.n ? d.row[t] : d.col[t - d.n];
    for (int j = 0; j < t; ++j)
        if ((j < d.n ? d.row[j] : d.col[j - d.n]) == u) return;
    if (mode == 1) {
        dinv_io[u] = dinv_base[u]; iso_io[u] = iso_base[u]; x0_io[u] = x0_base[u]; y0_io[u] = y0_base[u];
        return;
    }
    float dw = 0.f, dr = 0.f;
    for (int j = 0; j < d.n; ++j) {
        if (d.col[j] == u && d.row[j] != u) dw += d.val[j];     // in-degree (self loops excluded)
        if (d.row[j] == u) dr += d.val[j];                      // row sum (self loops count)
    }
    float dv;
    uint8_t is;
    normaliser_from_w(w_base[u] + dw, dv, is);
    const float x = (float)log1p((double)(rowsum_base[u] + dr));
    dinv_io[u] = dv; iso_io[u] = is; x0_io[u] = x; y0_io[u] = dv * x;
}

}  // namespace egnn
