// Column-blocked sliced-ELL ("SELL-32, 16-bit local indices") path for the
// narrow (F = 1) Chebyshev orders - the reference's default configuration
// (X0 = log1p(degree), calibration/WATS.py:58-59) on graphs whose CSR streams
// from HBM (Reddit shape: 114.6 M entries, 459 MB of int32 indices per order).
//
// Why: with F = 1 every stored entry costs one 4-byte random gather of
// T_{k-1}[j].  Through L1/L2 that gather rate (~1.5 per clock per SM, one
// 128-byte line per wavefront) is 8x short of what HBM can stream.  Shared
// memory serves ~9 random 4-byte reads per clock per SM, so the operand vector
// is staged in shared memory one COLUMN BLOCK (<= 49152 floats) at a time and
// the matrix is re-laid-out once per graph so that
//   * entries are grouped by column block and carry 16-bit block-local column
//     indices (halves the index stream: 2 B per entry),
//   * rows are cut into "virtual rows" of <= lmax entries (hub rows split, so
//     power-law graphs balance), sorted by length and packed 32 to a slice in
//     lane-interleaved order: lane r of a warp walks virtual row r with 16-byte
//     loads (8 indices), the warp's loads are one contiguous 512-byte segment,
//   * stored self loops and padding point at a zero slot, so the inner loop has
//     no predicate: load 8 indices -> 8 LDS -> 8 FADD.
// Per order: sell_spmv_kernel (persistent, one CTA per SM, entries split evenly
// over CTAs) writes one partial sum per virtual row; sell_epilogue_kernel adds
// each row's partials in a fixed order (deterministic), applies the Laplacian
// scaling, the three-term recurrence and the scale accumulation.
#pragma once

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "cheb.cuh"
#include "common.cuh"

namespace egnn {

constexpr int kSellThreads = 1024;
constexpr int kSellSliceRows = 32;
constexpr int kSellGroup = 8;           // indices per 16-byte load
constexpr int kSellMaxBlocks = 64;

// totals the host needs after the prepare pass (device int64[8])
enum SellTotals { kTotU = 0, kTotV = 1, kTotSlices = 2, kTotEntries = 3, kTotRowV = 4 };

// ---- build pass 1: per (column block, row) segment of the sorted CSR row ----
__global__ void __launch_bounds__(256)
sell_count_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, int n, int C,
                  int CB, int lmax, int32_t* __restrict__ seg_start, int32_t* __restrict__ seg_cnt,
                  int32_t* __restrict__ nv, int32_t* __restrict__ nvrow) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int e = rowptr[i + 1];
    int prev = rowptr[i];
    int total = 0;
    for (int c = 0; c < C; ++c) {
        int pos = e;
        if (c + 1 < C) {                       // first position with column >= (c+1)*CB
            const int bound = (c + 1) * CB;
            int lo = prev, hi = e;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (__ldg(colidx + mid) < bound) lo = mid + 1; else hi = mid;
            }
            pos = lo;
        }
        const int cnt = pos - prev;
        const int v = (cnt + lmax - 1) / lmax;
        seg_start[(size_t)c * n + i] = prev;
        seg_cnt[(size_t)c * n + i] = cnt;
        nv[(size_t)c * n + i] = v;
        total += v;
        prev = pos;
    }
    nvrow[i] = total;
}

// ---- build pass 2: one record per virtual row -------------------------------
__global__ void __launch_bounds__(256)
sell_emit_kernel(int n, int C, int lmax, const int32_t* __restrict__ seg_start,
                 const int32_t* __restrict__ seg_cnt, const int32_t* __restrict__ nv,
                 const int32_t* __restrict__ u_off, const int32_t* __restrict__ rv_ptr,
                 uint32_t* __restrict__ key, int32_t* __restrict__ uval, int32_t* __restrict__ vsrc,
                 int32_t* __restrict__ vslot, int32_t* __restrict__ vrow) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)C * n) return;
    const int c = (int)(t / n), i = (int)(t % n);
    const int cnt = seg_cnt[t];
    if (cnt == 0) return;
    int slot = rv_ptr[i];
    for (int cc = 0; cc < c; ++cc) slot += nv[(size_t)cc * n + i];
    const int u0 = u_off[t];
    const int start = seg_start[t];
    const int chunks = nv[t];
    for (int j = 0; j < chunks; ++j) {
        const int len = min(lmax, cnt - j * lmax);
        const int u = u0 + j;
        key[u] = ((uint32_t)c << 16) | (uint32_t)(lmax - len);     // block-major, longest first
        uval[u] = u;
        vsrc[u] = start + j * lmax;
        vslot[u] = slot + j;
        vrow[u] = i;
    }
}

// ---- build pass 3: block pointers (padded to whole slices) ------------------
__global__ void sell_blocks_kernel(int n, int C, const int32_t* __restrict__ u_off,
                                   int32_t* __restrict__ q_ptr, int32_t* __restrict__ vp_ptr,
                                   int32_t* __restrict__ blk_slice_ptr, int64_t* __restrict__ totals) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int vp = 0;
    for (int c = 0; c < C; ++c) {
        const int q0 = u_off[(size_t)c * n];
        const int q1 = u_off[(size_t)(c + 1) * n];
        q_ptr[c] = q0;
        vp_ptr[c] = vp;
        blk_slice_ptr[c] = vp / kSellSliceRows;
        vp += (q1 - q0 + kSellSliceRows - 1) / kSellSliceRows * kSellSliceRows;
    }
    q_ptr[C] = u_off[(size_t)C * n];
    vp_ptr[C] = vp;
    blk_slice_ptr[C] = vp / kSellSliceRows;
    totals[kTotU] = q_ptr[C];
    totals[kTotV] = vp;
    totals[kTotSlices] = vp / kSellSliceRows;
}

__device__ __forceinline__ int sell_block_of_slice(const int32_t* blk_slice_ptr, int C, int s) {
    int c = 0;
    while (c + 1 < C && s >= blk_slice_ptr[c + 1]) ++c;
    return c;
}

// ---- build pass 4: padded size of every slice --------------------------------
__global__ void __launch_bounds__(256)
sell_slice_kernel(int C, int lmax, const uint32_t* __restrict__ key_sorted,
                  const int32_t* __restrict__ q_ptr, const int32_t* __restrict__ blk_slice_ptr,
                  const int64_t* __restrict__ totals, int32_t* __restrict__ slice_sz) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= (int)totals[kTotSlices]) return;
    const int c = sell_block_of_slice(blk_slice_ptr, C, s);
    const int rank0 = (s - blk_slice_ptr[c]) * kSellSliceRows;
    const int len = lmax - (int)(key_sorted[q_ptr[c] + rank0] & 0xffffu);     // longest row of the slice
    slice_sz[s] = (len + kSellGroup - 1) / kSellGroup * kSellGroup * kSellSliceRows;
}

__global__ void sell_totals_kernel(const int32_t* __restrict__ slice_off, const int32_t* __restrict__ rv_ptr,
                                   int n, int64_t* __restrict__ totals) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    totals[kTotEntries] = slice_off[totals[kTotSlices]];
    totals[kTotRowV] = rv_ptr[n];
}

// ---- build pass 5: write the lane-interleaved 16-bit index stream -----------
__global__ void __launch_bounds__(256)
sell_fill_kernel(const int32_t* __restrict__ colidx, int C, int CB, int lmax, int n_slices,
                 const uint32_t* __restrict__ key_sorted, const int32_t* __restrict__ perm,
                 const int32_t* __restrict__ vsrc, const int32_t* __restrict__ vslot,
                 const int32_t* __restrict__ vrow, const int32_t* __restrict__ q_ptr,
                 const int32_t* __restrict__ vp_ptr, const int32_t* __restrict__ blk_slice_ptr,
                 const int32_t* __restrict__ slice_off, uint16_t* __restrict__ idx,
                 int32_t* __restrict__ rv_idx, int row0) {
    const int lane = threadIdx.x & 31;
    const int s = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (s >= n_slices) return;
    const int c = sell_block_of_slice(blk_slice_ptr, C, s);
    const int v = s * kSellSliceRows + lane;
    const int rank = v - vp_ptr[c];
    const bool valid = rank < (q_ptr[c + 1] - q_ptr[c]);
    int len = 0, src = 0, row = -1;
    if (valid) {
        const int q = q_ptr[c] + rank;
        const int u = perm[q];
        len = lmax - (int)(key_sorted[q] & 0xffffu);
        src = vsrc[u];
        row = vrow[u] + row0;             // global id of the row (diagonal test)
        rv_idx[vslot[u]] = v;
    }
    const int off = slice_off[s];
    const int groups = (slice_off[s + 1] - off) / (kSellGroup * kSellSliceRows);
    const int col0 = c * CB;
    uint4* dst = reinterpret_cast<uint4*>(idx + off) + lane;
    for (int g = 0; g < groups; ++g) {
        uint32_t w[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            uint32_t pair = 0;
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int j = g * kSellGroup + h * 2 + t;
                uint32_t loc = (uint32_t)CB;                      // zero slot: padding and self loops
                if (j < len) {
                    const int col = __ldg(colidx + src + j);
                    if (col != row) loc = (uint32_t)(col - col0);
                }
                pair |= loc << (16 * t);
            }
            w[h] = pair;
        }
        dst[(size_t)g * kSellSliceRows] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// ---- hot kernel ---------------------------------------------------------------
// y: the gather operand dinv (.) T_{k-1}, [n] float32.  One partial sum per
// virtual row goes to vpart[V].  Dynamic shared memory: (CB + 1) floats.
__device__ __forceinline__ uint4 ld_stream_u32x4(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ float sell_gather8(const float* ysm, const uint4 q) {
    const float a0 = ysm[q.x & 0xffffu], a1 = ysm[q.x >> 16];
    const float a2 = ysm[q.y & 0xffffu], a3 = ysm[q.y >> 16];
    const float a4 = ysm[q.z & 0xffffu], a5 = ysm[q.z >> 16];
    const float a6 = ysm[q.w & 0xffffu], a7 = ysm[q.w >> 16];
    return ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

template <int UNROLL>
__global__ void __launch_bounds__(kSellThreads, 1)
sell_spmv_kernel(const uint16_t* __restrict__ idx, const int32_t* __restrict__ slice_off,
                 const int32_t* __restrict__ blk_slice_ptr, int C, int CB, int n_slices,
                 const float* __restrict__ y, int n, float* __restrict__ vpart) {
    extern __shared__ __align__(16) float ysm[];
    __shared__ int next_slice;
    __shared__ int range[2];
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        // even split of the padded entries over the CTAs, on slice boundaries
        const int64_t total = slice_off[n_slices];
        const int64_t lo_e = total * blockIdx.x / gridDim.x;
        const int64_t hi_e = total * (blockIdx.x + 1) / gridDim.x;
        int bounds[2];
        const int64_t want[2] = {lo_e, hi_e};
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            int lo = 0, hi = n_slices;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if ((int64_t)slice_off[mid] < want[t]) lo = mid + 1; else hi = mid;
            }
            bounds[t] = lo;
        }
        if (blockIdx.x == gridDim.x - 1) bounds[1] = n_slices;
        range[0] = bounds[0];
        range[1] = bounds[1];
    }
    __syncthreads();
    const int s_begin = range[0], s_end = range[1];
    if (s_begin >= s_end) return;

    for (int c = sell_block_of_slice(blk_slice_ptr, C, s_begin); c < C; ++c) {
        const int sub_begin = max(s_begin, blk_slice_ptr[c]);
        const int sub_end = min(s_end, blk_slice_ptr[c + 1]);
        if (sub_begin >= s_end) break;
        if (sub_begin >= sub_end) continue;
        __syncthreads();                                   // previous block's gathers are done
        // stage the column block of the operand in shared memory
        const int col0 = c * CB;
        const int cnt = min(CB, n - col0);
        const float* src = y + col0;
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            const int cnt4 = cnt >> 2;
            const float4* src4 = reinterpret_cast<const float4*>(src);
            float4* dst4 = reinterpret_cast<float4*>(ysm);
            for (int t = threadIdx.x; t < cnt4; t += kSellThreads) dst4[t] = __ldg(src4 + t);
            for (int t = (cnt4 << 2) + threadIdx.x; t < cnt; t += kSellThreads) ysm[t] = __ldg(src + t);
        } else {
            for (int t = threadIdx.x; t < cnt; t += kSellThreads) ysm[t] = __ldg(src + t);
        }
        for (int t = cnt + threadIdx.x; t <= CB; t += kSellThreads) ysm[t] = 0.f;     // includes the zero slot
        if (threadIdx.x == 0) next_slice = sub_begin;
        __syncthreads();

        while (true) {
            int s = 0;
            if (lane == 0) s = atomicAdd(&next_slice, 1);
            s = __shfl_sync(0xffffffffu, s, 0);
            if (s >= sub_end) break;
            const int off = __ldg(slice_off + s);
            const int groups = (__ldg(slice_off + s + 1) - off) / (kSellGroup * kSellSliceRows);
            const uint4* p = reinterpret_cast<const uint4*>(idx + off) + lane;
            float acc0 = 0.f, acc1 = 0.f;
            int g = 0;
            for (; g + UNROLL <= groups; g += UNROLL) {
                uint4 q[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) q[u] = ld_stream_u32x4(p + (size_t)(g + u) * kSellSliceRows);
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    if (u & 1) acc1 += sell_gather8(ysm, q[u]);
                    else acc0 += sell_gather8(ysm, q[u]);
                }
            }
            for (; g < groups; ++g) acc0 += sell_gather8(ysm, ld_stream_u32x4(p + (size_t)g * kSellSliceRows));
            vpart[(size_t)s * kSellSliceRows + lane] = acc0 + acc1;
        }
    }
}

// ---- per-row epilogue -----------------------------------------------------------
struct SellEpilogueParams {
    const int32_t* rv_ptr;
    const int32_t* rv_idx;
    const float* vpart;
    const float* y_prev;     // gather operand (for the edge flips)
    const float* dinv;
    const uint8_t* iso;
    const float* tprev;
    const float* tprev2;     // may alias tk
    float* tk;               // or NULL
    float* y_out;            // or NULL
    float* out;              // [n, S]
    int32_t n, S, first, normalize;   // n: rows of this launch
    int32_t row0;                     // global id of local row 0 (dinv/iso/deltas are global)
    float a, b;
    float c_prev[EGNN_MAX_SCALES];
    float c_k[EGNN_MAX_SCALES];
    DeltaList delta;
};

__global__ void __launch_bounds__(256)
sell_epilogue_kernel(const __grid_constant__ SellEpilogueParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    // a hub row owns hundreds of virtual rows: add their partials in float64,
    // always in the same order (deterministic)
    double accd = 0.0;
    const int e = __ldg(p.rv_ptr + i + 1);
    for (int t = __ldg(p.rv_ptr + i); t < e; ++t) accd += (double)__ldg(p.vpart + __ldg(p.rv_idx + t));
    for (int d = 0; d < p.delta.n; ++d)
        if (p.delta.row[d] == i + p.row0 && p.delta.col[d] != i + p.row0)
            accd += (double)p.delta.val[d] * (double)__ldg(p.y_prev + p.delta.col[d]);
    const float acc = (float)accd;
    const float di = __ldg(p.dinv + p.row0 + i);
    const float theta = fmaf(p.a, 1.f - (float)__ldg(p.iso + p.row0 + i), p.b);
    float xprev = 0.f;
    if (theta != 0.f || p.first) xprev = p.tprev[i];
    const float lap = fmaf(theta, xprev, -p.a * di * acc);
    const float tk = p.first ? lap : fmaf(2.f, lap, -p.tprev2[i]);
    if (p.tk) p.tk[i] = tk;
    if (p.y_out) p.y_out[i] = di * tk;
    for (int s = 0; s < p.S; ++s) {
        float o = p.first ? fmaf(p.c_k[s], tk, p.c_prev[s] * xprev)
                          : fmaf(p.c_k[s], tk, p.out[(size_t)i * p.S + s]);
        if (p.normalize) o = o / (fabsf(o) + 1e-8f);
        p.out[(size_t)i * p.S + s] = o;
    }
}

}  // namespace egnn
