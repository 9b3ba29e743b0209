// Plan-free narrow (F = 1) orders for the FIRST use of a large graph - calibrator
// construction (calibration/WATS.py:99 computes the features once) and the
// host-buffer entry that bench.py times end to end.  The SELL plan (sell.cuh)
// costs 2.6 ms to build on the Reddit shape and pays off from the second use
// on; the generic CSR kernel (cheb.cuh) gathers through L1/L2 at ~1.5 4-byte
// gathers per clock per SM (0.54 ms per order there).  This path takes the two
// ideas of the SELL kernel that need no re-layout:
//   * the operand dinv (.) T_{k-1} is staged in shared memory one column block
//     (<= 49152 nodes) at a time, by TMA bulk copies, so every entry is a
//     shared-memory gather;
//   * every CTA serves one column block; the rows' segments inside that block
//     ([first entry with column >= c CB, first with column >= (c+1) CB), found
//     once per graph by binary search in the column-sorted rows) are handed out
//     in batches of 32 rows per warp from a per-block counter.
// A warp sums a segment with coalesced 4-byte index loads (int32 CSR as it
// came from the host: 4 B per entry instead of the plan's 2), one partial sum
// per (column block, row); the epilogue kernel adds a row's C partials in
// float64 in fixed order and applies the Laplacian scaling, the recurrence and
// the scale accumulation (same arithmetic as the SELL epilogue).
#pragma once

#include "common.cuh"
#include "prep.cuh"
#include "sell.cuh"
#include "sell_step.cuh"

namespace egnn {

constexpr int kBlockedThreads = 1024;
constexpr int kBlockedBatch = 32;         // rows per warp and grab

// seg[c * n + i] = first position of row i with column >= c * CB, c = 0 .. C (rows must be column-sorted)
__global__ void __launch_bounds__(256)
blocked_bounds_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, int n, int C, int CB,
                      int32_t* __restrict__ seg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = rowptr[i], e = rowptr[i + 1];
    seg[i] = b;
    int prev = b;
    for (int c = 1; c < C; ++c) {
        const int bound = c * CB;
        int lo = prev, hi = e;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(colidx + mid) < bound) lo = mid + 1; else hi = mid;
        }
        seg[(size_t)c * n + i] = lo;
        prev = lo;
    }
    seg[(size_t)C * n + i] = e;
}

struct BlockedParams {
    const int32_t* colidx;
    const float* vals;            // NULL: binary adjacency
    const int32_t* seg;           // [(C + 1) * n]
    const float* y;               // operand dinv (.) T_{k-1}, [n]
    float* part;                  // [C * n] partial sums
    unsigned* counter;            // [C] zeroed: next batch of rows of every column block
    int32_t n, C, CB, n_cta;
};

template <bool HAS_VALS>
__global__ void __launch_bounds__(kBlockedThreads, 1)
blocked_spmv_kernel(const __grid_constant__ BlockedParams p) {
    extern __shared__ __align__(128) float ysm[];
    __shared__ __align__(8) unsigned long long stage_bar;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int kWarps = kBlockedThreads / 32;
    // CTAs cycle over the column blocks: block c is served by CTAs c, c + C, ...
    const int c = blockIdx.x % p.C;
    const int rank_in_block = blockIdx.x / p.C;
    const int ctas_in_block = (p.n_cta - c + p.C - 1) / p.C;
    const int col0 = c * p.CB;
    const int cnt = min(p.CB, p.n - col0);
    const int n_batches = (p.n + kBlockedBatch - 1) / kBlockedBatch;

    // stage the column block (bulk copies by one thread, partial 16-byte tail by the others)
    const float* src = p.y + col0;
    const int cnt4 = cnt & ~3;
    if (tid == 0) {
        mbar_init(&stage_bar, 1);
        fence_proxy_async();
    }
    __syncthreads();
    if (tid == 0 && cnt4 > 0) {
        mbar_expect_tx(&stage_bar, (unsigned)cnt4 * 4u);
        for (int o = 0; o < cnt4; o += kStageChunkFloats)
            bulk_g2s(ysm + o, src + o, (unsigned)min(kStageChunkFloats, cnt4 - o) * 4u, &stage_bar);
    }
    if (tid >= 32 && tid - 32 < cnt - cnt4) ysm[cnt4 + tid - 32] = __ldg(src + cnt4 + tid - 32);

    // the warp's first batch is fixed, later ones come from the block's counter (one ahead)
    int batch = wid * ctas_in_block + rank_in_block;
    unsigned raw_next = 0;
    if (lane == 0) raw_next = atomicAdd(p.counter + c, 1u);
    const int32_t* seg_lo = p.seg + (size_t)c * p.n;
    const int32_t* seg_hi = p.seg + (size_t)(c + 1) * p.n;
    int my_lo = 0, my_hi = 0;
    if (batch < n_batches) {
        const int row = batch * kBlockedBatch + lane;
        if (row < p.n) { my_lo = __ldg(seg_lo + row); my_hi = __ldg(seg_hi + row); }
    }
    if (cnt4 > 0) mbar_wait(&stage_bar, 0);
    __syncthreads();

    while (batch < n_batches) {
        const int batch_next = kWarps * ctas_in_block + (int)__shfl_sync(0xffffffffu, raw_next, 0);
        if (lane == 0) raw_next = atomicAdd(p.counter + c, 1u);
        int nx_lo = 0, nx_hi = 0;
        if (batch_next < n_batches) {
            const int row = batch_next * kBlockedBatch + lane;
            if (row < p.n) { nx_lo = __ldg(seg_lo + row); nx_hi = __ldg(seg_hi + row); }
        }
        const int row_base = batch * kBlockedBatch;
        float mine = 0.f;                                      // lane l keeps the sum of row row_base + l
        // the first four index loads of the NEXT row are in flight while this row is summed
        int cj[4];
        float wj[4];
        auto load4 = [&](int e, int hi, int row, int (&cc)[4], float (&ww)[4]) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int q = e + 32 * u;
                cc[u] = q < hi ? ld_stream_i32(p.colidx + q) : row;
                ww[u] = (HAS_VALS && q < hi) ? ld_stream_f32(p.vals + q) : 1.f;
            }
        };
        {
            const int lo0 = __shfl_sync(0xffffffffu, my_lo, 0), hi0 = __shfl_sync(0xffffffffu, my_hi, 0);
            load4(lo0 + lane, hi0, row_base, cj, wj);
        }
        for (int r = 0; r < kBlockedBatch; ++r) {
            const int lo = __shfl_sync(0xffffffffu, my_lo, r);
            const int hi = __shfl_sync(0xffffffffu, my_hi, r);
            const int row = row_base + r;
            int cn[4];
            float wn[4];
            if (r + 1 < kBlockedBatch) {
                const int lo_n = __shfl_sync(0xffffffffu, my_lo, r + 1), hi_n = __shfl_sync(0xffffffffu, my_hi, r + 1);
                load4(lo_n + lane, hi_n, row + 1, cn, wn);
            }
            // two-level sum (the reference's scipy path accumulates in float64): float32 over 64 entries per lane,
            // those chunks and the 32 lanes in float64, always in the same order
            float acc = 0.f;
            double accd = 0.0;
            int since = 0;
            for (int e = lo + lane; e < hi; e += 128) {        // four coalesced index loads per trip
                if (e != lo + lane) load4(e, hi, row, cj, wj);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float w = cj[u] == row ? 0.f : wj[u];           // stored self loops are not part of L; padding lanes too
                    const int loc = cj[u] == row ? 0 : cj[u] - col0;
                    acc = fmaf(w, ysm[loc], acc);
                }
                if (++since == 16) { accd += (double)acc; acc = 0.f; since = 0; }
            }
            accd += (double)acc;
            if (hi - lo > 2048) {                              // long segment: lanes combined in float64
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) accd += __shfl_xor_sync(0xffffffffu, accd, o);
                if (lane == r) mine = (float)accd;
            } else {
                const float a = warp_sum((float)accd);
                if (lane == r) mine = a;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) { cj[u] = cn[u]; wj[u] = wn[u]; }
        }
        const int row = row_base + lane;
        if (row < p.n) p.part[(size_t)c * p.n + row] = mine;
        batch = batch_next; my_lo = nx_lo; my_hi = nx_hi;
    }
}

struct BlockedEpilogueParams {
    const float* part;       // [C * n]
    const float* y_prev;     // operand of this order (for the edge flips)
    const float* dinv;
    const uint8_t* iso;
    const float* tprev;
    const float* tprev2;     // may alias tk
    float* tk;               // or NULL
    float* y_out;            // or NULL
    float* out;              // [n, S]
    int32_t n, C, S, first, normalize;
    float a, b;
    float c_prev[EGNN_MAX_SCALES];
    float c_k[EGNN_MAX_SCALES];
    DeltaList delta;
};

__global__ void __launch_bounds__(256)
blocked_epilogue_kernel(const __grid_constant__ BlockedEpilogueParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const float di = __ldg(p.dinv + i);
    const float theta = fmaf(p.a, 1.f - (float)__ldg(p.iso + i), p.b);
    const float xprev = p.tprev[i];
    const float t2 = p.first ? 0.f : p.tprev2[i];
    double accd = 0.0;
    for (int c = 0; c < p.C; ++c) accd += (double)p.part[(size_t)c * p.n + i];       // fixed order
    for (int d = 0; d < p.delta.n; ++d)
        if (p.delta.row[d] == i && p.delta.col[d] != i)
            accd += (double)p.delta.val[d] * (double)__ldg(p.y_prev + p.delta.col[d]);
    const float acc = (float)accd;
    const float lap = fmaf(theta, xprev, -p.a * di * acc);
    const float tk = p.first ? lap : fmaf(2.f, lap, -t2);
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
