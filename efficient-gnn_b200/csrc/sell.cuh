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
//   * stored self loops and padding point at zero slots, so the inner loop has
//     no predicate: load 8 indices -> 8 LDS -> 8 FADD,
//   * the order of the entries inside a virtual row is free, so it is chosen
//     to spread the warp's 32 simultaneous gathers over the 32 banks (greedy
//     edge colouring of lanes x banks per slice, see sell_fill_kernel), and the slices of a
//     block are stored in a strided order of their length so every CTA's
//     contiguous range of slices carries the same mix of long and short rows.
// Per order: sell_spmv_kernel (persistent, one CTA per SM, entries split evenly
// over CTAs) writes one partial sum per virtual row into a row-major array;
// sell_epilogue_kernel adds each row's partials in a fixed order
// (deterministic), applies the Laplacian scaling, the three-term recurrence and
// the scale accumulation.
#pragma once

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "cheb.cuh"
#include "common.cuh"
#include "peer.cuh"

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
// Slices of a block are stored in a strided order of their length rank (slice
// t holds the rows of rank (t * stride) mod n_c, stride ~ 0.618 n_c, coprime to
// n_c): any contiguous range of slices - what one CTA of the hot kernel gets -
// then carries the same mix of long and short rows, so equal entries mean equal time.
__device__ __forceinline__ int sell_gcd(int a, int b) {
    while (b) { const int t = a % b; a = b; b = t; }
    return a;
}

__global__ void sell_blocks_kernel(int n, int C, const int32_t* __restrict__ u_off,
                                   int32_t* __restrict__ q_ptr, int32_t* __restrict__ vp_ptr,
                                   int32_t* __restrict__ blk_slice_ptr, int32_t* __restrict__ blk_stride,
                                   int64_t* __restrict__ totals) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int vp = 0;
    for (int c = 0; c < C; ++c) {
        const int q0 = u_off[(size_t)c * n];
        const int q1 = u_off[(size_t)(c + 1) * n];
        q_ptr[c] = q0;
        vp_ptr[c] = vp;
        blk_slice_ptr[c] = vp / kSellSliceRows;
        const int n_c = (q1 - q0 + kSellSliceRows - 1) / kSellSliceRows;
        int stride = (int)(0.6180339887 * (double)n_c);
        if (stride < 1) stride = 1;
        while (sell_gcd(stride, n_c > 0 ? n_c : 1) != 1) ++stride;
        blk_stride[c] = stride;
        vp += n_c * kSellSliceRows;
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

// length rank (in slices) of the rows stored in slice s of block c
__device__ __forceinline__ int sell_rank_of_slice(const int32_t* blk_slice_ptr, const int32_t* blk_stride, int c, int s) {
    const int n_c = blk_slice_ptr[c + 1] - blk_slice_ptr[c];
    return (int)(((int64_t)(s - blk_slice_ptr[c]) * blk_stride[c]) % n_c);
}

// ---- build pass 4: padded size of every slice --------------------------------
__global__ void __launch_bounds__(256)
sell_slice_kernel(int C, int lmax, const uint32_t* __restrict__ key_sorted,
                  const int32_t* __restrict__ q_ptr, const int32_t* __restrict__ blk_slice_ptr,
                  const int32_t* __restrict__ blk_stride, const int64_t* __restrict__ totals,
                  int32_t* __restrict__ slice_sz) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= (int)totals[kTotSlices]) return;
    const int c = sell_block_of_slice(blk_slice_ptr, C, s);
    const int rank0 = sell_rank_of_slice(blk_slice_ptr, blk_stride, c, s) * kSellSliceRows;
    const int len = lmax - (int)(key_sorted[q_ptr[c] + rank0] & 0xffffu);     // longest row of the slice
    slice_sz[s] = (len + kSellGroup - 1) / kSellGroup * kSellGroup * kSellSliceRows;
}

// CTA g of the hot kernel handles slices [cta_ptr[g], cta_ptr[g+1]): an even
// split of the padded entries on slice boundaries, computed once per plan (in
// the hot kernel the two binary searches were ~30 dependent loads per CTA).
// cta_ptr[n_cta + 1 + g] = column block of that CTA's first slice.
__global__ void sell_cta_ranges_kernel(const int32_t* __restrict__ slice_off, const int32_t* __restrict__ blk_slice_ptr,
                                       int C, int n_slices, int n_cta, int32_t* __restrict__ cta_ptr) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > n_cta) return;
    if (g == n_cta) { cta_ptr[g] = n_slices; cta_ptr[n_cta + 1 + g] = C; return; }
    const int64_t total = slice_off[n_slices];
    const int64_t want = total * g / n_cta;
    int lo = 0, hi = n_slices;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((int64_t)slice_off[mid] < want) lo = mid + 1; else hi = mid;
    }
    cta_ptr[g] = lo;
    cta_ptr[n_cta + 1 + g] = sell_block_of_slice(blk_slice_ptr, C, lo);
}

__global__ void sell_totals_kernel(const int32_t* __restrict__ slice_off, const int32_t* __restrict__ rv_ptr,
                                   int n, int64_t* __restrict__ totals) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    totals[kTotEntries] = slice_off[totals[kTotSlices]];
    totals[kTotRowV] = rv_ptr[n];
}

// ---- build pass 5: write the lane-interleaved 16-bit index stream -----------
// Entry order inside a virtual row is free (the partial sum is a plain sum), so
// it is chosen to keep the hot kernel's shared-memory gathers off each other's
// banks: the warp's LDS for position j reads one entry per lane, so a slice is
// conflict-free when, at every position, the 32 lanes hold 32 different banks
// (bank = local column mod 32).  That is an edge colouring of the bipartite
// multigraph lanes x banks with the positions as colours; it is done greedily
// (first fit): lane r walks its own virtual row, and each entry takes the
// lowest position that is free both in the lane and in the entry's bank.  One
// warp per slice; lanes whose current entries share a bank go one at a time
// (match_any), so a bank's position mask has one writer per round.  The few
// entries (2-5 %) with no common free position take the lane's lowest free
// one; unused positions point at the zero slot of bank (r + j) mod 32 (32 zero
// slots at CB .. CB+31; stored self loops too).  Measured on the Reddit shape:
// 3.7 wavefronts per LDS unordered, 2.5 with a fixed diagonal rule, see
// DESIGN.md for the first-fit figure.  The 16-bit stream is written straight
// to global memory (2-byte stores that L2 merges).
constexpr int kSellFillWarps = 8;
constexpr int kSellLmaxCap = 256;          // longest virtual row
constexpr int kSellMaskWords = kSellLmaxCap / 32;
constexpr size_t kSellFillSmem = (size_t)kSellFillWarps * 32 * kSellMaskWords * sizeof(uint32_t);   // bank masks

__global__ void __launch_bounds__(kSellFillWarps * 32)
sell_fill_kernel(const int32_t* __restrict__ colidx, int C, int CB, int lmax, int n_slices,
                 const uint32_t* __restrict__ key_sorted, const int32_t* __restrict__ perm,
                 const int32_t* __restrict__ vsrc, const int32_t* __restrict__ vslot_u,
                 const int32_t* __restrict__ vrow, const int32_t* __restrict__ q_ptr,
                 const int32_t* __restrict__ vp_ptr, const int32_t* __restrict__ blk_slice_ptr,
                 const int32_t* __restrict__ blk_stride, const int32_t* __restrict__ slice_off,
                 uint16_t* __restrict__ idx, int32_t* __restrict__ vslot_out, int row0) {
    extern __shared__ __align__(16) uint32_t fill_smem[];
    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    const int s = blockIdx.x * kSellFillWarps + wid;
    if (s >= n_slices) return;
    uint32_t* bank_used = fill_smem + (size_t)wid * 32 * kSellMaskWords;   // [bank][word]: positions taken in that bank

    const int c = sell_block_of_slice(blk_slice_ptr, C, s);
    const int v = s * kSellSliceRows + lane;                          // storage slot of this lane's virtual row
    const int rank = sell_rank_of_slice(blk_slice_ptr, blk_stride, c, s) * kSellSliceRows + lane;
    const bool valid = rank < (q_ptr[c + 1] - q_ptr[c]);
    int len = 0, src = 0, row = -1, slot = -1;
    if (valid) {
        const int q = q_ptr[c] + rank;
        const int u = perm[q];
        len = lmax - (int)(key_sorted[q] & 0xffffu);
        src = vsrc[u];
        row = vrow[u] + row0;             // global id of the row (diagonal test)
        slot = vslot_u[u];
    }
    vslot_out[v] = slot;                  // -1: padding lane of the block's last slice
    const int off = slice_off[s];
    const int lpad = (slice_off[s + 1] - off) / kSellSliceRows;      // positions per lane
    const int col0 = c * CB;
    // position j of this lane lives at idx[off + ((j >> 3) * 32 + lane) * 8 + (j & 7)]
    uint16_t* mine = idx + off + lane * kSellGroup;

    uint32_t lane_used[kSellMaskWords];                               // positions taken in this lane (>= lpad: never free)
#pragma unroll
    for (int w = 0; w < kSellMaskWords; ++w) {
        const int lo = w * 32;
        lane_used[w] = lpad >= lo + 32 ? 0u : (lpad <= lo ? 0xffffffffu : ~((1u << (lpad - lo)) - 1u));
        bank_used[lane * kSellMaskWords + w] = 0u;
    }
    __syncwarp();

    // the lane's next CSR entry is requested one step ahead of its use
    int i = 0, loc = -1;
    int col_next = len > 0 ? __ldg(colidx + src) : 0;
    while (true) {
        while (loc < 0 && i < len) {                                  // take the next entry (skip a stored self loop)
            const int col = col_next;
            ++i;
            if (i < len) col_next = __ldg(colidx + src + i);
            if (col != row) loc = col - col0;
        }
        const bool pending = loc >= 0;
        if (!__any_sync(0xffffffffu, pending)) break;
        const int b = loc & 31;
        const unsigned grp = __match_any_sync(0xffffffffu, pending ? b : 32 + lane);
        if (pending && (__ffs(grp) - 1) == lane) {                    // one lane per bank and round
            uint32_t* bu = bank_used + b * kSellMaskWords;
            int p = -1;
#pragma unroll
            for (int w = 0; w < kSellMaskWords; ++w) {
                const uint32_t m = ~(lane_used[w] | bu[w]);
                if (p < 0 && m) p = w * 32 + __ffs(m) - 1;
            }
            const bool clean = p >= 0;
            if (!clean) {                                             // no common free position: lowest free in the lane
#pragma unroll
                for (int w = 0; w < kSellMaskWords; ++w) {
                    const uint32_t m = ~lane_used[w];
                    if (p < 0 && m) p = w * 32 + __ffs(m) - 1;
                }
            }
            const int wp = p >> 5;
            const uint32_t bit = 1u << (p & 31);
#pragma unroll
            for (int w = 0; w < kSellMaskWords; ++w) lane_used[w] |= (w == wp) ? bit : 0u;   // stays in registers
            if (clean) bu[wp] |= bit;
            mine[(p >> 3) * (kSellGroup * kSellSliceRows) + (p & 7)] = (uint16_t)loc;
            loc = -1;
        }
        __syncwarp();
    }
    // unused positions -> zero slot of bank (lane + p) mod 32
#pragma unroll
    for (int w = 0; w < kSellMaskWords; ++w) {
        uint32_t m = ~lane_used[w];
        while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            const int p = w * 32 + bit;
            mine[(p >> 3) * (kSellGroup * kSellSliceRows) + (p & 7)] = (uint16_t)(CB + ((lane + p) & 31));
        }
    }
}

// ---- hot kernel ---------------------------------------------------------------
// y: the gather operand dinv (.) T_{k-1}, [n] float32.  The partial sum of
// virtual row v goes to vpart[vslot[v]] (row-major: a row's partials are
// contiguous for the epilogue).  Dynamic shared memory: (CB + 32) floats.
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

constexpr int kSellZeroSlots = 32;
constexpr int kSellUnroll = 8;            // 16-byte index loads in flight per lane (measured best of 2/4/6/8)

// PEER: the operand lives in this rank's exchange window and is written by the
// other GPUs over NVLink; the CTA first waits for their flags, and stages with
// L2-coherent loads (ld.global.cg) instead of the non-coherent path.
struct SellPeerWait {
    const unsigned* local_flags;
    const unsigned* epoch;
    unsigned* error;
    int32_t world, rank;
};

template <int UNROLL, bool PEER>
__global__ void __launch_bounds__(kSellThreads, 1)
sell_spmv_kernel(const uint16_t* __restrict__ idx, const int32_t* __restrict__ slice_off,
                 const int32_t* __restrict__ blk_slice_ptr, const int32_t* __restrict__ vslot,
                 const int32_t* __restrict__ cta_ptr, int n_cta, int C, int CB, const float* y, int n,
                 float* __restrict__ vpart, const SellPeerWait pw) {
    extern __shared__ __align__(16) float ysm[];
    __shared__ int next_slice;
    __shared__ int bsp[kSellMaxBlocks + 1];
    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    constexpr int kWarps = kSellThreads / 32;

    // everything that does not depend on the operand is requested first: this CTA's slice
    // range and first column block (precomputed per plan), the block pointers, and - below -
    // the index groups of every warp's first slice, which then fly during the flag wait
    // (row-sharded) and the staging of the operand
    const int s_begin = __ldg(cta_ptr + blockIdx.x), s_end = __ldg(cta_ptr + blockIdx.x + 1);
    const int c_first = __ldg(cta_ptr + n_cta + 1 + blockIdx.x);
    if (threadIdx.x <= C) bsp[threadIdx.x] = __ldg(blk_slice_ptr + threadIdx.x);
    __syncthreads();

    bool waited = !PEER;
    for (int c = c_first; c < C; ++c) {
        const int sub_begin = max(s_begin, bsp[c]);
        const int sub_end = min(s_end, bsp[c + 1]);
        if (sub_begin >= s_end) break;
        if (sub_begin >= sub_end) continue;

        // this warp's first slice of the block is fixed (sub_begin + warp id); the rest are
        // handed out dynamically from a shared counter
        int s = sub_begin + wid;
        int off = 0, end = 0, slot = -1;
        if (s < sub_end) {
            off = __ldg(slice_off + s);
            end = __ldg(slice_off + s + 1);
            slot = __ldg(vslot + (size_t)s * kSellSliceRows + lane);
        }
        // ask L2 for the head of that slice's index stream (one 128-byte line per lane, no registers held)
        if (s < sub_end) {
            const char* head = reinterpret_cast<const char*>(idx + off) + lane * 128;
            if (head < reinterpret_cast<const char*>(idx + end)) asm volatile("prefetch.global.L2 [%0];" ::"l"(head));
        }
        if (!waited) {                                     // operand written by the other GPUs: wait for their flags
            peer_consumer_wait(pw.local_flags, pw.epoch, pw.world, pw.rank, pw.error);
            waited = true;
        }
        __syncthreads();                                   // previous block's gathers are done
        // stage the column block of the operand in shared memory
        const int col0 = c * CB;
        const int cnt = min(CB, n - col0);
        const float* src = y + col0;
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            const int cnt4 = cnt >> 2;
            const float4* src4 = reinterpret_cast<const float4*>(src);
            float4* dst4 = reinterpret_cast<float4*>(ysm);
            for (int t = threadIdx.x; t < cnt4; t += kSellThreads) dst4[t] = PEER ? __ldcg(src4 + t) : __ldg(src4 + t);
            for (int t = (cnt4 << 2) + threadIdx.x; t < cnt; t += kSellThreads) ysm[t] = PEER ? __ldcg(src + t) : __ldg(src + t);
        } else {
            for (int t = threadIdx.x; t < cnt; t += kSellThreads) ysm[t] = PEER ? __ldcg(src + t) : __ldg(src + t);
        }
        for (int t = cnt + threadIdx.x; t < CB + kSellZeroSlots; t += kSellThreads) ysm[t] = 0.f;   // includes the zero slots
        if (threadIdx.x == 0) next_slice = sub_begin + kWarps;
        __syncthreads();

        while (s < sub_end) {
            // the NEXT slice's offsets and slot are fetched while this one is summed
            int s_next = 0;
            if (lane == 0) s_next = atomicAdd(&next_slice, 1);
            s_next = __shfl_sync(0xffffffffu, s_next, 0);
            int off_next = 0, end_next = 0, slot_next = -1;
            if (s_next < sub_end) {
                off_next = __ldg(slice_off + s_next);
                end_next = __ldg(slice_off + s_next + 1);
                slot_next = __ldg(vslot + (size_t)s_next * kSellSliceRows + lane);
            }
            const int groups = (end - off) / (kSellGroup * kSellSliceRows);
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
            if (slot >= 0) vpart[slot] = acc0 + acc1;
            s = s_next; off = off_next; end = end_next; slot = slot_next;
        }
    }
    if (!waited) peer_consumer_wait(pw.local_flags, pw.epoch, pw.world, pw.rank, pw.error);   // CTA without slices
}

// ---- per-row epilogue -----------------------------------------------------------
struct SellEpilogueParams {
    const int32_t* rv_ptr;
    const float* vpart;      // row-major partial sums: row i owns [rv_ptr[i], rv_ptr[i+1])
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
    PeerPush peer;           // world > 1: dinv (.) T_k goes into every rank's exchange window
};

__device__ __forceinline__ void sell_epilogue_row(const SellEpilogueParams& p, int i) {
    // everything that does not depend on the partial sums is requested first, so the row pays
    // two dependent round trips (row pointers -> partials) instead of three
    const int e = __ldg(p.rv_ptr + i + 1);
    int t = __ldg(p.rv_ptr + i);
    const float di = __ldg(p.dinv + p.row0 + i);
    const float theta = fmaf(p.a, 1.f - (float)__ldg(p.iso + p.row0 + i), p.b);
    float xprev = 0.f;
    if (theta != 0.f || p.first) xprev = p.tprev[i];
    const float t2 = p.first ? 0.f : p.tprev2[i];
    float prev_out = 0.f;
    if (!p.first && p.S == 1) prev_out = p.out[i];
    // a hub row owns hundreds of virtual rows: add their partials in float64,
    // always in the same order (deterministic)
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;          // four interleaved chains, always combined the same way
    for (; t + 4 <= e; t += 4) {
        a0 += (double)p.vpart[t];
        a1 += (double)p.vpart[t + 1];
        a2 += (double)p.vpart[t + 2];
        a3 += (double)p.vpart[t + 3];
    }
    for (; t < e; ++t) a0 += (double)p.vpart[t];
    double accd = (a0 + a1) + (a2 + a3);
    for (int d = 0; d < p.delta.n; ++d)
        if (p.delta.row[d] == i + p.row0 && p.delta.col[d] != i + p.row0)
            accd += (double)p.delta.val[d] * (double)__ldg(p.y_prev + p.delta.col[d]);
    const float acc = (float)accd;
    const float lap = fmaf(theta, xprev, -p.a * di * acc);
    const float tk = p.first ? lap : fmaf(2.f, lap, -t2);
    if (p.tk) p.tk[i] = tk;
    if (p.y_out) p.y_out[i] = di * tk;
    if (p.peer.world > 1 && p.peer.has_data) {
        const float yv = di * tk;
        for (int r = 0; r < p.peer.world; ++r) p.peer.dst[r][p.row0 + i] = yv;
    }
    for (int s = 0; s < p.S; ++s) {
        float o = p.first ? fmaf(p.c_k[s], tk, p.c_prev[s] * xprev)
                          : fmaf(p.c_k[s], tk, p.S == 1 ? prev_out : p.out[(size_t)i * p.S + s]);
        if (p.normalize) o = o / (fabsf(o) + 1e-8f);
        p.out[(size_t)i * p.S + s] = o;
    }
}

// Launched with one row per thread (256-thread CTAs) on a single GPU; with the exchange
// fused it runs as at most one 1024-thread CTA per SM striding over the rows, because every
// CTA ends with a system-scope fence and those serialise per SM (8 CTAs per SM measured
// ~8 us of fences per order, one CTA per SM ~3 us).
__global__ void __launch_bounds__(1024)
sell_epilogue_kernel(const __grid_constant__ SellEpilogueParams p) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += gridDim.x * blockDim.x) sell_epilogue_row(p, i);
    peer_producer_signal(p.peer);
}

}  // namespace egnn
