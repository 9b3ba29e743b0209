// Wide-signal Chebyshev order kernel (F >= 8): CSR SpMM over the implicit
// scaled Laplacian with the recurrence, the scale accumulation and the L1
// normalisation fused, as cheb.cuh, but organised around what bounds it at
// these widths - the latency and L2 bandwidth of the T_{k-1} row gathers:
//
//  * the gather source is always y = dinv (.) T_{k-1} (pre-scaled once per
//    order by the producing epilogue), so an entry costs one index, no dinv
//    lookup; the own-row T_{k-1}, T_{k-2} are recovered as y / dinv_i;
//  * a warp owns one row x one feature tile (<= 128 columns: 32 lanes x
//    float4, or NZ entries in parallel when the tile is narrower); the row's
//    indices are fetched 32 at a time with one coalesced load and broadcast
//    by shuffle, 8 row gathers are in flight per lane, and the NEXT row's
//    pointers and first index batch are fetched while this row is summed -
//    short rows (arxiv / Physics shape: ~14 entries) no longer pay five
//    dependent round trips each;
//  * rows are visited in degree-descending order (egnn_row_order) by a
//    persistent grid, longest first, so the tail of the launch is made of
//    short rows; rows above kHubDegree are summed by a whole CTA (8 warps,
//    fixed-order shared-memory reduction: deterministic);
//  * gathers carry an L2 evict_last policy, the streamed own-row operands and
//    results evict_first, so T_{k-1} stays resident in the 126 MB L2 while
//    T_{k-2}, the accumulators and T_k stream through it.
// Reference semantics: calibration/WATS.py:29-37, :55, :65-68, :71-72.
#pragma once

#include "cheb.cuh"
#include "common.cuh"
#include "peer.cuh"

namespace egnn {

constexpr int kWideBlock = 256;
constexpr int kWideWarps = kWideBlock / 32;
constexpr int kWideUnroll = 8;           // row gathers in flight per lane
constexpr int kWideMinBlocks = 3;        // CTAs per SM the register budget is sized for
constexpr int kHubDegree = 2048;         // rows at least this long are summed by a whole CTA

struct WideParams {
    const int32_t* rowptr;
    const int32_t* colidx;
    const float* vals;          // NULL: binary adjacency
    const int32_t* perm;        // processing order (degree-descending) or NULL: identity
    const int32_t* n_hub;       // device scalar: leading rows of perm that are CTA-cooperative (NULL: none)
    const float* dinv;          // [n_global]
    const uint8_t* iso;         // [n_global]
    const float* ysrc;          // dinv (.) T_{k-1}, indexed by GLOBAL column
    const float* x0_own;        // order 1: exact T_0 rows of this launch; NULL later
    const float* y2_own;        // dinv (.) T_{k-2} rows of this launch (may alias y_out)
    float* y_out;               // dinv (.) T_k rows, or NULL
    float* tk_out;              // T_k rows (all orders requested), or NULL
    float* out;                 // [n_rows, S, F]
    int64_t n_rows;
    int64_t row0;
    int32_t F, S;
    float a, b;                 // operator = a * L_sym + b * I
    int32_t first, normalize, fl_log2;
    float c_prev[EGNN_MAX_SCALES];
    float c_k[EGNN_MAX_SCALES];
    DeltaList delta;
    PeerPush peer;              // world > 1: y_out rows also go to every rank's exchange window
};

__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// feature-vector loads/stores with an L2 cache-policy hint
template <int VEC> struct HintVec;
template <> struct HintVec<1> {
    float v[1];
    __device__ __forceinline__ void gather(const float* p, uint64_t pol) {
        asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v[0]) : "l"(p), "l"(pol));
    }
    __device__ __forceinline__ void load_stream(const float* p, uint64_t pol) {
        asm volatile("ld.global.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v[0]) : "l"(p), "l"(pol));
    }
    __device__ __forceinline__ void store_stream(float* p, uint64_t pol) const {
        asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v[0]), "l"(pol) : "memory");
    }
    __device__ __forceinline__ void store(float* p) const { *p = v[0]; }
};
template <> struct HintVec<4> {
    float v[4];
    __device__ __forceinline__ void gather(const float* p, uint64_t pol) {
        asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                     : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p), "l"(pol));
    }
    __device__ __forceinline__ void load_stream(const float* p, uint64_t pol) {
        asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                     : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p), "l"(pol));
    }
    __device__ __forceinline__ void store_stream(float* p, uint64_t pol) const {
        asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                     ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "l"(pol) : "memory");
    }
    __device__ __forceinline__ void store(float* p) const {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};

// Sum of w_e * y[c_e, tile] over entries q in [qs, qe) taken by this lane's
// entry slot (nzl of NZ), starting from the first index batch already held in
// (pre_c, pre_w) when have_pre.  ybase[u] = ysrc + (clamped) column offset of
// the lane: lanes past the end of a ragged last tile read a valid column and
// are masked at the stores, so the loop carries no per-lane predicate.  Full
// groups of UNR entries per slot run branch-free; only the tail is guarded.
// Two-level float32 summation: acc folded into hi every 64 entries of a chain.
template <int VEC, int U, int NZ_LOG2, bool HAS_VALS>
__device__ __forceinline__ void wide_accumulate(const WideParams& p, int qs, int qe, int grow, int lane, int nzl,
                                                const float* const (&ybase)[U], bool have_pre, int pre_c, float pre_w,
                                                uint64_t pol_keep, float (&sum)[U][VEC]) {
    constexpr int NZ = 1 << NZ_LOG2;
    constexpr int UNR = NZ <= 4 ? kWideUnroll : 32 / NZ;      // entries in flight per slot; NZ * UNR <= 32
    constexpr int STEP = NZ * UNR;
    const int64_t F = p.F;
    float acc[U][VEC], hi[U][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int v = 0; v < VEC; ++v) { acc[u][v] = 0.f; hi[u][v] = 0.f; }
    int since_fold = 0;
    for (int b0 = qs; b0 < qe; b0 += 32) {
        const int cnt = min(32, qe - b0);
        int cj = grow;
        float wj = 0.f;
        if (have_pre && b0 == qs) {
            cj = pre_c;
            wj = pre_w;
        } else if (lane < cnt) {
            cj = ld_stream_i32(p.colidx + b0 + lane);
            wj = HAS_VALS ? ld_stream_f32(p.vals + b0 + lane) : 1.f;
        }
        if (lane >= cnt) { cj = grow; wj = 0.f; }
        if (cj == grow) wj = 0.f;                      // stored self loops are not part of L
        int e0 = 0;
        for (; e0 + STEP <= cnt; e0 += STEP) {         // full groups: no guards
            HintVec<VEC> x[UNR][U];
#pragma unroll
            for (int j = 0; j < UNR; ++j) {
                const int c = __shfl_sync(0xffffffffu, cj, e0 + j * NZ + nzl);
#pragma unroll
                for (int u = 0; u < U; ++u) x[j][u].gather(ybase[u] + c * F, pol_keep);
            }
#pragma unroll
            for (int j = 0; j < UNR; ++j) {
                const float w = __shfl_sync(0xffffffffu, wj, e0 + j * NZ + nzl);
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) acc[u][v] = fmaf(w, x[j][u].v[v], acc[u][v]);
            }
        }
        if (e0 < cnt) {                                // tail group (lanes past cnt hold c = own row, w = 0)
            HintVec<VEC> x[UNR][U];
#pragma unroll
            for (int j = 0; j < UNR; ++j) {
                if (e0 + j * NZ < cnt) {               // warp-uniform
                    const int c = __shfl_sync(0xffffffffu, cj, (e0 + j * NZ + nzl) & 31);
#pragma unroll
                    for (int u = 0; u < U; ++u) x[j][u].gather(ybase[u] + c * F, pol_keep);
                }
            }
#pragma unroll
            for (int j = 0; j < UNR; ++j) {
                if (e0 + j * NZ < cnt) {
                    const float w = __shfl_sync(0xffffffffu, wj, (e0 + j * NZ + nzl) & 31);
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int v = 0; v < VEC; ++v) acc[u][v] = fmaf(w, x[j][u].v[v], acc[u][v]);
                }
            }
        }
        since_fold += 32;
        if (since_fold >= 64 * NZ) {                   // 64 entries per chain
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int v = 0; v < VEC; ++v) { hi[u][v] += acc[u][v]; acc[u][v] = 0.f; }
            since_fold = 0;
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int v = 0; v < VEC; ++v) sum[u][v] = hi[u][v] + acc[u][v];
}

// Fused epilogue of one row tile (run by the lanes with nzl == 0).
template <int VEC, int U>
__device__ __forceinline__ void wide_epilogue(const WideParams& p, int64_t row, int grow, int lane, int FL,
                                              const int (&fidx)[U], const bool (&fok)[U], bool writer,
                                              float (&acc)[U][VEC], uint64_t pol_stream) {
    const int F = p.F;
    float di = 1.f, theta = 0.f;
    if (writer) {
        di = __ldg(p.dinv + grow);
        theta = fmaf(p.a, 1.f - (float)__ldg(p.iso + grow), p.b);
    }
    // edge flips on top of the CSR (UGCA recompute); tiny host-provided list
    if (p.delta.n > 0 && writer) {
        for (int e = 0; e < p.delta.n; ++e) {
            if (p.delta.row[e] == grow && p.delta.col[e] != grow) {
                const float* src = p.ysrc + (int64_t)p.delta.col[e] * F;
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (fok[u]) {
                        FeatVec<VEC> x;
                        x.load(src + fidx[u]);
#pragma unroll
                        for (int v = 0; v < VEC; ++v) acc[u][v] = fmaf(p.delta.val[e], x.v[v], acc[u][v]);
                    }
            }
        }
    }
    const float nscale = -p.a * di;
    const float inv_di = 1.f / di;
    float tk[U][VEC], xprev[U][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) { tk[u][v] = 0.f; xprev[u][v] = 0.f; }
        if (writer && fok[u]) {
            const int64_t off = row * F + fidx[u];
            if (p.first) {
                HintVec<VEC> t;
                t.load_stream(p.x0_own + off, pol_stream);
#pragma unroll
                for (int v = 0; v < VEC; ++v) xprev[u][v] = t.v[v];
            } else if (theta != 0.f) {                  // T_{k-1} of the own row = y / dinv
                FeatVec<VEC> t;
                t.load(p.ysrc + (int64_t)grow * F + fidx[u]);
#pragma unroll
                for (int v = 0; v < VEC; ++v) xprev[u][v] = t.v[v] * inv_di;
            }
            float t2[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) t2[v] = 0.f;
            if (!p.first) {
                HintVec<VEC> t;
                t.load_stream(p.y2_own + off, pol_stream);
#pragma unroll
                for (int v = 0; v < VEC; ++v) t2[v] = t.v[v] * inv_di;
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float lap = fmaf(theta, xprev[u][v], nscale * acc[u][v]);
                tk[u][v] = p.first ? lap : fmaf(2.f, lap, -t2[v]);
            }
            HintVec<VEC> o;
            if (p.tk_out) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) o.v[v] = tk[u][v];
                o.store_stream(p.tk_out + off, pol_stream);
            }
            if (p.y_out || (p.peer.world > 1 && p.peer.has_data)) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) o.v[v] = di * tk[u][v];
                if (p.y_out) o.store(p.y_out + off);
                if (p.peer.world > 1 && p.peer.has_data) {
                    const int64_t goff = (int64_t)grow * F + fidx[u];
                    for (int r = 0; r < p.peer.world; ++r) o.store(p.peer.dst[r] + goff);
                }
            }
        }
    }
    for (int s = 0; s < p.S; ++s) {
        float o[U][VEC];
        float l1 = 0.f;
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) o[u][v] = 0.f;
            if (writer && fok[u]) {
                const int64_t off = (row * p.S + s) * F + fidx[u];
                if (p.first) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) o[u][v] = fmaf(p.c_k[s], tk[u][v], p.c_prev[s] * xprev[u][v]);
                } else {
                    HintVec<VEC> t;
                    t.load_stream(p.out + off, pol_stream);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) o[u][v] = fmaf(p.c_k[s], tk[u][v], t.v[v]);
                }
#pragma unroll
                for (int v = 0; v < VEC; ++v) l1 += fabsf(o[u][v]);
            }
        }
        if (p.normalize) {       // uniform branch: every lane takes part in the shuffles
            for (int off = 1; off < FL; off <<= 1) l1 += __shfl_xor_sync(0xffffffffu, l1, off);
            const float inv = 1.f / (l1 + 1e-8f);
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int v = 0; v < VEC; ++v) o[u][v] *= inv;
        }
        if (writer) {
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (fok[u]) {
                    HintVec<VEC> t;
#pragma unroll
                    for (int v = 0; v < VEC; ++v) t.v[v] = o[u][v];
                    t.store_stream(p.out + (row * p.S + s) * F + fidx[u], pol_stream);
                }
        }
    }
}

template <int VEC, int U, int NZ_LOG2, bool HAS_VALS>
__global__ void __launch_bounds__(kWideBlock, kWideMinBlocks)
cheb_wide_kernel(const __grid_constant__ WideParams p) {
    __shared__ float hub_part[kWideWarps][32 * VEC * U];
    constexpr int FL_LOG2 = 5 - NZ_LOG2;
    constexpr int FL = 1 << FL_LOG2;
    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    const int fl = lane & (FL - 1);
    const int nzl = lane >> FL_LOG2;
    const int f_tile = blockIdx.y * (FL * VEC * U);
    int fidx[U];
    bool fok[U];
    const float* ybase[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        fidx[u] = f_tile + (u * FL + fl) * VEC;
        fok[u] = fidx[u] < p.F;
        if (!fok[u]) fidx[u] = p.F - VEC;                      // clamped: read something valid, never stored
        ybase[u] = p.ysrc + fidx[u];
    }
    const uint64_t pol_keep = l2_policy_evict_last();
    const uint64_t pol_stream = l2_policy_evict_first();
    const int n_rows = (int)p.n_rows;
    int n_hub = p.n_hub ? __ldg(p.n_hub) : 0;
    if (n_hub > n_rows) n_hub = n_rows;

    // ---- phase A: hub rows, one CTA each, longest first ---------------------
    for (int h = blockIdx.x; h < n_hub; h += gridDim.x) {
        const int row = p.perm ? __ldg(p.perm + h) : h;
        const int grow = (int)p.row0 + row;
        const int start = __ldg(p.rowptr + row), end = __ldg(p.rowptr + row + 1);
        // warp w sums a contiguous run of whole 32-entry batches
        const int n_batches = (end - start + 31) >> 5;
        const int per_warp = (n_batches + kWideWarps - 1) / kWideWarps;
        const int qs = min(end, start + wid * per_warp * 32);
        const int qe = min(end, qs + per_warp * 32);
        float part[U][VEC];
        wide_accumulate<VEC, U, NZ_LOG2, HAS_VALS>(p, qs, qe, grow, lane, nzl, ybase, false, 0, 0.f, pol_keep, part);
#pragma unroll
        for (int o = FL; o < 32; o <<= 1) {
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int v = 0; v < VEC; ++v) part[u][v] += __shfl_xor_sync(0xffffffffu, part[u][v], o);
        }
        __syncthreads();                                       // previous hub row's readers are done
        if (nzl == 0) {
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int v = 0; v < VEC; ++v) hub_part[wid][(u * FL + fl) * VEC + v] = part[u][v];
        }
        __syncthreads();
        if (wid == 0) {
            float acc[U][VEC];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    float t = 0.f;
                    for (int w = 0; w < kWideWarps; ++w) t += hub_part[w][(u * FL + fl) * VEC + v];   // fixed order
                    acc[u][v] = t;
                }
            wide_epilogue<VEC, U>(p, row, grow, lane, FL, fidx, fok, nzl == 0, acc, pol_stream);
        }
    }

    // ---- phase B: one row per warp, persistent over the ordered rows --------
    const int total_warps = gridDim.x * kWideWarps;
    int r = n_hub + blockIdx.x * kWideWarps + wid;
    int row = 0, start = 0, end = 0, pre_c = 0;
    float pre_w = 0.f;
    if (r < n_rows) {
        row = p.perm ? __ldg(p.perm + r) : r;
        start = __ldg(p.rowptr + row);
        end = __ldg(p.rowptr + row + 1);
        if (start + lane < end) {
            pre_c = ld_stream_i32(p.colidx + start + lane);
            pre_w = HAS_VALS ? ld_stream_f32(p.vals + start + lane) : 1.f;
        }
    }
    while (r < n_rows) {
        // the next row's pointers are requested now and its first index batch once they are back
        const int r_next = r + total_warps;
        int row_n = 0, start_n = 0, end_n = 0;
        if (r_next < n_rows) {
            row_n = p.perm ? __ldg(p.perm + r_next) : r_next;
            start_n = __ldg(p.rowptr + row_n);
            end_n = __ldg(p.rowptr + row_n + 1);
        }
        const int grow = (int)p.row0 + row;
        float acc[U][VEC];
        wide_accumulate<VEC, U, NZ_LOG2, HAS_VALS>(p, start, end, grow, lane, nzl, ybase, true, pre_c, pre_w, pol_keep, acc);
        int pre_c_n = 0;
        float pre_w_n = 0.f;
        if (r_next < n_rows && start_n + lane < end_n) {
            pre_c_n = ld_stream_i32(p.colidx + start_n + lane);
            pre_w_n = HAS_VALS ? ld_stream_f32(p.vals + start_n + lane) : 1.f;
        }
#pragma unroll
        for (int o = FL; o < 32; o <<= 1) {
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc[u][v] += __shfl_xor_sync(0xffffffffu, acc[u][v], o);
        }
        wide_epilogue<VEC, U>(p, row, grow, lane, FL, fidx, fok, nzl == 0, acc, pol_stream);
        r = r_next; row = row_n; start = start_n; end = end_n; pre_c = pre_c_n; pre_w = pre_w_n;
    }
    peer_producer_signal(p.peer);
}

// ---- processing order ---------------------------------------------------------
// key = ~degree so an ascending radix sort yields the longest rows first
__global__ void __launch_bounds__(256)
row_order_keys_kernel(const int32_t* __restrict__ rowptr, int n, uint32_t* __restrict__ keys,
                      int32_t* __restrict__ ids) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = ~(uint32_t)(rowptr[i + 1] - rowptr[i]);
    ids[i] = i;
}

__global__ void row_order_hubs_kernel(const uint32_t* __restrict__ keys_sorted, int n, int32_t* __restrict__ n_hub) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int lo = 0, hi = n;                          // first position whose degree < kHubDegree
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((int)(~keys_sorted[mid]) >= kHubDegree) lo = mid + 1; else hi = mid;
    }
    *n_hub = lo;
}

}  // namespace egnn
