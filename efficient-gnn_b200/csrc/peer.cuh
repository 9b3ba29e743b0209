// Exchange over peer memory (NVLink / NVSwitch) for the row-sharded path.
//
// Every rank of the node owns one "exchange window" - a device allocation that
// all ranks map through CUDA IPC (egnn_peer_alloc / egnn_peer_open).  The
// kernel that PRODUCES a rank's slab of the next order's operand stores it
// straight into every rank's window (plain st.global on mapped peer pointers),
// then the last CTA to finish raises this rank's flag in every window; the
// kernel that CONSUMES the operand spins on its own window's flags first.  One
// order therefore costs no collective launch: the transfer rides inside the
// epilogue and the barrier is a flag wait at the head of the next kernel.
//
// Window layout (same on every rank; offsets in bytes):
//   [0, 64)      uint32 flags[EGNN_MAX_RANKS]  flags[p]: last epoch rank p has signalled here
//   [256, 260)   uint32 epoch                  this rank's own epoch (local use only)
//   [260, 264)   uint32 done_ctr               CTAs of the producing kernel that have finished
//   [264, 268)   uint32 error                  set when a wait timed out
//   [4096, ...)  two operand buffers, each `slab_stride` bytes: the full operand
//                [world * rows_per, f] in global row order; order k reads buffer
//                (k-1) & 1 and its epilogue fills buffer k & 1.
// Epochs advance in lockstep (every rank runs the same kernel sequence), so a
// buffer is overwritten only after every reader of its previous contents has
// signalled a later epoch.
#pragma once

#include "common.cuh"

namespace egnn {

constexpr size_t kPeerFlagsOff = 0;
constexpr size_t kPeerEpochOff = 256;
constexpr size_t kPeerDoneOff = 260;
constexpr size_t kPeerErrorOff = 264;
constexpr size_t kPeerWaitNsOff = 512;     // uint64: total ns CTA 0 spent in flag waits (diagnostic)
constexpr size_t kPeerWaitCntOff = 520;    // uint64: number of such waits
constexpr size_t kPeerHeaderBytes = 4096;
constexpr unsigned long long kPeerTimeoutNs = 30000000000ull;     // 30 s: a rank that never shows up

struct PeerPush {
    int32_t world;                       // 0: no peer exchange
    int32_t rank;
    int32_t wait_first;                  // first kernel of a step: wait for the peers' current epoch before storing
    int32_t has_data;                    // 0: signal only (last order)
    float* dst[EGNN_MAX_RANKS];          // operand buffer being filled, in every rank's window
    unsigned* flag[EGNN_MAX_RANKS];      // &flags[rank] in every rank's window
    unsigned* local_flags;               // this rank's flags[EGNN_MAX_RANKS]
    unsigned* epoch;
    unsigned* done_ctr;
    unsigned* error;
};

// Flags are polled and raised with relaxed system-scope accesses; ordering comes from ONE
// fence.sys after the poll loop / before the flag stores (an acquire load or release store per
// peer would pay a fence each: measured ~4 us per peer on the 8-GPU box).
__device__ __forceinline__ unsigned ld_relaxed_sys_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_u32(unsigned* p, unsigned v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// One thread: wait until every peer has signalled `epoch` in this rank's window.
__device__ __forceinline__ void peer_wait_epoch(const unsigned* local_flags, int world, int rank, unsigned epoch,
                                                unsigned* error) {
    const unsigned long long t0 = global_timer_ns();
    for (int p = 0; p < world; ++p) {
        if (p == rank) continue;
        while ((int)(ld_relaxed_sys_u32(local_flags + p) - epoch) < 0) {
            __nanosleep(32);
            if (global_timer_ns() - t0 > kPeerTimeoutNs) {
                atomicExch(error, 1u);
                return;
            }
        }
    }
    __threadfence_system();            // acquire: the peers' operand stores are visible past this point
    if (blockIdx.x == 0 && blockIdx.y == 0) {          // diagnostic: how long the first CTA waited
        unsigned long long* stat = reinterpret_cast<unsigned long long*>(
            reinterpret_cast<char*>(const_cast<unsigned*>(local_flags)) - kPeerFlagsOff + kPeerWaitNsOff);
        stat[0] += global_timer_ns() - t0;
        stat[1] += 1ull;
    }
}

// Head of a consuming kernel: thread 0 of every CTA waits for the epoch the
// previous kernel on this stream left in the window, then the CTA proceeds.
__device__ __forceinline__ void peer_consumer_wait(const unsigned* local_flags, const unsigned* epoch_ptr, int world,
                                                   int rank, unsigned* error) {
    if (world > 1) {
        if (threadIdx.x == 0) peer_wait_epoch(local_flags, world, rank, *epoch_ptr, error);
        __syncthreads();
    }
}

// Tail of a producing kernel (all threads of every CTA call it after their
// stores): the last CTA to arrive bumps the epoch and raises this rank's flag
// in every window.  The CTA barrier orders every thread's peer stores before thread 0's
// fence.sys (cumulative), which orders them before the counter; the last CTA fences once
// more and raises the flags.
__device__ __forceinline__ void peer_producer_signal(const PeerPush& pp) {
    if (pp.world <= 1) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned total = gridDim.x * gridDim.y * gridDim.z;
        const unsigned prev = atomicAdd(pp.done_ctr, 1u);
        if (prev == total - 1) {
            *pp.done_ctr = 0u;                       // ready for the next producing launch
            const unsigned e = *pp.epoch + 1u;
            *pp.epoch = e;
            __threadfence_system();                  // release: every CTA's stores (seen through the counter) first
            for (int p = 0; p < pp.world; ++p)
                if (p != pp.rank) st_relaxed_sys_u32(pp.flag[p], e);
        }
    }
}

// First kernel of a step: before overwriting operand buffer 0, make sure every
// peer has finished the previous step (it signals once more after its last order).
__device__ __forceinline__ void peer_producer_wait_first(const PeerPush& pp) {
    if (pp.world > 1 && pp.wait_first) {
        if (threadIdx.x == 0) peer_wait_epoch(pp.local_flags, pp.world, pp.rank, *pp.epoch, pp.error);
        __syncthreads();
    }
}

// y = dinv (.) x for the local rows, stored into every rank's operand buffer 0
// (the order-1 operand).  x: [n_rows, f]; the window rows are ldy >= f wide and
// the padding columns are written as zeros.
__global__ void __launch_bounds__(1024)
peer_prescale_push_kernel(const float* __restrict__ x, const float* __restrict__ dinv, int64_t n_rows, int64_t row0,
                          int32_t f, int32_t ldy, const __grid_constant__ PeerPush pp) {
    peer_producer_wait_first(pp);
    const int64_t total = n_rows * ldy;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / ldy;
        const int c = (int)(i - r * ldy);
        const float v = c < f ? dinv[row0 + r] * x[r * f + c] : 0.f;
        for (int p = 0; p < pp.world; ++p) pp.dst[p][(row0 + r) * ldy + c] = v;
    }
    peer_producer_signal(pp);
}

}  // namespace egnn
