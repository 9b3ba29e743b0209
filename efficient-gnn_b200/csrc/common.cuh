// Shared helpers for libegnn_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/egnn_b200.h"

namespace egnn {

// thread-local last-error string (the only state the library keeps)
char* last_error_buf();
void set_error(const char* fmt, ...);

inline int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return EGNN_OK;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return EGNN_ERR_CUDA;
}

#define EGNN_REQUIRE(cond, msg)                         \
    do {                                                \
        if (!(cond)) {                                  \
            ::egnn::set_error("invalid argument: %s", msg); \
            return EGNN_ERR_INVALID_ARG;                \
        }                                               \
    } while (0)

#define EGNN_LAUNCH_CHECK(what)                                        \
    do {                                                               \
        int _rc = ::egnn::check_cuda(cudaGetLastError(), what);        \
        if (_rc != EGNN_OK) return _rc;                                \
    } while (0)

constexpr int kSmCountB200 = 148;

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming (read-once) loads: keep the gather operands in L1/L2, not the CSR
__device__ __forceinline__ int ld_stream_i32(const int* p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_stream_f32(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

}  // namespace egnn
