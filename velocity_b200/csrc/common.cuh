// Shared device/host helpers for libvelocity_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/velocity_b200.h"

#define VEL_API extern "C" __attribute__((visibility("default")))

// thread-local last-error buffer (defined in api.cu)
void vel_set_error(const char* fmt, ...);

// Stream-ordered scratch (cudaMallocAsync) would be returned to the OS at every synchronisation with the default release
// threshold of the device's memory pool and re-mapped by the next call (milliseconds for a few MB); entry points that
// take scratch call this first: it raises the threshold once per device (defined in api.cu).
void vel_keep_async_pool_cached();

#define VEL_CHECK_ARG(cond, ...)          \
    do {                                  \
        if (!(cond)) {                    \
            vel_set_error(__VA_ARGS__);   \
            return VEL_ERR_INVALID;       \
        }                                 \
    } while (0)

#define VEL_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            vel_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return VEL_ERR_CUDA;                                                            \
        }                                                                                   \
    } while (0)

#define VEL_LAUNCH_CHECK(name)                                                         \
    do {                                                                               \
        cudaError_t e__ = cudaGetLastError();                                          \
        if (e__ != cudaSuccess) {                                                      \
            vel_set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));   \
            return VEL_ERR_CUDA;                                                       \
        }                                                                              \
    } while (0)

// dense_f64.cu: bytes of scratch vel_syrk_lower_sub needs for an m x k operand (ba.cu sizes its workspace with it)
size_t vel_dense_syrk_workspace(int m, int k);
// dense_f64.cu: vel_syrk_lower_sub_rows / vel_spd_solve with a device-side gate (NULL = always run): when *gate != 0 at launch
// time the kernels return immediately.  ba_loop.cu enqueues a whole iteration loop and lets the convergence test close the gate.
int vel_dense_syrk_rows_gated(const double* E, int64_t ld, int32_t m, int32_t k, double* S, int64_t lds, void* work, size_t work_bytes,
                              int32_t blk_lo, int32_t blk_hi, const int* gate, vel_stream_t stream);
int vel_dense_spd_solve_gated(double* S, int64_t lds, int32_t n, double* b, int32_t* info, const int* gate, vel_stream_t stream);

static constexpr int kNumSMs = 148;  // B200

// BORDER_REFLECT_101 (gfedcb|abcdefgh|gfedcba).  Valid for any overshoot when n > 1.
__host__ __device__ __forceinline__ int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        i = i < 0 ? -i : 2 * n - 2 - i;
    }
    return i;
}

// single-overshoot version for hot loops (callers guarantee |overshoot| < n)
__device__ __forceinline__ int reflect101_1(int i, int n)
{
    i = i < 0 ? -i : i;
    return i >= n ? 2 * n - 2 - i : i;
}

__device__ __forceinline__ unsigned ldg_u8(const uint8_t* p) { return (unsigned)__ldg(p); }
