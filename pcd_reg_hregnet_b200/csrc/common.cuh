// Shared device/host helpers for the sm_100a kernels of the HRegNet registration forward path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define HRN_API extern "C" __attribute__((visibility("default")))

// Error codes of the C ABI (include/hregnet_b200.h).  >= 1000: argument errors; otherwise cudaError_t.
#define HRN_OK 0
#define HRN_ERR_BAD_ARG 1001
#define HRN_ERR_UNSUPPORTED 1002

#define HRN_LAUNCH_CHECK()                                   \
    do {                                                     \
        cudaError_t e__ = cudaGetLastError();                \
        if (e__ != cudaSuccess) return (int)e__;             \
    } while (0)

#define HRN_CUDA(call)                                       \
    do {                                                     \
        cudaError_t e__ = (call);                            \
        if (e__ != cudaSuccess) return (int)e__;             \
    } while (0)

static inline int hrn_divup(long long a, long long b) { return (int)((a + b - 1) / b); }

// Order-preserving map float -> uint32 (a < b  <=>  ord(a) < ord(b) for non-NaN a, b).
__device__ __forceinline__ unsigned hrn_ford(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float hrn_ford_inv(unsigned o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__device__ __forceinline__ float hrn_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float hrn_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
