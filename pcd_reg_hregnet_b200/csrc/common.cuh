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

// cudaFuncSetAttribute is PER DEVICE: a call site that raises a kernel's dynamic shared-memory limit must do so once on
// every device the process launches on (one process may drive several GPUs), not once per process.
struct hrn_once_per_device {
    unsigned long long mask = 0;
    bool need() {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d > 63) return true;
        const unsigned long long bit = 1ull << d;
        return (__atomic_fetch_or(&mask, bit, __ATOMIC_RELAXED) & bit) == 0;
    }
};

static inline int hrn_divup(long long a, long long b) { return (int)((a + b - 1) / b); }

// Order-preserving map float -> uint32 (a < b  <=>  ord(a) < ord(b) for non-NaN a, b).
__device__ __forceinline__ unsigned hrn_ford(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float hrn_ford_inv(unsigned o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// Packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2 on sm_100): two points per instruction, each half rounded exactly
// like the scalar __fsub_rn / __fmul_rn / __fmaf_rn -- halves the arithmetic instructions of the distance loops (FPS, kNN).
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t f2_pack(float lo, float hi) {
    f32x2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void f2_unpack(f32x2_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2_t f2_add(f32x2_t a, f32x2_t b) { f32x2_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2_t f2_sub(f32x2_t a, f32x2_t b) { f32x2_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2_t f2_mul(f32x2_t a, f32x2_t b) { f32x2_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2_t f2_fma(f32x2_t a, f32x2_t b, f32x2_t c) {
    f32x2_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
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
