// Micro-probe: sustained tcgen05.mma issue rate for the operand layout used by the shared-MLP kernels
// (kind::f16, bf16, M=128, K=16, no-swizzle K-major core matrices), with different commit cadences.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../pcd_reg_hregnet_b200/csrc mma_probe.cu -o mma_probe
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

template <int commit_every, int uniform>
__global__ void __launch_bounds__(128) probe(int N, int n_mma, int rotate, long long* out) {
    const int two_acc = 0;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 48 * 1024; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // 192 KB of bf16 ~0.0078
    if (tid == 0) { mbar_init(smem_u32(&bar[0]), 1); mbar_init(smem_u32(&bar[1]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    if (uniform ? (warp == 0) : (tid == 0)) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t base = smem_u32(smem);
        const uint32_t a_bytes = 2 * 128 * 16, b_bytes = 2 * N * 16;        // one K=16 piece of A / B
        const int n_a = 8, n_b = rotate ? (int)((160 * 1024) / b_bytes) : 1;
        long long t0 = clock64();
        uint32_t ph = 0;
        const uint64_t ad0 = umma_desc(base, 128 * 16, 128), bd0 = umma_desc(base + 32 * 1024, N * 16, 128);
        const uint32_t a_step = rotate ? (a_bytes >> 4) : 0, b_step = rotate ? (b_bytes >> 4) : 0;
        for (int i = 0; i < n_mma; i += 6) {
#pragma unroll
            for (int u = 0; u < 6; ++u) {
                if (!uniform || elect_one()) {
                    umma_bf16(tmem, ad0 + (uint64_t)(u * a_step), bd0 + (uint64_t)(u * b_step), idesc, (i + u) > 0 ? 1u : 0u);
                    if (commit_every > 0 && (u + 1) % commit_every == 0) umma_commit(smem_u32(&bar[1]));
                }
                if (uniform) __syncwarp();
            }
        }
        if (!uniform || elect_one()) umma_commit(smem_u32(&bar[0]));
        long long t1 = clock64();
        mbar_wait(smem_u32(&bar[0]), ph);
        long long t2 = clock64();
        if (tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int CE, int U>
void run(long long* d, int N, int rotate) {
    const int n_mma = 3000;
    cudaFuncSetAttribute(probe<CE, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int rep = 0; rep < 2; ++rep) probe<CE, U><<<1, 128, 196 * 1024>>>(N, n_mma, rotate, d);
    long long h[2];
    cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    printf("N=%3d commit_every=%d rotate=%d uniform=%d : issue %.1f cyc/mma, complete %.1f cyc/mma\n", N, CE, rotate, U,
           (double)h[0] / n_mma, (double)h[1] / n_mma);
}

int main() {
    long long* d; cudaMalloc(&d, 16);
    for (int N : {256, 128, 64}) {
        run<0, 0>(d, N, 1); run<6, 0>(d, N, 1); run<3, 0>(d, N, 1); run<1, 0>(d, N, 1);
        run<0, 1>(d, N, 1); run<3, 1>(d, N, 1);
    }
    return 0;
}
