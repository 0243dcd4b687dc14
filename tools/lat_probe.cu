// Micro-probe: dependent-chain latency (cycles/op) of the warp-collective primitives used on the FPS critical path.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 lat_probe.cu -o lat_probe
#include <cstdio>
#include <cuda_runtime.h>
#define CHAIN 64
template <int OP>
__global__ void probe(unsigned* out, long long* cyc, unsigned seed) {
    __shared__ unsigned s[1024];
    const int tid = threadIdx.x, lane = tid & 31;
    s[tid] = tid * 2654435761u + seed;
    __syncthreads();
    unsigned v = s[tid] ^ seed;
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int i = 0; i < CHAIN / 4; ++i) {
            if (OP == 0) v = __reduce_max_sync(0xffffffffu, v + lane);
            if (OP == 1) { v += lane; for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o)); }
            if (OP == 2) v = __ballot_sync(0xffffffffu, (v + lane) & 1) + v;
            if (OP == 3) v = s[(v + lane) & 1023];
            if (OP == 4) v = __shfl_sync(0xffffffffu, v + 1, (v + lane) & 31);
            if (OP == 5) v = __reduce_min_sync(0xffffffffu, v ^ lane) + 3;
            if (OP == 6) v = __float_as_uint(fmaxf(__uint_as_float(v & 0x3fffffffu), 1.0f)) + 1;
            if (OP == 7) { v += lane; v = max(v, __shfl_xor_sync(0xffffffffu, v, 16)); }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + tid] = v;
    if (tid == 0 && blockIdx.x == 0) cyc[OP] = t1 - t0;
}
int main() {
    unsigned* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 64);
    const char* names[] = {"redux.max u32", "5x shfl_xor max", "ballot", "lds", "shfl idx", "redux.min u32", "fmnmx", "1x shfl_xor max"};
    for (int threads : {32, 512, 1024}) {
        probe<0><<<1, threads>>>(out, cyc, 1); probe<1><<<1, threads>>>(out, cyc, 1); probe<2><<<1, threads>>>(out, cyc, 1);
        probe<3><<<1, threads>>>(out, cyc, 1); probe<4><<<1, threads>>>(out, cyc, 1); probe<5><<<1, threads>>>(out, cyc, 1);
        probe<6><<<1, threads>>>(out, cyc, 1); probe<7><<<1, threads>>>(out, cyc, 1);
        long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        printf("threads/CTA=%d:", threads);
        for (int i = 0; i < 8; ++i) printf("  %s %.1f", names[i], (double)h[i] / CHAIN);
        printf("\n");
    }
    return 0;
}
