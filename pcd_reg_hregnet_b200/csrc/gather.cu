// Index gathers of the HRegNet forward path (HBM-bound; coalesced on the contiguous axis).
//
//   hrn_gather_points       <- reference furthest_point_sampling_gpu.cu:7-39  (out[b,c,m] = points[b,c,idx[b,m]])
//   hrn_gather_points_grad  <- reference furthest_point_sampling_gpu.cu:41-73 (atomic scatter-add, backward only)
//   hrn_gather_rows         <- the `gather_operation(xyz^T, idx)^T` idiom of layers.py:140,143 without the two
//                              permute+contiguous passes: out[b,m,:] = x[b,idx[b,m],:] (row gather, int32 idx)
//   hrn_knn_gather          <- pytorch3d.ops.knn_gather (call sites layers.py:25,279,288,303,...): int64 idx
//   hrn_knn_gather_grad, hrn_knn_dists_grad <- the backward passes of knn_gather / knn_points that pytorch3d's autograd
//                              functions provide to the reference's training scripts (train/train_reg_v0.py:281-296)
//   hrn_transpose_bcn_bnc   <- the permute(0,2,1).contiguous() glue between the [B,C,N] API layout and the
//                              channels-last rows the kernels consume (tiled through shared memory)
#include "common.cuh"

namespace {

__global__ void gather_points_kernel(const float* __restrict__ points, const int32_t* __restrict__ idx,
                                     float* __restrict__ out, int C, int N, int M) {
    const int b = blockIdx.z, c = blockIdx.y;
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    out[((size_t)b * C + c) * M + m] = __ldg(points + ((size_t)b * C + c) * N + idx[(size_t)b * M + m]);
}

__global__ void gather_points_grad_kernel(const float* __restrict__ grad_out, const int32_t* __restrict__ idx,
                                          float* __restrict__ grad_points, int C, int N, int M) {
    const int b = blockIdx.z, c = blockIdx.y;
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    atomicAdd(grad_points + ((size_t)b * C + c) * N + idx[(size_t)b * M + m], grad_out[((size_t)b * C + c) * M + m]);
}

// out[(b*M + m)*U + u] = x[(b*N + idx[b*M+m])*U + u]; one thread per output float (coalesced over u).
template <typename IdxT>
__global__ void gather_rows_kernel(const float* __restrict__ x, const IdxT* __restrict__ idx, float* __restrict__ out,
                                   long long total, int rows_per_batch, int N, int U) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / U;
        const int u = (int)(i - r * U);
        const long long b = r / rows_per_batch;
        out[i] = __ldg(x + ((size_t)b * N + (size_t)idx[r]) * U + u);
    }
}

// Backward of knn_gather: grad_x[b, idx[b,m,k], u] += grad_out[b,m,k,u]  (grad_x pre-zeroed by the caller; fp32 atomics,
// like the reference's own gather gradient, furthest_point_sampling_gpu.cu:41-55)
__global__ void knn_gather_grad_kernel(const float* __restrict__ grad_out, const int64_t* __restrict__ idx,
                                       float* __restrict__ grad_x, long long total, int rows_per_batch, int N, int U) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / U;
        const int u = (int)(i - r * U);
        const long long b = r / rows_per_batch;
        atomicAdd(grad_x + ((size_t)b * N + (size_t)idx[r]) * U + u, grad_out[i]);
    }
}

// Backward of knn_points' squared distances  d[b,m,k] = sum_j (p1[b,m,j] - p2[b,idx,j])^2 :
//   grad_p1[b,m,j]   = sum_k 2 g[b,m,k] (p1 - p2[idx])          (one thread per (b,m,j): plain sum over k)
//   grad_p2[b,idx,j] -=       2 g[b,m,k] (p1 - p2[idx])          (atomics; pre-zeroed by the caller)
__global__ void knn_dists_grad_kernel(const float* __restrict__ p1, const float* __restrict__ p2,
                                      const int64_t* __restrict__ idx, const float* __restrict__ g,
                                      float* __restrict__ grad_p1, float* __restrict__ grad_p2, long long total, int M,
                                      int N, int D, int K) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long bm = i / D;
        const int j = (int)(i - bm * D);
        const long long b = bm / M;
        const float q = p1[i];
        float acc = 0.f;
        for (int k = 0; k < K; ++k) {
            const size_t n = (size_t)b * N + (size_t)idx[bm * K + k];
            const float t = 2.f * g[bm * K + k] * (q - p2[n * D + j]);
            acc += t;
            if (grad_p2) atomicAdd(grad_p2 + n * D + j, -t);
        }
        if (grad_p1) grad_p1[i] = acc;
    }
}

// [B, R, Cc] -> [B, Cc, R]  (32x32 tiles through padded shared memory; both sides coalesced)
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int Cc) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    in += (size_t)b * R * Cc;
    out += (size_t)b * R * Cc;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        if (r < R && c < Cc) tile[i][threadIdx.x] = in[(size_t)r * Cc + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < R && c < Cc) out[(size_t)c * R + r] = tile[threadIdx.x][i];
    }
}

}  // namespace

HRN_API int hrn_gather_points(const float* points, const int32_t* idx, float* out, int B, int C, int N, int M,
                              void* stream) {
    if (!points || !idx || !out || B < 0 || C < 0 || N <= 0 || M < 0) return HRN_ERR_BAD_ARG;
    if (B == 0 || C == 0 || M == 0) return HRN_OK;
    dim3 grid(hrn_divup(M, 256), C, B);
    gather_points_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(points, idx, out, C, N, M);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}

HRN_API int hrn_gather_points_grad(const float* grad_out, const int32_t* idx, float* grad_points, int B, int C, int N,
                                   int M, void* stream) {
    if (!grad_out || !idx || !grad_points || B < 0 || C < 0 || N <= 0 || M < 0) return HRN_ERR_BAD_ARG;
    if (B == 0 || C == 0 || M == 0) return HRN_OK;
    dim3 grid(hrn_divup(M, 256), C, B);
    gather_points_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(grad_out, idx, grad_points, C, N, M);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}

HRN_API int hrn_gather_rows(const float* x, const int32_t* idx, float* out, int B, int N, int M, int U, void* stream) {
    if (!x || !idx || !out || B < 0 || N <= 0 || M < 0 || U <= 0) return HRN_ERR_BAD_ARG;
    const long long total = (long long)B * M * U;
    if (total == 0) return HRN_OK;
    const int blocks = (int)((total + 255) / 256 < 148LL * 16 ? (total + 255) / 256 : 148LL * 16);
    gather_rows_kernel<int32_t><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, idx, out, total, M, N, U);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}

HRN_API int hrn_knn_gather(const float* x, const int64_t* idx, float* out, int B, int N, int M, int K, int U,
                           void* stream) {
    if (!x || !idx || !out || B < 0 || N <= 0 || M < 0 || K < 0 || U <= 0) return HRN_ERR_BAD_ARG;
    const long long total = (long long)B * M * K * U;
    if (total == 0) return HRN_OK;
    const int blocks = (int)((total + 255) / 256 < 148LL * 16 ? (total + 255) / 256 : 148LL * 16);
    gather_rows_kernel<int64_t><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, idx, out, total, M * K, N, U);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}

HRN_API int hrn_knn_gather_grad(const float* grad_out, const int64_t* idx, float* grad_x, int B, int N, int M, int K,
                                int U, void* stream) {
    if (!grad_out || !idx || !grad_x || B < 0 || N <= 0 || M < 0 || K < 0 || U <= 0) return HRN_ERR_BAD_ARG;
    const long long total = (long long)B * M * K * U;
    if (total == 0) return HRN_OK;
    const int blocks = (int)((total + 255) / 256 < 148LL * 16 ? (total + 255) / 256 : 148LL * 16);
    knn_gather_grad_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(grad_out, idx, grad_x, total, M * K, N, U);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}

HRN_API int hrn_knn_dists_grad(const float* p1, const float* p2, const int64_t* idx, const float* grad_dists,
                               float* grad_p1, float* grad_p2, int B, int M, int N, int D, int K, void* stream) {
    if (!p1 || !p2 || !idx || !grad_dists || (!grad_p1 && !grad_p2) || B < 0 || M < 0 || N <= 0 || D <= 0 || K <= 0)
        return HRN_ERR_BAD_ARG;
    const long long total = (long long)B * M * D;
    if (total == 0) return HRN_OK;
    const int blocks = (int)((total + 255) / 256 < 148LL * 16 ? (total + 255) / 256 : 148LL * 16);
    knn_dists_grad_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p1, p2, idx, grad_dists, grad_p1, grad_p2, total, M, N,
                                                                    D, K);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}

HRN_API int hrn_transpose(const float* in, float* out, int B, int R, int Cc, void* stream) {
    if (!in || !out || B < 0 || R < 0 || Cc < 0) return HRN_ERR_BAD_ARG;
    if (B == 0 || R == 0 || Cc == 0) return HRN_OK;
    dim3 grid(hrn_divup(Cc, 32), hrn_divup(R, 32), B), block(32, 8);
    transpose_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(in, out, R, Cc);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}
