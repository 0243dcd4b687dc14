// Descriptor-similarity features of CoarseReg (reference models/HRegNet/layers.py:29-41, 290-313, 339-362).
//
// The reference materialises two [B,N2,N1,C] `repeat`ed tensors (2 x 2.1 GB at B=32), reduces them
// elementwise into cos[B,N2,N1], normalises by row / column maxima, gathers [B,N1,k,N1] slabs with knn_gather
// and finally picks the diagonal with 4 x N1 Python-loop iterations.  Here:
//   hrn_cosine_matrix : cos[b,n2,n1] = <D[n2],S[n1]> / (|D[n2]| |S[n1]| + 1e-6)   -- one dense contraction
//   hrn_cosine_maxima : rowmax[b,n2] = max_n1 cos,  colmax[b,n1] = max_n2 cos
//   hrn_cosine_pick   : src_dst[b,i,j] = cos[b,idx,i]/(colmax[b,i]+1e-6),  dst_src[b,i,j] = cos[b,idx,i]/(rowmax[b,idx]+1e-6)
#include "common.cuh"
#include <math_constants.h>

namespace {

// norms[r] = sqrt(sum_c x[r,c]^2); one warp per row
__global__ void row_norm_kernel(const float* __restrict__ x, int C, long long rows, float* __restrict__ out) {
    const long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) { const float v = x[r * C + c]; acc = fmaf(v, v, acc); }
    acc = hrn_warp_sum(acc);
    if (lane == 0) out[r] = sqrtf(acc);
}

// 64x64 output tile per CTA (256 threads, 4x4 per thread), K-chunks of 16 through shared memory.
__global__ void __launch_bounds__(256)
cosine_matrix_kernel(const float* __restrict__ S, const float* __restrict__ Dd, const float* __restrict__ nS,
                     const float* __restrict__ nD, float* __restrict__ cosm, int N1, int N2, int C) {
    __shared__ __align__(16) float As[16][68];  // D rows (n2)
    __shared__ __align__(16) float Bs[16][68];  // S rows (n1)
    const int b = blockIdx.z;
    S += (size_t)b * N1 * C; Dd += (size_t)b * N2 * C; nS += (size_t)b * N1; nD += (size_t)b * N2;
    cosm += (size_t)b * N2 * N1;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, kk = tid & 15, rr = tid >> 4;
    const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < C; k0 += 16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rl = rr + 16 * i;
            As[kk][rl] = (r0 + rl < N2 && k0 + kk < C) ? Dd[(size_t)(r0 + rl) * C + k0 + kk] : 0.f;
            Bs[kk][rl] = (c0 + rl < N1 && k0 + kk < C) ? S[(size_t)(c0 + rl) * C + k0 + kk] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float4 a = *reinterpret_cast<const float4*>(&As[q][ty * 4]);
            const float4 w = *reinterpret_cast<const float4*>(&Bs[q][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n2 = r0 + ty * 4 + i;
        if (n2 >= N2) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n1 = c0 + tx * 4 + j;
            if (n1 < N1) cosm[(size_t)n2 * N1 + n1] = acc[i][j] / (nD[n2] * nS[n1] + 1e-6f);
        }
    }
}

// grid (N2 + N1, B): blocks [0,N2) reduce a row, blocks [N2, N2+N1) reduce a column
__global__ void __launch_bounds__(128)
cosine_maxima_kernel(const float* __restrict__ cosm, int N1, int N2, float* __restrict__ rowmax,
                     float* __restrict__ colmax) {
    __shared__ float s_part[4];
    const int b = blockIdx.y;
    cosm += (size_t)b * N2 * N1;
    float m = -CUDART_INF_F;
    const int id = blockIdx.x;
    if (id < N2) {
        for (int n1 = threadIdx.x; n1 < N1; n1 += blockDim.x) m = fmaxf(m, cosm[(size_t)id * N1 + n1]);
    } else {
        const int n1 = id - N2;
        for (int n2 = threadIdx.x; n2 < N2; n2 += blockDim.x) m = fmaxf(m, cosm[(size_t)n2 * N1 + n1]);
    }
    m = hrn_warp_max(m);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        m = fmaxf(fmaxf(s_part[0], s_part[1]), fmaxf(s_part[2], s_part[3]));
        if (id < N2) rowmax[(size_t)b * N2 + id] = m; else colmax[(size_t)b * N1 + (id - N2)] = m;
    }
}

__global__ void cosine_pick_kernel(const float* __restrict__ cosm, const float* __restrict__ rowmax,
                                   const float* __restrict__ colmax, const int32_t* __restrict__ idx, long long rows,
                                   int N1, int N2, int k, float* __restrict__ out, int ldo, int col_src_dst,
                                   int col_dst_src) {
    const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const long long bi = r / k;
    const long long b = bi / N1;
    const int i = (int)(bi - b * N1);
    const int n2 = idx[r];
    const float c = cosm[((size_t)b * N2 + n2) * N1 + i];
    out[r * ldo + col_src_dst] = c / (colmax[b * N1 + i] + 1e-6f);
    out[r * ldo + col_dst_src] = c / (rowmax[b * N2 + n2] + 1e-6f);
}

}  // namespace

// S [B,N1,C] source descriptors, D [B,N2,C] target descriptors (channels-last) -> cosm [B,N2,N1];
// scratch norms nS [B,N1], nD [B,N2]; rowmax [B,N2], colmax [B,N1].
HRN_API int hrn_cosine_matrix(const float* S, const float* D, int B, int N1, int N2, int C, float* nS, float* nD,
                              float* cosm, float* rowmax, float* colmax, void* stream) {
    if (!S || !D || !nS || !nD || !cosm || !rowmax || !colmax || B < 0 || N1 <= 0 || N2 <= 0 || C <= 0) return HRN_ERR_BAD_ARG;
    if (B == 0) return HRN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    row_norm_kernel<<<hrn_divup((long long)B * N1 * 32, 256), 256, 0, st>>>(S, C, (long long)B * N1, nS);
    row_norm_kernel<<<hrn_divup((long long)B * N2 * 32, 256), 256, 0, st>>>(D, C, (long long)B * N2, nD);
    dim3 grid(hrn_divup(N1, 64), hrn_divup(N2, 64), B);
    cosine_matrix_kernel<<<grid, 256, 0, st>>>(S, D, nS, nD, cosm, N1, N2, C);
    dim3 g2(N1 + N2, B);
    cosine_maxima_kernel<<<g2, 128, 0, st>>>(cosm, N1, N2, rowmax, colmax);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}

// idx [B,N1,k] int32 (neighbours in the target cloud); writes two columns of out [B*N1*k, ldo].
HRN_API int hrn_cosine_pick(const float* cosm, const float* rowmax, const float* colmax, const int32_t* idx, int B,
                            int N1, int N2, int k, float* out, int ldo, int col_src_dst, int col_dst_src,
                            void* stream) {
    if (!cosm || !rowmax || !colmax || !idx || !out || B < 0 || N1 <= 0 || N2 <= 0 || k <= 0) return HRN_ERR_BAD_ARG;
    const long long rows = (long long)B * N1 * k;
    if (rows == 0) return HRN_OK;
    cosine_pick_kernel<<<hrn_divup(rows, 256), 256, 0, (cudaStream_t)stream>>>(cosm, rowmax, colmax, idx, rows, N1, N2,
                                                                              k, out, ldo, col_src_dst, col_dst_src);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}
