// Per-neighbourhood ("group" = the k rows of one keypoint) reductions and feature assembly, channels-last.
//
//   hrn_group_attention   <- layers.py:151-152, 330-331, 385-386, 447-448:  a = softmax_k( max_c E )
//   hrn_group_weighted_sum<- layers.py:154-159, 332, 388-390, 449-450:      out[g,:] = sum_j a[g,j] * V[row(g,j),:]
//   hrn_group_max         <- layers.py:202, 208:                            out[g,:] = max_j X[g*k+j,:]
//   hrn_group_geometry    <- layers.py:20-23 (knn_group), 284-288, 318-319, 364-365, 438-445:
//                            the small "geometry + weights" channels of every grouped feature tensor
//   hrn_sigma_to_weights  <- models.py:30-32:  w = 1/(sigma+1e-5); w /= mean(w)
//   hrn_transform_points  <- models.py:91-92,113-114:  x' = R x + t
// All are HBM-bound single-pass kernels.
#include "common.cuh"
#include <math_constants.h>

namespace {

// one CTA (128 threads) per group; E [G*k, C] (ld = ldE)
__global__ void __launch_bounds__(128)
group_attention_kernel(const float* __restrict__ E, int ldE, int C, int k, float* __restrict__ a) {
    __shared__ float s_x[64];
    const long long g = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = warp; j < k; j += 4) {
        const float* row = E + (g * k + j) * ldE;
        float m = -CUDART_INF_F;
        for (int c = lane; c < C; c += 32) m = fmaxf(m, row[c]);
        m = hrn_warp_max(m);
        if (lane == 0) s_x[j] = m;
    }
    __syncthreads();
    if (warp == 0) {
        const float v0 = lane < k ? s_x[lane] : -CUDART_INF_F;
        const float v1 = lane + 32 < k ? s_x[lane + 32] : -CUDART_INF_F;
        const float mx = hrn_warp_max(fmaxf(v0, v1));
        const float e0 = lane < k ? expf(v0 - mx) : 0.f;
        const float e1 = lane + 32 < k ? expf(v1 - mx) : 0.f;
        const float sum = hrn_warp_sum(e0 + e1);
        if (lane < k) a[g * k + lane] = e0 / sum;
        if (lane + 32 < k) a[g * k + lane + 32] = e1 / sum;
    }
}

// out[g, c] = sum_j a[g*k+j] * V[src(g,j), c];  src = direct row g*k+j, or b*N + idx[g*k+j]
__global__ void __launch_bounds__(128)
group_weighted_sum_kernel(const float* __restrict__ a, const float* __restrict__ V, int ldV, int C, int k,
                          const int32_t* __restrict__ idx, int groups_per_batch, int N, float* __restrict__ out,
                          int ldo) {
    const long long g = blockIdx.x;
    const long long b = g / groups_per_batch;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float acc = 0.f;
        for (int j = 0; j < k; ++j) {
            const long long r = g * k + j;
            const long long sr = idx ? b * N + idx[r] : r;
            acc = fmaf(a[r], V[sr * ldV + c], acc);
        }
        out[g * ldo + c] = acc;
    }
}

// The same for C % 4 == 0 and 16-byte aligned rows: C/4 lanes per group, each with a float4 of channels (a quarter of the
// load instructions), several groups per 128-thread CTA; narrow outputs (the 3 correspondence coordinates) go through the
// third kernel, which packs 128 (group, channel) pairs into a CTA instead of one group per 32-thread CTA.  Per output the
// fma chain over j is unchanged: same bits.
__global__ void __launch_bounds__(128)
group_weighted_sum_vec4_kernel(const float* __restrict__ a, const float* __restrict__ V, int ldV, int C4, int k,
                               const int32_t* __restrict__ idx, int groups_per_batch, int N, float* __restrict__ out,
                               int ldo, long long groups) {
    const int per_cta = 128 / C4;                           // groups per CTA (C4 <= 128 divides 128)
    const long long g = (long long)blockIdx.x * per_cta + threadIdx.x / C4;
    const int c = threadIdx.x % C4;
    if (g >= groups) return;
    const long long b = g / groups_per_batch;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < k; ++j) {
        const long long r = g * k + j;
        const long long sr = idx ? b * N + __ldg(idx + r) : r;
        const float w = __ldg(a + r);
        const float4 v = __ldg(reinterpret_cast<const float4*>(V + sr * ldV) + c);
        acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
    }
    *reinterpret_cast<float4*>(out + g * ldo + 4 * c) = acc;
}

__global__ void __launch_bounds__(128)
group_weighted_sum_narrow_kernel(const float* __restrict__ a, const float* __restrict__ V, int ldV, int C, int k,
                                 const int32_t* __restrict__ idx, int groups_per_batch, int N, float* __restrict__ out,
                                 int ldo, long long groups) {
    const long long t = (long long)blockIdx.x * 128 + threadIdx.x;
    const long long g = t / C;
    const int c = (int)(t - g * C);
    if (g >= groups) return;
    const long long b = g / groups_per_batch;
    float acc = 0.f;
    for (int j = 0; j < k; ++j) {
        const long long r = g * k + j;
        const long long sr = idx ? b * N + __ldg(idx + r) : r;
        acc = fmaf(__ldg(a + r), __ldg(V + sr * ldV + c), acc);
    }
    out[g * ldo + c] = acc;
}

// Attention tail of the correspondence heads in ONE pass over the k rows of a group (layers.py:385-390, 447-450):
//   a = softmax_k(max_c E), af[g,:] = sum_j a_j E[g*k+j,:], cor[g,:] = sum_j a_j xyz[b*N + idx[g*k+j],:]
// The rows are read once (float4, coalesced) into shared memory while their maxima are taken; the three separate
// kernels read the [rows, C] tensor twice (134 MB each at the coarse level).  Same operation order as
// group_attention_kernel / group_weighted_sum_kernel, hence the same bits.
__global__ void __launch_bounds__(128)
group_attend_kernel(const float* __restrict__ E, int ldE, int C, int k, float* __restrict__ a_out,
                    float* __restrict__ af, int ldaf, const float* __restrict__ xyz, const int32_t* __restrict__ idx,
                    int groups_per_batch, int N, float* __restrict__ cor) {
    extern __shared__ __align__(16) float s_E[];            // [k][C]
    __shared__ float s_x[64], s_a[64];
    const long long g = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C4 = C >> 2;
    for (int j = warp; j < k; j += 4) {
        const float4* row = reinterpret_cast<const float4*>(E + (g * k + j) * ldE);
        float4* dst = reinterpret_cast<float4*>(s_E + j * C);
        float m = -CUDART_INF_F;
        for (int c = lane; c < C4; c += 32) {
            const float4 v = __ldg(row + c);
            dst[c] = v;
            m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
        }
        m = hrn_warp_max(m);
        if (lane == 0) s_x[j] = m;
    }
    __syncthreads();
    if (warp == 0) {
        const float v0 = lane < k ? s_x[lane] : -CUDART_INF_F;
        const float v1 = lane + 32 < k ? s_x[lane + 32] : -CUDART_INF_F;
        const float mx = hrn_warp_max(fmaxf(v0, v1));
        const float e0 = lane < k ? expf(v0 - mx) : 0.f;
        const float e1 = lane + 32 < k ? expf(v1 - mx) : 0.f;
        const float sum = hrn_warp_sum(e0 + e1);
        if (lane < k) { s_a[lane] = e0 / sum; if (a_out) a_out[g * k + lane] = e0 / sum; }
        if (lane + 32 < k) { s_a[lane + 32] = e1 / sum; if (a_out) a_out[g * k + lane + 32] = e1 / sum; }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float acc = 0.f;
        for (int j = 0; j < k; ++j) acc = fmaf(s_a[j], s_E[j * C + c], acc);
        af[g * ldaf + c] = acc;
    }
    if (cor && threadIdx.x < 3) {
        const long long b = g / groups_per_batch;
        float acc = 0.f;
        for (int j = 0; j < k; ++j) acc = fmaf(s_a[j], xyz[(b * N + idx[g * k + j]) * 3 + threadIdx.x], acc);
        cor[g * 3 + threadIdx.x] = acc;
    }
}
__global__ void __launch_bounds__(128)
group_max_kernel(const float* __restrict__ X, int ldX, int C, int k, float* __restrict__ out, int ldo) {
    const long long g = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float m = -CUDART_INF_F;
        for (int j = 0; j < k; ++j) m = fmaxf(m, X[(g * k + j) * ldX + c]);
        out[g * ldo + c] = m;
    }
}

// One thread per row r = (b, m, j).  q [B,M,3] queries, p [B,N,3] references, idx [B,M,k] int32.
// Writes `out` [rows, ldo] columns:
//   0..2 rel = p[idx] - q ; 3 = |rel| ;
//   if (pair): 4..6 q ; 7..9 p[idx] ; 10 = wq[b,m] ; 11 = wp[b,idx]
// Optionally nn [rows,3] = p[idx].
__global__ void group_geometry_kernel(const float* __restrict__ q, const float* __restrict__ p,
                                      const int32_t* __restrict__ idx, const float* __restrict__ wq,
                                      const float* __restrict__ wp, int pair, long long rows, int M, int k, int N,
                                      float* __restrict__ out, int ldo, float* __restrict__ nn) {
    const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const long long bm = r / k;
    const long long b = bm / M;
    const int n = idx[r];
    const float* qq = q + bm * 3;
    const float* pp = p + (b * N + n) * 3;
    const float qx = qq[0], qy = qq[1], qz = qq[2], x = pp[0], y = pp[1], z = pp[2];
    const float rx = x - qx, ry = y - qy, rz = z - qz;
    float* o = out + r * ldo;
    o[0] = rx; o[1] = ry; o[2] = rz;
    o[3] = sqrtf(rx * rx + ry * ry + rz * rz);
    if (pair) {
        o[4] = qx; o[5] = qy; o[6] = qz;
        o[7] = x; o[8] = y; o[9] = z;
        o[10] = wq[bm];
        o[11] = wp[b * N + n];
    }
    if (nn) { nn[r * 3 + 0] = x; nn[r * 3 + 1] = y; nn[r * 3 + 2] = z; }
}

// w[b,:] = (1/(sigma+1e-5)) / mean_m(1/(sigma+1e-5));  one CTA per cloud (reference models.py:30-32).
// The mean is accumulated in fp64 and rounded once, i.e. it is the correctly rounded fp32 mean of the fp32 terms --
// independent of any summation order.  torch.mean (the reference) sums in fp32 with its own vectorised order and can
// differ from that value by an ulp; a 1-ulp difference of the weights can flip a weighted-FPS pick downstream, which is
// why the index gates are applied teacher-forced.
__global__ void __launch_bounds__(256)
sigma_to_weights_kernel(const float* __restrict__ sigma, float* __restrict__ w, int M) {
    __shared__ double s_part[8];
    const float* s = sigma + (size_t)blockIdx.x * M;
    float* o = w + (size_t)blockIdx.x * M;
    double acc = 0.0;
    for (int i = threadIdx.x; i < M; i += blockDim.x) acc += (double)(1.0f / (s[i] + 1e-5f));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    double tot = 0.0;
    for (int i = 0; i < 8; ++i) tot += s_part[i];
    const float mean = (float)(tot / (double)M);
    for (int i = threadIdx.x; i < M; i += blockDim.x) o[i] = (1.0f / (s[i] + 1e-5f)) / mean;
}

// out[b,n,:] = R[b] x[b,n,:] + t[b]
__global__ void transform_points_kernel(const float* __restrict__ x, const float* __restrict__ R,
                                        const float* __restrict__ t, float* __restrict__ out, long long total, int N) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long b = i / N;
    const float* Rb = R + b * 9;
    const float* tb = t + b * 3;
    const float px = x[i * 3 + 0], py = x[i * 3 + 1], pz = x[i * 3 + 2];
    out[i * 3 + 0] = fmaf(Rb[2], pz, fmaf(Rb[1], py, Rb[0] * px)) + tb[0];
    out[i * 3 + 1] = fmaf(Rb[5], pz, fmaf(Rb[4], py, Rb[3] * px)) + tb[1];
    out[i * 3 + 2] = fmaf(Rb[8], pz, fmaf(Rb[7], py, Rb[6] * px)) + tb[2];
}

}  // namespace

HRN_API int hrn_group_attention(const float* E, int ldE, int C, long long groups, int k, float* a, void* stream) {
    if (!E || !a || C <= 0 || k <= 0 || k > 64 || groups < 0) return HRN_ERR_BAD_ARG;
    if (groups == 0) return HRN_OK;
    group_attention_kernel<<<(unsigned)groups, 128, 0, (cudaStream_t)stream>>>(E, ldE, C, k, a);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}

HRN_API int hrn_group_weighted_sum(const float* a, const float* V, int ldV, int C, long long groups, int k,
                                   const int32_t* idx, int groups_per_batch, int N, float* out, int ldo,
                                   void* stream) {
    if (!a || !V || !out || C <= 0 || k <= 0 || groups < 0) return HRN_ERR_BAD_ARG;
    if (groups == 0) return HRN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int C4 = C >> 2;
    if ((C & 3) == 0 && C4 <= 128 && 128 % C4 == 0 && (ldV & 3) == 0 && (ldo & 3) == 0 && ((uintptr_t)V & 15) == 0 &&
        ((uintptr_t)out & 15) == 0) {
        const int per_cta = 128 / C4;
        group_weighted_sum_vec4_kernel<<<(unsigned)((groups + per_cta - 1) / per_cta), 128, 0, st>>>(
            a, V, ldV, C4, k, idx, groups_per_batch, N, out, ldo, groups);
        HRN_LAUNCH_CHECK();
        return HRN_OK;
    }
    if (C <= 8) {
        group_weighted_sum_narrow_kernel<<<(unsigned)((groups * C + 127) / 128), 128, 0, st>>>(
            a, V, ldV, C, k, idx, groups_per_batch, N, out, ldo, groups);
        HRN_LAUNCH_CHECK();
        return HRN_OK;
    }
    const int threads = C >= 128 ? 128 : (C >= 64 ? 64 : 32);
    group_weighted_sum_kernel<<<(unsigned)groups, threads, 0, (cudaStream_t)stream>>>(a, V, ldV, C, k, idx,
                                                                                     groups_per_batch, N, out, ldo);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}

HRN_API int hrn_group_attend(const float* E, int ldE, int C, long long groups, int k, float* a, float* af, int ldaf,
                             const float* xyz, const int32_t* idx, int groups_per_batch, int N, float* cor,
                             void* stream) {
    if (!E || !af || C <= 0 || k <= 0 || k > 64 || groups < 0) return HRN_ERR_BAD_ARG;
    if (cor && (!xyz || !idx || groups_per_batch <= 0 || N <= 0)) return HRN_ERR_BAD_ARG;
    if ((C & 3) || (ldE & 3) || ((uintptr_t)E & 15)) return HRN_ERR_UNSUPPORTED;
    const size_t smem = (size_t)k * C * sizeof(float);
    if (smem > 96 * 1024) return HRN_ERR_UNSUPPORTED;
    if (groups == 0) return HRN_OK;
    static hrn_once_per_device attr_set;
    if (attr_set.need()) {
        HRN_CUDA(cudaFuncSetAttribute(group_attend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    }
    group_attend_kernel<<<(unsigned)groups, 128, smem, (cudaStream_t)stream>>>(E, ldE, C, k, a, af, ldaf, xyz, idx,
                                                                              groups_per_batch, N, cor);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}
HRN_API int hrn_group_max(const float* X, int ldX, int C, long long groups, int k, float* out, int ldo,
                          void* stream) {
    if (!X || !out || C <= 0 || k <= 0 || groups < 0) return HRN_ERR_BAD_ARG;
    if (groups == 0) return HRN_OK;
    const int threads = C >= 128 ? 128 : (C >= 64 ? 64 : 32);
    group_max_kernel<<<(unsigned)groups, threads, 0, (cudaStream_t)stream>>>(X, ldX, C, k, out, ldo);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}

HRN_API int hrn_group_geometry(const float* q, const float* p, const int32_t* idx, const float* wq, const float* wp,
                               int B, int M, int k, int N, float* out, int ldo, float* nn, void* stream) {
    if (!q || !p || !idx || !out || B < 0 || M < 0 || k <= 0 || N <= 0) return HRN_ERR_BAD_ARG;
    const int pair = (wq && wp) ? 1 : 0;
    if (ldo < (pair ? 12 : 4)) return HRN_ERR_BAD_ARG;
    const long long rows = (long long)B * M * k;
    if (rows == 0) return HRN_OK;
    group_geometry_kernel<<<hrn_divup(rows, 256), 256, 0, (cudaStream_t)stream>>>(q, p, idx, wq, wp, pair, rows, M, k,
                                                                                 N, out, ldo, nn);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}

HRN_API int hrn_sigma_to_weights(const float* sigma, float* w, int B, int M, void* stream) {
    if (!sigma || !w || B < 0 || M <= 0) return HRN_ERR_BAD_ARG;
    if (B == 0) return HRN_OK;
    sigma_to_weights_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(sigma, w, M);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}

HRN_API int hrn_transform_points(const float* x, const float* R, const float* t, float* out, int B, int N,
                                 void* stream) {
    if (!x || !R || !t || !out || B < 0 || N < 0) return HRN_ERR_BAD_ARG;
    const long long total = (long long)B * N;
    if (total == 0) return HRN_OK;
    transform_points_kernel<<<hrn_divup(total, 256), 256, 0, (cudaStream_t)stream>>>(x, R, t, out, total, N);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}
