// Weighted Kabsch / SVD pose head, fused: weights normalisation, weighted means, 3x3 weighted covariance,
// 3x3 SVD (Jacobi, fp64) and R|t -- one CTA per registration pair, nothing but R (9 floats) and t (3 floats)
// leaves the SM.
//
// Replaces reference models/HRegNet/layers.py:469-504 (WeightedSVDHead.forward): ~12 ATen launches, a
// diag_embed [B,N,N] matrix (134 MB at B=32, N=1024), two bmm and a cuSOLVER batched SVD; and the 4x4
// homogeneous pose composition of models/HRegNet/models.py:100-110,120-127 (optional `prev` pose).
//
// Formula kept identical to the reference (eps = 1e-4 appears twice, layers.py:470-476):
//   w' = w / (sum w + eps);  xbar = (sum w' x) / (sum w' + eps);  ybar likewise
//   H = sum w' (x - xbar)(y - ybar)^T ;  H = U S V^T ;  R = V diag(1,1,det(V U^T)) U^T ;  t = ybar - R xbar
// R is evaluated as [v0 v1 v0xv1][u0 u1 u0xu1]^T with u_i = H v_i / s_i, which equals the expression above for
// either sign of det(H) and never divides by the smallest singular value.  Sums are accumulated in fp64
// (inputs stay the caller's fp32), so the result sits closer to the exact answer than the reference's own
// fp32 evaluation ("pose tolerance").  Rank-deficient H (s1 ~ 0) or non-finite input -> R = I,
// t = 0, the reference's SVD-failure fallback (layers.py:487-493).
#include "common.cuh"

namespace {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* s_red /*[8][NV]*/, double (&out)[NV]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum_d(v[i]);
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < NV; ++i) s_red[warp * NV + i] = v[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double a = 0.0;
        for (int w = 0; w < 8; ++w) a += s_red[w * NV + i];
        out[i] = a;
    }
}

// cyclic Jacobi on a symmetric 3x3 (A), accumulates eigenvectors in V (columns)
__host__ __device__ inline void jacobi_eig3(double A[3][3], double V[3][3]) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        const double diag = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
        if (off <= 1e-300 || off <= 1e-17 * diag) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (A[p][q] == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;
                for (int k = 0; k < 3; ++k) {  // A <- A J
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {  // A <- J^T A
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
                }
            }
    }
}

// H (row-major 3x3 = sum w' xc yc^T), weighted means xb / yb  ->  R, t  (see file header)
__host__ __device__ inline void pose_from_covariance(const double* H9, const double* xb, const double* yb,
                                                     double R[3][3], double tt[3]) {
    double H[3][3] = {{H9[0], H9[1], H9[2]}, {H9[3], H9[4], H9[5]}, {H9[6], H9[7], H9[8]}};
    double A[3][3], V[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A[i][j] = H[0][i] * H[0][j] + H[1][i] * H[1][j] + H[2][i] * H[2][j];  // H^T H
    jacobi_eig3(A, V);
    // two largest eigenvalues -> v0, v1
    int o0 = 0, o1 = 1, o2 = 2;
    double e[3] = {A[0][0], A[1][1], A[2][2]};
    if (e[o0] < e[o1]) { int t = o0; o0 = o1; o1 = t; }
    if (e[o1] < e[o2]) { int t = o1; o1 = o2; o2 = t; }
    if (e[o0] < e[o1]) { int t = o0; o0 = o1; o1 = t; }
    double v0[3] = {V[0][o0], V[1][o0], V[2][o0]}, v1[3] = {V[0][o1], V[1][o1], V[2][o1]};
    double u0[3], u1[3];
    for (int i = 0; i < 3; ++i) {
        u0[i] = H[i][0] * v0[0] + H[i][1] * v0[1] + H[i][2] * v0[2];
        u1[i] = H[i][0] * v1[0] + H[i][1] * v1[1] + H[i][2] * v1[2];
    }
    const double n0 = sqrt(u0[0] * u0[0] + u0[1] * u0[1] + u0[2] * u0[2]);
    bool ok = isfinite(n0) && n0 > 0.0;
    for (int i = 0; i < 3; ++i) { tt[i] = 0.0; for (int j = 0; j < 3; ++j) R[i][j] = (i == j) ? 1.0 : 0.0; }
    if (ok) {
        for (int i = 0; i < 3; ++i) u0[i] /= n0;
        const double pr = u1[0] * u0[0] + u1[1] * u0[1] + u1[2] * u0[2];
        for (int i = 0; i < 3; ++i) u1[i] -= pr * u0[i];
        const double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
        ok = isfinite(n1) && n1 > 1e-14 * n0;
        if (ok) {
            for (int i = 0; i < 3; ++i) u1[i] /= n1;
            const double u2[3] = {u0[1] * u1[2] - u0[2] * u1[1], u0[2] * u1[0] - u0[0] * u1[2], u0[0] * u1[1] - u0[1] * u1[0]};
            const double v2[3] = {v0[1] * v1[2] - v0[2] * v1[1], v0[2] * v1[0] - v0[0] * v1[2], v0[0] * v1[1] - v0[1] * v1[0]};
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) R[i][j] = v0[i] * u0[j] + v1[i] * u1[j] + v2[i] * u2[j];
            for (int i = 0; i < 3; ++i) tt[i] = yb[i] - (R[i][0] * xb[0] + R[i][1] * xb[1] + R[i][2] * xb[2]);
        }
    }
}

__global__ void __launch_bounds__(256)
kabsch_kernel(const float* __restrict__ src, const float* __restrict__ cor, const float* __restrict__ w, int N,
              const float* __restrict__ R_prev, const float* __restrict__ t_prev, float* __restrict__ R_out,
              float* __restrict__ t_out, float* __restrict__ R_cmp, float* __restrict__ t_cmp, float* __restrict__ pose12) {
    __shared__ double s_red[8 * 9];
    const int b = blockIdx.x;
    src += (size_t)b * N * 3; cor += (size_t)b * N * 3; w += (size_t)b * N;
    const float eps = 1e-4f;

    double a1[1] = {0.0}, sw[1];
    for (int n = threadIdx.x; n < N; n += blockDim.x) a1[0] += (double)w[n];
    block_sum<1>(a1, s_red, sw);
    const float denom = (float)sw[0] + eps;                               // layers.py:471

    double a7[7] = {0, 0, 0, 0, 0, 0, 0}, m7[7];
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const double wn = (double)(w[n] / denom);                          // layers.py:472 (fp32 division)
        a7[0] += wn;
        a7[1] += wn * src[n * 3 + 0]; a7[2] += wn * src[n * 3 + 1]; a7[3] += wn * src[n * 3 + 2];
        a7[4] += wn * cor[n * 3 + 0]; a7[5] += wn * cor[n * 3 + 1]; a7[6] += wn * cor[n * 3 + 2];
    }
    block_sum<7>(a7, s_red, m7);
    const double d2 = (double)((float)m7[0] + eps);                       // layers.py:475-476
    const double xb[3] = {m7[1] / d2, m7[2] / d2, m7[3] / d2};
    const double yb[3] = {m7[4] / d2, m7[5] / d2, m7[6] / d2};
    const float xbf[3] = {(float)xb[0], (float)xb[1], (float)xb[2]};
    const float ybf[3] = {(float)yb[0], (float)yb[1], (float)yb[2]};

    double a9[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, H9[9];
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const double wn = (double)(w[n] / denom);
        const double x0 = (double)(src[n * 3 + 0] - xbf[0]), x1 = (double)(src[n * 3 + 1] - xbf[1]),
                     x2 = (double)(src[n * 3 + 2] - xbf[2]);                // layers.py:478 (fp32 subtraction)
        const double y0 = wn * (double)(cor[n * 3 + 0] - ybf[0]), y1 = wn * (double)(cor[n * 3 + 1] - ybf[1]),
                     y2 = wn * (double)(cor[n * 3 + 2] - ybf[2]);
        a9[0] += x0 * y0; a9[1] += x0 * y1; a9[2] += x0 * y2;
        a9[3] += x1 * y0; a9[4] += x1 * y1; a9[5] += x1 * y2;
        a9[6] += x2 * y0; a9[7] += x2 * y1; a9[8] += x2 * y2;
    }
    block_sum<9>(a9, s_red, H9);
    if (threadIdx.x != 0) return;

    double R[3][3], tt[3];
    // the reference forms t from its fp32 means (layers.py:501); use the same rounded means
    const double xbd[3] = {(double)xbf[0], (double)xbf[1], (double)xbf[2]};
    const double ybd[3] = {(double)ybf[0], (double)ybf[1], (double)ybf[2]};
    pose_from_covariance(H9, xbd, ybd, R, tt);
    float Rf[9], tf[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) Rf[i * 3 + j] = (float)R[i][j];
        tf[i] = (float)tt[i];
    }
    for (int i = 0; i < 9; ++i) R_out[b * 9 + i] = Rf[i];
    for (int i = 0; i < 3; ++i) t_out[b * 3 + i] = tf[i];
    if (R_cmp) {   // T = [R|t] * [R_prev|t_prev]   (models.py:108,125)
        const float* Rp = R_prev + b * 9;
        const float* tp = t_prev + b * 3;
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j)
                R_cmp[b * 9 + i * 3 + j] = fmaf(Rf[i * 3 + 2], Rp[6 + j], fmaf(Rf[i * 3 + 1], Rp[3 + j], Rf[i * 3] * Rp[j]));
            t_cmp[b * 3 + i] = fmaf(Rf[i * 3 + 2], tp[2], fmaf(Rf[i * 3 + 1], tp[1], Rf[i * 3] * tp[0])) + tf[i];
        }
    }
    if (pose12) {  // the final pose of this call (composed when a previous pose is given) as one [R | t] row: the message
                   // of the multi-GPU pose gather, written here so that no packing kernel runs in front of the collective
        const float* Rs = R_cmp ? R_cmp + b * 9 : R_out + b * 9;
        const float* ts = R_cmp ? t_cmp + b * 3 : t_out + b * 3;
        for (int i = 0; i < 9; ++i) pose12[b * 12 + i] = Rs[i];
        for (int i = 0; i < 3; ++i) pose12[b * 12 + 9 + i] = ts[i];
    }
}

}  // namespace

// src, cor [B,N,3]; w [B,N] -> R [B,9] row-major, t [B,3].  If R_prev/t_prev are given, also the composed pose
// R_cmp = R R_prev, t_cmp = R t_prev + t.  pose12 (nullable) [B,12]: the final pose of the call as packed rows [R | t].
HRN_API int hrn_weighted_kabsch(const float* src, const float* cor, const float* w, int B, int N, const float* R_prev,
                                const float* t_prev, float* R, float* t, float* R_cmp, float* t_cmp, float* pose12,
                                void* stream) {
    if (!src || !cor || !w || !R || !t || B < 0 || N <= 0) return HRN_ERR_BAD_ARG;
    if ((R_cmp || t_cmp) && !(R_prev && t_prev && R_cmp && t_cmp)) return HRN_ERR_BAD_ARG;
    if (B == 0) return HRN_OK;
    kabsch_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(src, cor, w, N, R_prev, t_prev, R, t, R_cmp, t_cmp, pose12);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}

// Host-side evaluation of the same closed form (unit-testable without a GPU): H row-major 3x3.
HRN_API int hrn_pose_from_covariance_host(const double* H9, const double* xbar, const double* ybar, double* R9,
                                          double* t3) {
    if (!H9 || !xbar || !ybar || !R9 || !t3) return HRN_ERR_BAD_ARG;
    double R[3][3], tt[3];
    pose_from_covariance(H9, xbar, ybar, R, tt);
    for (int i = 0; i < 3; ++i) { t3[i] = tt[i]; for (int j = 0; j < 3; ++j) R9[i * 3 + j] = R[i][j]; }
    return HRN_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// RegressionHead (reference models/model_v2/layers.py:625-668, used by model_v3): the pose is REGRESSED from the two
// weighted means instead of solved -- w' = w / (sum w + 1e-4); x = [sum w' src | sum w' cor] (6 values);
// rotation = fc3_rot(relu(fc2_rot(relu(fc1_rot(x))))), translation likewise through the *_trans layers.  One CTA per
// pair: the means (fp64 sums of the fp32 products), then both three-layer perceptrons from shared memory.
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct RegMlp {
    const float *W1, *b1, *W2, *b2, *W3, *b3;      // nn.Linear layout [out, in]
};

__global__ void __launch_bounds__(128)
regression_head_kernel(const float* __restrict__ src, const float* __restrict__ cor, const float* __restrict__ w, int N,
                       RegMlp rot, RegMlp trn, int H1, int H2, int n_rot, float* __restrict__ rot_out,
                       float* __restrict__ trans_out) {
    __shared__ double s_red[4][7];
    __shared__ float s_x[6], s_h1[2][256], s_h2[2][256];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* sp = src + (size_t)b * N * 3;
    const float* cp = cor + (size_t)b * N * 3;
    const float* wp = w + (size_t)b * N;
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = tid; i < N; i += blockDim.x) acc[6] += (double)wp[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[6] += __shfl_xor_sync(0xffffffffu, acc[6], o);
    if (lane == 0) s_red[warp][6] = acc[6];
    __syncthreads();
    const float wsum = (float)(s_red[0][6] + s_red[1][6] + s_red[2][6] + s_red[3][6]) + 1e-4f;   // layers.py:648
    for (int i = tid; i < N; i += blockDim.x) {
        const float wn = wp[i] / wsum;                                                            // layers.py:649
        acc[0] += (double)(wn * sp[3 * i]); acc[1] += (double)(wn * sp[3 * i + 1]); acc[2] += (double)(wn * sp[3 * i + 2]);
        acc[3] += (double)(wn * cp[3 * i]); acc[4] += (double)(wn * cp[3 * i + 1]); acc[5] += (double)(wn * cp[3 * i + 2]);
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        if (lane == 0) s_red[warp][k] = acc[k];
    }
    __syncthreads();
    if (tid < 6) s_x[tid] = (float)(s_red[0][tid] + s_red[1][tid] + s_red[2][tid] + s_red[3][tid]);
    __syncthreads();
    for (int j = tid; j < 2 * H1; j += blockDim.x) {                 // fc1 of both branches
        const RegMlp& m = j < H1 ? rot : trn;
        const int o = j < H1 ? j : j - H1;
        float v = m.b1[o];
        for (int k = 0; k < 6; ++k) v = fmaf(m.W1[o * 6 + k], s_x[k], v);
        s_h1[j < H1 ? 0 : 1][o] = fmaxf(v, 0.f);
    }
    __syncthreads();
    for (int j = tid; j < 2 * H2; j += blockDim.x) {                 // fc2
        const int br = j < H2 ? 0 : 1;
        const RegMlp& m = br ? trn : rot;
        const int o = br ? j - H2 : j;
        float v = m.b2[o];
        for (int k = 0; k < H1; ++k) v = fmaf(m.W2[o * H1 + k], s_h1[br][k], v);
        s_h2[br][o] = fmaxf(v, 0.f);
    }
    __syncthreads();
    if (tid < n_rot + 3) {                                           // fc3
        const int br = tid < n_rot ? 0 : 1;
        const RegMlp& m = br ? trn : rot;
        const int o = br ? tid - n_rot : tid;
        float v = m.b3[o];
        for (int k = 0; k < H2; ++k) v = fmaf(m.W3[o * H2 + k], s_h2[br][k], v);
        if (br) trans_out[b * 3 + o] = v; else rot_out[b * n_rot + o] = v;
    }
}

}  // namespace

// src, cor [B,N,3]; w [B,N]; params = 12 device pointers {W1,b1,W2,b2,W3,b3} of the rotation branch then of the translation
// branch (nn.Linear layout, fp32): 6 -> H1 -> H2 -> n_rot / 3 (H1, H2 <= 256).  rot_out [B,n_rot], trans_out [B,3].
HRN_API int hrn_regression_head(const float* src, const float* cor, const float* w, int B, int N, const float* const* params,
                                int H1, int H2, int n_rot, float* rot_out, float* trans_out, void* stream) {
    if (!src || !cor || !w || !params || !rot_out || !trans_out || B < 0 || N <= 0) return HRN_ERR_BAD_ARG;
    for (int i = 0; i < 12; ++i) if (!params[i]) return HRN_ERR_BAD_ARG;
    if (H1 <= 0 || H2 <= 0 || H1 > 256 || H2 > 256 || n_rot <= 0 || n_rot > 9) return HRN_ERR_UNSUPPORTED;
    if (B == 0) return HRN_OK;
    RegMlp r{params[0], params[1], params[2], params[3], params[4], params[5]};
    RegMlp t{params[6], params[7], params[8], params[9], params[10], params[11]};
    regression_head_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(src, cor, w, N, r, t, H1, H2, n_rot, rot_out, trans_out);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}
