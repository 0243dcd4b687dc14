// Pose-error metrics on device, fused after the pose head / the pose all-gather.
//
// Replaces, per batch of registration pairs (reference file:line):
//   losses/losses.py:138-164   calc_rot_rre_err / calc_tran_rte_err: R_err = pred_R^T gt_R, Euler XYZ angles of R_err
//                              (mean |deg| per axis), geodesic distance acos((tr R_err - 1)/2) in degrees per pair,
//                              mean |pred_t - gt_t| per axis and its Euclidean norm per pair
//   models/utils.py:132-138    calc_error_np (the same two scalars for one pair, numpy)
//   metrics/calibeval.py:72-106,172-196  CalibEval.add_batch / geodesic_distance: error = pred_tf . gt_tf (4x4, gt =
//                              the applied perturbation), batch means of the geodesic angle and of ||t_err||
// -- a dozen ATen launches, two host round trips (.cpu()) and Python list appends per batch in the reference.
// Euler angles follow pytorch3d.transforms.matrix_to_euler_angles(M, "XYZ") = (atan2(-M12, M22), asin(M02),
// atan2(-M01, M00)) (pytorch3d is not vendored in the reference: restated, see oracle/ref_metrics.py).
// One thread per pair; batch sums by one block-wide reduction + atomicAdd (B is at most a few thousand).
#include "common.cuh"

namespace {

constexpr float RAD2DEG = 57.29577951308232f;

// mode 0: R_err = pred_R^T gt_R,  t_err = pred_t - gt_t            (losses.py, models/utils.py)
// mode 1: R_err = pred_R gt_R,    t_err = pred_R gt_t + pred_t      (calibeval.py: pred_tf . gt_tf)
__global__ void __launch_bounds__(128)
pose_error_kernel(const float* __restrict__ pR, const float* __restrict__ pt, const float* __restrict__ gR,
                  const float* __restrict__ gt, int B, int mode, float* __restrict__ geo, float* __restrict__ eucl,
                  float* __restrict__ euler, float* __restrict__ terr, float* __restrict__ sums) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // geo, eucl, |euler| xyz, |t_err| xyz
    if (b < B) {
        float P[9], G[9], E[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) { P[i] = pR[b * 9 + i]; G[i] = gR[b * 9 + i]; }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                float a = 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) a = fmaf(mode == 0 ? P[k * 3 + i] : P[i * 3 + k], G[k * 3 + j], a);
                E[i * 3 + j] = a;
            }
        const float tp[3] = {pt[b * 3], pt[b * 3 + 1], pt[b * 3 + 2]}, tg[3] = {gt[b * 3], gt[b * 3 + 1], gt[b * 3 + 2]};
        float te[3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
            te[i] = mode == 0 ? tp[i] - tg[i] : fmaf(P[i * 3], tg[0], fmaf(P[i * 3 + 1], tg[1], fmaf(P[i * 3 + 2], tg[2], tp[i])));
        const float c = fminf(fmaxf((E[0] + E[4] + E[8] - 1.f) * 0.5f, -1.f), 1.f);
        v[0] = acosf(c) * RAD2DEG;
        v[1] = sqrtf(te[0] * te[0] + te[1] * te[1] + te[2] * te[2]);
        const float ex = atan2f(-E[5], E[8]) * RAD2DEG, ey = asinf(fminf(fmaxf(E[2], -1.f), 1.f)) * RAD2DEG,
                    ez = atan2f(-E[1], E[0]) * RAD2DEG;
        if (geo) geo[b] = v[0];
        if (eucl) eucl[b] = v[1];
        if (euler) { euler[b * 3] = ex; euler[b * 3 + 1] = ey; euler[b * 3 + 2] = ez; }
        if (terr) { terr[b * 3] = te[0]; terr[b * 3 + 1] = te[1]; terr[b * 3 + 2] = te[2]; }
        v[2] = fabsf(ex); v[3] = fabsf(ey); v[4] = fabsf(ez);
        v[5] = fabsf(te[0]); v[6] = fabsf(te[1]); v[7] = fabsf(te[2]);
    }
    if (!sums) return;
    __shared__ float s_red[4][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float a = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) s_red[warp][i] = a;
    }
    __syncthreads();
    if (threadIdx.x < 8) atomicAdd(sums + threadIdx.x, s_red[0][threadIdx.x] + s_red[1][threadIdx.x] + s_red[2][threadIdx.x] + s_red[3][threadIdx.x]);
}

}  // namespace

// pred_R, gt_R [B,9] row-major; pred_t, gt_t [B,3].  Per-pair outputs (each nullable): geo [B] degrees, eucl [B],
// euler [B,3] degrees (XYZ convention), terr [B,3]; sums [8] (nullable, ACCUMULATED into -- zero it first):
// sum geo, sum eucl, sum |euler| xyz, sum |terr| xyz, so that means over several batches / ranks are one division.
HRN_API int hrn_pose_errors(const float* pred_R, const float* pred_t, const float* gt_R, const float* gt_t, int B, int mode,
                            float* geo, float* eucl, float* euler, float* terr, float* sums, void* stream) {
    if (!pred_R || !pred_t || !gt_R || !gt_t || B < 0 || (mode != 0 && mode != 1)) return HRN_ERR_BAD_ARG;
    if (B == 0) return HRN_OK;
    pose_error_kernel<<<hrn_divup(B, 128), 128, 0, (cudaStream_t)stream>>>(pred_R, pred_t, gt_R, gt_t, B, mode, geo, eucl,
                                                                         euler, terr, sums);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}
