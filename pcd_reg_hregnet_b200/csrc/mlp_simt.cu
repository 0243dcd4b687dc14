// Shared-MLP layer, exact-fp32 variant (CUDA-core FFMA):  Y[r, n] = act( sum_k X[r, k] * W[n, k] + bias[n] ).
//
// Replaces the reference's  nn.Conv2d(1x1, bias=False) + nn.BatchNorm2d (eval) + nn.ReLU  triples and the
// Conv1d(+BN)(+ReLU) heads (models/HRegNet/layers.py:115-130,183-198,246-268,417-431): BatchNorm running
// statistics are folded into (W, bias) on the host (pcd_reg_hregnet_b200/fold.py), the activation runs in
// the epilogue, and X is the virtual "rows" concatenation of rows.cuh -- the repeat / knn_gather / cat /
// permute().contiguous() tensors of the reference (layers.py:21-27,281-288,364-384,437-446) are never written.
//
// This variant is the fp32 parity mode (precision == "fp32"); the tensor-core variant lives in mlp_tc.cu.
#include "common.cuh"
#include "rows.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == HRN_ACT_RELU) return fmaxf(v, 0.f);
    if (act == HRN_ACT_SOFTPLUS_EPS) return (v > 20.f ? v : log1pf(expf(v))) + 0.001f;  // nn.Softplus() + 0.001, layers.py:162
    if (act == HRN_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
    return v;
}

__global__ void __launch_bounds__(256)
layer_simt_kernel(const hrn_rows_t in, const float* __restrict__ W, const float* __restrict__ bias, int act,
                  float* __restrict__ Y, int ldy, long long rows, int Cout, int K) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Ws[BK][BN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long row0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int kk = tid & 15, rr = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += BK) {
        const int kg = k0 + kk;
        // locate the segment of column kg
        int s = 0, ks = 0;
        while (s < in.n_seg - 1 && kg >= ks + in.seg[s].channels) { ks += in.seg[s].channels; ++s; }
        const hrn_seg_t sg = in.seg[s];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rl = rr + 16 * i;
            const long long r = row0 + rl;
            float v = 0.f;
            if (r < rows && kg < K) {
                const long long sr = hrn_src_row(in, sg.mode, r);
                v = __ldg(sg.ptr + sr * sg.ld + sg.col0 + (kg - ks));
                if (sg.row_scale) v *= __ldg(sg.row_scale + r);
            }
            As[kk][rl] = v;
            const int n = n0 + rl;
            Ws[kk][rl] = (n < Cout && kg < K) ? __ldg(W + (size_t)n * K + kg) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < BK; ++q) {
            const float4 a = *reinterpret_cast<const float4*>(&As[q][ty * 4]);
            const float4 w = *reinterpret_cast<const float4*>(&Ws[q][tx * 4]);
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
        const long long r = row0 + ty * 4 + i;
        if (r >= rows) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < Cout) Y[r * ldy + n] = apply_act(acc[i][j] + (bias ? __ldg(bias + n) : 0.f), act);
        }
    }
}

}  // namespace

HRN_API int hrn_layer_fp32(const hrn_rows_t* in, const float* W, const float* bias, int act, float* Y, int ldy,
                           long long rows, int Cout, void* stream) {
    if (!in || !W || !Y || rows < 0 || Cout <= 0 || in->n_seg < 1 || in->n_seg > 4) return HRN_ERR_BAD_ARG;
    if (rows == 0) return HRN_OK;
    int K = 0;
    for (int s = 0; s < in->n_seg; ++s) {
        if (!in->seg[s].ptr || in->seg[s].channels <= 0) return HRN_ERR_BAD_ARG;
        if (in->seg[s].mode == HRN_SEG_GATHER && !in->gather_idx) return HRN_ERR_BAD_ARG;
        if (in->seg[s].mode == HRN_SEG_BROADCAST && in->group <= 0) return HRN_ERR_BAD_ARG;
        K += in->seg[s].channels;
    }
    dim3 grid(hrn_divup(rows, BM), hrn_divup(Cout, BN));
    layer_simt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*in, W, bias, act, Y, ldy, rows, Cout, K);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}
