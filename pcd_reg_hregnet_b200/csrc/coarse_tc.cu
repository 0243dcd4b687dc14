// Descriptor-similarity features of CoarseReg on the tensor cores: the cosine matrix of a pair IS a dense
// [N1 x C] . [C x N2] contraction (reference models/HRegNet/layers.py:29-41, 290-313, 339-362; the north star names it
// as tcgen05 work).  One CTA per pair computes
//     cosT[n1, n2] = <S[n1], D[n2]> / (|S[n1]| |D[n2]| + 1e-6)
// with tcgen05.mma (operands split into bf16 hi + lo, all FOUR partial products hi.hi + lo.hi + hi.lo + lo.lo, fp32
// accumulation in TMEM: the dot products carry fp32-class accuracy, ~1e-6 on the cosine), takes both families of maxima
// (over n2 per source keypoint, over n1 per target keypoint) and the k picks per source keypoint from the accumulators,
// and writes ONLY the two feature columns
//     src_dst[b,i,j] = cos[idx[i,j], i] / (max_n2 cos[:, i] + 1e-6)      dst_src[b,i,j] = cos[idx[i,j], i] / (max_n1 cos[idx[i,j], :] + 1e-6)
// The [N2 x N1] matrix never reaches HBM (coarse.cu writes and re-reads it: 5 launches per similarity; this is 1).
//
// CTA = 256 threads; thread t owns source keypoint n1 = t = TMEM lane (t % 128) of M-tile (t / 128), and loads / splits
// row t of S (A operand) and row t of D (B operand) itself, 32 channels per stage, two stages in flight.
#include "common.cuh"
#include "tc_common.cuh"
#include <math_constants.h>

namespace {

constexpr int CT_ROWS = 256;                         // rows of a staged operand (N1, N2 <= 256)
constexpr int CT_KC = 32;                            // channels per stage
constexpr int CT_PLANE = (CT_KC / 8) * CT_ROWS * 16; // one bf16 plane of one operand of a stage: 16 KB
constexpr int CT_STAGE = 4 * CT_PLANE;               // S hi | S lo | D hi | D lo
constexpr int CT_TP = 36;                            // pick tile pitch (floats)
constexpr int CT_SMEM = 2 * CT_STAGE + CT_ROWS * CT_TP * 4 + 2 * CT_ROWS * 4;

__device__ __forceinline__ float ct_warp_transpose_max(float (&v)[32], int lane) {   // lane c <- max over lanes of v[c]
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = up ? v[i] : v[i + off];
            const float keep = up ? v[i + off] : v[i];
            v[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, off));
        }
    }
    return v[0];
}

template <int K>
__global__ void __launch_bounds__(256, 1)
cosine_features_tc_kernel(const float* __restrict__ S, const float* __restrict__ D, const int32_t* __restrict__ idx,
                          int N1, int N2, int C, float* __restrict__ out, int ldo, int col_sd, int col_ds) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_bar[3];          // [0,1] stage consumed by its MMAs, [2] accumulators complete
    __shared__ uint32_t s_tmem;
    float* sTile = reinterpret_cast<float*>(smem + 2 * CT_STAGE);       // [256][CT_TP]
    float* sND = sTile + CT_ROWS * CT_TP;                               // |D[n2]|
    unsigned* sRmax = reinterpret_cast<unsigned*>(sND + CT_ROWS);       // max_n1 cos[n2, :], order-preserving uint

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x;
    S += (size_t)b * N1 * C;
    D += (size_t)b * N2 * C;
    const bool has_s = tid < N1, has_d = tid < N2;
    const int m_tiles = N1 / 128;

    if (tid == 0) {
        for (int i = 0; i < 3; ++i) mbar_init(smem_u32(&s_bar[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    sRmax[tid] = hrn_ford(-CUDART_INF_F);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    const int n_stage = C / CT_KC;
    const float4* srow = reinterpret_cast<const float4*>(S + (size_t)tid * C);
    const float4* drow = reinterpret_cast<const float4*>(D + (size_t)tid * C);
    float ss = 0.f, dd = 0.f;                          // squared norms of this thread's rows
    const uint32_t idesc = umma_idesc_m128<3>(N2);
    const uint64_t fix = umma_desc_fixed(CT_ROWS * 16, 128);           // LBO = 4096 B between K-adjacent core matrices

    for (int st = 0; st < n_stage; ++st) {
        const int slot = st & 1;
        uint8_t* base = smem + slot * CT_STAGE;
        // the MMAs that read this slot two stages ago must have completed
        if (st >= 2) mbar_wait(smem_u32(&s_bar[slot]), ((st >> 1) - 1) & 1);
        float4 sv[CT_KC / 4], dv[CT_KC / 4];
#pragma unroll
        for (int i = 0; i < CT_KC / 4; ++i) {
            sv[i] = has_s ? __ldg(srow + st * (CT_KC / 4) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            dv[i] = has_d ? __ldg(drow + st * (CT_KC / 4) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        uint4* s_hi = reinterpret_cast<uint4*>(base), *s_lo = s_hi + CT_PLANE / 16;
        uint4* d_hi = s_lo + CT_PLANE / 16, *d_lo = d_hi + CT_PLANE / 16;
#pragma unroll
        for (int c = 0; c < CT_KC / 8; ++c) {
            const float xs[8] = {sv[2 * c].x, sv[2 * c].y, sv[2 * c].z, sv[2 * c].w, sv[2 * c + 1].x, sv[2 * c + 1].y, sv[2 * c + 1].z, sv[2 * c + 1].w};
            const float xd[8] = {dv[2 * c].x, dv[2 * c].y, dv[2 * c].z, dv[2 * c].w, dv[2 * c + 1].x, dv[2 * c + 1].y, dv[2 * c + 1].z, dv[2 * c + 1].w};
#pragma unroll
            for (int e = 0; e < 8; ++e) { ss = fmaf(xs[e], xs[e], ss); dd = fmaf(xd[e], xd[e], dd); }
            split_store8(xs, s_hi + c * CT_ROWS + tid, s_lo + c * CT_ROWS + tid);
            split_store8(xd, d_hi + c * CT_ROWS + tid, d_lo + c * CT_ROWS + tid);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a0 = smem_u32(base) >> 4, lo = CT_PLANE >> 4, dB = (2 * CT_PLANE) >> 4;
            for (int m = 0; m < m_tiles; ++m) {
#pragma unroll
                for (int ks = 0; ks < CT_KC / 16; ++ks) {
                    const uint32_t ah = a0 + ((ks * 2 * CT_ROWS * 16) >> 4) + ((m * 128 * 16) >> 4);
                    const uint32_t bh = a0 + dB + ((ks * 2 * CT_ROWS * 16) >> 4);
                    const uint32_t d = tmem + m * 256;
                    umma_bf16(d, fix | ah, fix | bh, idesc, (st > 0 || ks > 0) ? 1u : 0u);
                    umma_bf16(d, fix | (ah + lo), fix | bh, idesc, 1u);
                    umma_bf16(d, fix | ah, fix | (bh + lo), idesc, 1u);
                    umma_bf16(d, fix | (ah + lo), fix | (bh + lo), idesc, 1u);
                }
            }
            umma_commit(smem_u32(&s_bar[slot]));
            if (st == n_stage - 1) umma_commit(smem_u32(&s_bar[2]));
        }
    }
    const float nS = sqrtf(ss);
    sND[tid] = sqrtf(dd);
    int nb[K];
    if (has_s) {
#pragma unroll
        for (int j = 0; j < K; ++j) nb[j] = __ldg(idx + ((size_t)b * N1 + tid) * K + j);
    }
    float pick[K];
#pragma unroll
    for (int j = 0; j < K; ++j) pick[j] = 0.f;
    __syncthreads();                                                    // sND complete
    mbar_wait(smem_u32(&s_bar[2]), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    float cmax = -CUDART_INF_F;                                         // max over n2 of cos[:, n1]
    if (warp < 4 * m_tiles) {
        const uint32_t acc = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
        float* trow = sTile + tid * CT_TP;
        for (int c0 = 0; c0 < N2; c0 += 32) {
            uint32_t v[32];
            float f[32];
            tmem_ld32(acc + c0, v);
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                const bool ok = c0 + e < N2;
                f[e] = ok ? __uint_as_float(v[e]) / (nS * sND[ok ? c0 + e : 0] + 1e-6f) : -CUDART_INF_F;
                cmax = fmaxf(cmax, f[e]);
            }
#pragma unroll
            for (int e = 0; e < 32; e += 4) *reinterpret_cast<float4*>(trow + e) = make_float4(f[e], f[e + 1], f[e + 2], f[e + 3]);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < K; ++j)
                if (nb[j] >= c0 && nb[j] < c0 + 32) pick[j] = trow[nb[j] - c0];
            const float colm = ct_warp_transpose_max(f, lane);          // lane e: max over this warp's 32 source keypoints
            if (c0 + lane < N2) atomicMax(&sRmax[c0 + lane], hrn_ford(colm));
            __syncwarp();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                                                    // sRmax complete
    if (has_s) {
        float* o = out + ((size_t)b * N1 + tid) * K * ldo;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            o[j * ldo + col_sd] = pick[j] / (cmax + 1e-6f);
            o[j * ldo + col_ds] = pick[j] / (hrn_ford_inv(sRmax[nb[j]]) + 1e-6f);
        }
    }
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

}  // namespace

// S [B,N1,C] source descriptors, D [B,N2,C] target descriptors (channels-last), idx [B,N1,k] int32 candidates in the
// target cloud -> two columns of out [B*N1*k, ldo].  Shapes of the model path: N1 = 128 or 256, N2 <= 256 and a multiple
// of 16, C a multiple of 32, k = 8; HRN_ERR_UNSUPPORTED otherwise (callers then run hrn_cosine_matrix + hrn_cosine_pick).
HRN_API int hrn_cosine_features_tc(const float* S, const float* D, const int32_t* idx, int B, int N1, int N2, int C, int k,
                                   float* out, int ldo, int col_src_dst, int col_dst_src, void* stream) {
    if (!S || !D || !idx || !out || B < 0 || N1 <= 0 || N2 <= 0 || C <= 0 || k <= 0) return HRN_ERR_BAD_ARG;
    if ((N1 != 128 && N1 != 256) || N2 > 256 || (N2 % 16) || (C % CT_KC) || k != 8) return HRN_ERR_UNSUPPORTED;
    if (((uintptr_t)S | (uintptr_t)D) & 15) return HRN_ERR_UNSUPPORTED;
    if (B == 0) return HRN_OK;
    static hrn_once_per_device attr;
    if (attr.need())
        HRN_CUDA(cudaFuncSetAttribute(cosine_features_tc_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM));
    cosine_features_tc_kernel<8><<<B, 256, CT_SMEM, (cudaStream_t)stream>>>(S, D, idx, N1, N2, C, out, ldo, col_src_dst, col_dst_src);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}
