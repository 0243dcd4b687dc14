// Fused hierarchy level (level 1 of HierFeatureExtraction): grouping + KeypointDetector + DescExtractor in ONE
// persistent kernel -- no per-neighbour tensor ever reaches HBM.
//
// Replaces, for the first level (reference models/HRegNet/models.py:27-28, in_channels = 0, k = 64):
//   knn_group (layers.py:9-27)  ->  KeypointDetector.convs / attention / keypoints / attentive feature
//   (layers.py:150-159)  ->  DescExtractor.convs / max / cat / mlp1 / mlp2 / max (layers.py:200-209).
// The unfused form of this level moves ~11 GB per 32-pair step through HBM (4.2 M rows x 64 channels x a dozen
// passes); here the inputs are 16 B per row (neighbour index + coordinates) and the outputs are per KEYPOINT only
// (xyz 12 B, attentive feature 256 B, descriptor 256 B).
//
// CTA = 128 rows = 2 keypoints x 64 neighbours; thread t <-> row t <-> TMEM lane t; persistent over tiles, two
// CTAs per SM (<= 113 KB smem, 256 TMEM columns each) so one CTA's epilogue math overlaps the other's MMAs.
// All ten weight matrices of the level (60 KB as bf16 hi/lo UMMA tiles) stay resident in shared memory.
// Chain per tile (every layer: A operand in smem -> tcgen05.mma bf16x3 -> TMEM -> tcgen05.ld -> bias+ReLU ->
// bf16 hi/lo split -> written back IN PLACE as the next layer's A operand):
//   G(4->16) -d1-> 32 -d2-> 32 -d3-> E(64, stays in TMEM)   | max_c, softmax over the 64 neighbours, keypoint
//   G        -x1-> 32 -x2-> 32 -x3-> X1(64)                 | column max over the group
//   mlp1 = Wb.X1 + Wa.max_k(X1) + Wc.(E*a)  (three K-segments accumulated in TMEM) -> 32 -mlp2-> 64 -> max_k
#include "common.cuh"
#include "tc_common.cuh"
#include <math_constants.h>

namespace {

constexpr int TMR = 128;

template <int KNBR, int C1, int C2, int CO, int CMID, int CD>
struct LevelCfg {
    static constexpr int KG = 16;                                   // padded grouped-input channels (4 used)
    // shared-memory weight tiles: per layer hi plane [K/8][N][16B] then lo plane
    static constexpr int W_D1 = 0;
    static constexpr int W_D2 = W_D1 + 4 * KG * C1;
    static constexpr int W_D3 = W_D2 + 4 * C1 * C2;
    static constexpr int W_X1 = W_D3 + 4 * C2 * CO;
    static constexpr int W_X2 = W_X1 + 4 * KG * C1;
    static constexpr int W_X3 = W_X2 + 4 * C1 * C2;
    static constexpr int W_MA = W_X3 + 4 * C2 * CO;                 // mlp1, columns of max_k(X1)
    static constexpr int W_MB = W_MA + 4 * CO * CMID;               // mlp1, columns of X1
    static constexpr int W_MC = W_MB + 4 * CO * CMID;               // mlp1, columns of the attentive feature map
    static constexpr int W_M2 = W_MC + 4 * CO * CMID;
    static constexpr int W_BYTES = W_M2 + 4 * CMID * CD;
    // biases (floats): d1,d2,d3,x1,x2,x3,m1,m2
    static constexpr int B_D1 = 0, B_D2 = B_D1 + C1, B_D3 = B_D2 + C2, B_X1 = B_D3 + CO, B_X2 = B_X1 + C1,
                         B_X3 = B_X2 + C2, B_M1 = B_X3 + CO, B_M2 = B_M1 + CMID, B_COUNT = B_M2 + CD;
    static constexpr int OPC = (CO > C1 ? CO : C1) / 8;             // chunks of the in-place operand buffer
    static constexpr int G_BYTES = 2 * (KG / 8) * TMR * 16;
    static constexpr int OP_BYTES = 2 * OPC * TMR * 16;
    static constexpr int SMEM = W_BYTES + B_COUNT * 4 + G_BYTES + OP_BYTES + 4 * CO * 4 * 2 + 256;
    // TMEM columns
    static constexpr int T_ACC0 = 0, T_ACCM = 32, T_ACCE = 64, T_ACCX = 128, T_COLS = 256;
};

__device__ __forceinline__ uint32_t make_idesc(int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TMR >> 4) << 24);
}

// one thread: D[tmem_col] (+)= A(smem, K, hi/lo planes) x W(smem, [N x K], hi/lo planes)^T, bf16x3
__device__ __forceinline__ void issue_layer(uint32_t a_hi, uint32_t a_lo, int K, uint32_t w_hi, int N, uint32_t tmem_d,
                                            bool accumulate) {
    const uint32_t idesc = make_idesc(N);
    const uint32_t a_lbo = TMR * 16, w_lbo = (uint32_t)N * 16;
    const uint32_t w_lo = w_hi + (uint32_t)(K / 8) * N * 16;
    for (int k = 0; k < K / 16; ++k) {
        const uint64_t ah = umma_desc(a_hi + k * 2 * a_lbo, a_lbo, 128);
        const uint64_t al = umma_desc(a_lo + k * 2 * a_lbo, a_lbo, 128);
        const uint64_t wh = umma_desc(w_hi + k * 2 * w_lbo, w_lbo, 128);
        const uint64_t wl = umma_desc(w_lo + k * 2 * w_lbo, w_lbo, 128);
        umma_bf16(tmem_d, ah, wh, idesc, (accumulate || k > 0) ? 1u : 0u);
        umma_bf16(tmem_d, al, wh, idesc, 1u);
        umma_bf16(tmem_d, ah, wl, idesc, 1u);
    }
}

// lane c receives op-reduction over the 32 lanes of v[c]  (31 shuffles)
template <bool IS_MAX>
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = up ? v[i] : v[i + off];
            const float keep = up ? v[i + off] : v[i];
            const float recv = __shfl_xor_sync(0xffffffffu, send, off);
            v[i] = IS_MAX ? fmaxf(keep, recv) : keep + recv;
        }
    }
    return v[0];
}

template <class Cfg, int KNBR, int C1, int C2, int CO, int CMID, int CD>
__global__ void __launch_bounds__(TMR, 2)
level1_fused_kernel(const float* __restrict__ q, const float* __restrict__ xyz, const int32_t* __restrict__ idx,
                    const uint8_t* __restrict__ Wpack, const float* __restrict__ biases, float* __restrict__ out_xyz,
                    float* __restrict__ out_af, float* __restrict__ out_desc, int M, int N, int n_tiles) {
    static_assert(KNBR == 64, "row-group reductions are written for 64 neighbours (2 warps per keypoint)");
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    __shared__ float s_red[4][8];

    uint8_t* sW = smem;
    float* sB = reinterpret_cast<float*>(smem + Cfg::W_BYTES);
    uint8_t* sG = smem + Cfg::W_BYTES + Cfg::B_COUNT * 4;
    uint8_t* sOp = sG + Cfg::G_BYTES;
    float* sCol = reinterpret_cast<float*>(sOp + Cfg::OP_BYTES);          // [2][4][CO] cross-warp column partials

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = warp >> 1;                                            // keypoint of this row inside the tile
    const uint32_t bar = smem_u32(&s_bar);
    uint32_t phase = 0;

    // ---- one-time: weights + biases resident, barrier, TMEM ---------------------------------------------------
    for (int i = tid; i < Cfg::W_BYTES / 16; i += TMR)
        reinterpret_cast<uint4*>(sW)[i] = __ldg(reinterpret_cast<const uint4*>(Wpack) + i);
    for (int i = tid; i < Cfg::B_COUNT; i += TMR) sB[i] = __ldg(biases + i);
    {   // K padding chunk of G (channels 8..15) is zero forever
        uint4* g_hi = reinterpret_cast<uint4*>(sG);
        uint4* g_lo = g_hi + (Cfg::KG / 8) * TMR;
        g_hi[1 * TMR + tid] = make_uint4(0, 0, 0, 0);
        g_lo[1 * TMR + tid] = make_uint4(0, 0, 0, 0);
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(Cfg::T_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t lane_base = ((uint32_t)(warp * 32) << 16);
    const uint32_t aG_hi = smem_u32(sG), aG_lo = aG_hi + (Cfg::KG / 8) * TMR * 16;
    const uint32_t aOp_hi = smem_u32(sOp), aOp_lo = aOp_hi + Cfg::OPC * TMR * 16;
    const uint32_t wbase = smem_u32(sW);
    uint4* op_hi = reinterpret_cast<uint4*>(sOp);
    uint4* op_lo = op_hi + Cfg::OPC * TMR;

    // operand ready in smem -> one thread issues the layer -> everybody waits for the accumulator
    auto run_layer = [&](uint32_t a_hi, uint32_t a_lo, int K, int w_off, int Nn, int tcol, bool acc) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            issue_layer(a_hi, a_lo, K, wbase + w_off, Nn, tmem + tcol, acc);
            umma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    };
    // accumulator [tcol, tcol+Nn) -> relu(x + b) -> bf16 hi/lo operand (in place)
    auto epi_to_operand = [&](int tcol, int Nn, const float* b) {
        for (int c0 = 0; c0 < Nn; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem + lane_base + tcol + c0, v);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                float x[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) x[e] = fmaxf(__uint_as_float(v[ch * 8 + e]) + b[c0 + ch * 8 + e], 0.f);
                split_store8(x, op_hi + (c0 / 8 + ch) * TMR + tid, op_lo + (c0 / 8 + ch) * TMR + tid);
            }
        }
    };

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // ---- S0: grouped input  [rel xyz, |rel|] ------------------------------------------------------------
        const long long r = (long long)tile * TMR + tid;
        const long long bm = r / KNBR;
        const long long b = bm / M;
        const int n = __ldg(idx + r);
        const float* pp = xyz + (b * N + n) * 3;
        const float* qq = q + bm * 3;
        const float nx = __ldg(pp), ny = __ldg(pp + 1), nz = __ldg(pp + 2);
        const float rx = nx - __ldg(qq), ry = ny - __ldg(qq + 1), rz = nz - __ldg(qq + 2);
        {
            const float x[8] = {rx, ry, rz, sqrtf(rx * rx + ry * ry + rz * rz), 0.f, 0.f, 0.f, 0.f};
            uint4* g_hi = reinterpret_cast<uint4*>(sG);
            split_store8(x, g_hi + tid, g_hi + (Cfg::KG / 8) * TMR + tid);
        }
        // ---- detector chain ------------------------------------------------------------------------------------
        run_layer(aG_hi, aG_lo, Cfg::KG, Cfg::W_D1, C1, Cfg::T_ACC0, false);
        epi_to_operand(Cfg::T_ACC0, C1, sB + Cfg::B_D1);
        run_layer(aOp_hi, aOp_lo, C1, Cfg::W_D2, C2, Cfg::T_ACC0, false);
        epi_to_operand(Cfg::T_ACC0, C2, sB + Cfg::B_D2);
        run_layer(aOp_hi, aOp_lo, C2, Cfg::W_D3, CO, Cfg::T_ACCE, false);
        // ---- attention: a = softmax_k(max_c E), keypoint = sum_k a * nn -------------------------------------
        float x1 = 0.f;                                                   // post-ReLU values are >= 0
        for (int c0 = 0; c0 < CO; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem + lane_base + Cfg::T_ACCE + c0, v);
#pragma unroll
            for (int e = 0; e < 32; ++e) x1 = fmaxf(x1, __uint_as_float(v[e]) + sB[Cfg::B_D3 + c0 + e]);
        }
        float wm = hrn_warp_max(x1);
        if (lane == 0) s_red[warp][0] = wm;
        __syncthreads();
        const float gmax = fmaxf(s_red[2 * grp][0], s_red[2 * grp + 1][0]);
        const float ex = expf(x1 - gmax);
        float s0 = hrn_warp_sum(ex), s1 = hrn_warp_sum(ex * nx), s2 = hrn_warp_sum(ex * ny), s3 = hrn_warp_sum(ex * nz);
        if (lane == 0) { s_red[warp][1] = s0; s_red[warp][2] = s1; s_red[warp][3] = s2; s_red[warp][4] = s3; }
        __syncthreads();
        const float ssum = s_red[2 * grp][1] + s_red[2 * grp + 1][1];
        const float a = ex / ssum;
        if ((tid & 63) < 3) {
            const int c = tid & 63;
            out_xyz[bm * 3 + c] = (s_red[2 * grp][2 + c] + s_red[2 * grp + 1][2 + c]) / ssum;
        }
        // ---- descriptor chain ----------------------------------------------------------------------------------
        run_layer(aG_hi, aG_lo, Cfg::KG, Cfg::W_X1, C1, Cfg::T_ACC0, false);
        epi_to_operand(Cfg::T_ACC0, C1, sB + Cfg::B_X1);
        run_layer(aOp_hi, aOp_lo, C1, Cfg::W_X2, C2, Cfg::T_ACC0, false);
        epi_to_operand(Cfg::T_ACC0, C2, sB + Cfg::B_X2);
        run_layer(aOp_hi, aOp_lo, C2, Cfg::W_X3, CO, Cfg::T_ACCX, false);
        // X1 -> operand, and its column maximum over the 64 rows of the group
        for (int c0 = 0; c0 < CO; c0 += 32) {
            uint32_t v[32];
            float f[32];
            tmem_ld32(tmem + lane_base + Cfg::T_ACCX + c0, v);
#pragma unroll
            for (int e = 0; e < 32; ++e) f[e] = fmaxf(__uint_as_float(v[e]) + sB[Cfg::B_X3 + c0 + e], 0.f);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                const float x[8] = {f[ch * 8], f[ch * 8 + 1], f[ch * 8 + 2], f[ch * 8 + 3], f[ch * 8 + 4], f[ch * 8 + 5],
                                    f[ch * 8 + 6], f[ch * 8 + 7]};
                split_store8(x, op_hi + (c0 / 8 + ch) * TMR + tid, op_lo + (c0 / 8 + ch) * TMR + tid);
            }
            const float cm = warp_transpose_reduce<true>(f, lane);
            sCol[warp * CO + c0 + lane] = cm;
        }
        run_layer(aOp_hi, aOp_lo, CO, Cfg::W_MB, CMID, Cfg::T_ACCM, false);      // (barrier inside also publishes sCol)
        // max_k(X1) broadcast over the group's rows as the next K-segment
#pragma unroll
        for (int ch = 0; ch < CO / 8; ++ch) {
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
                x[e] = fmaxf(sCol[(2 * grp) * CO + ch * 8 + e], sCol[(2 * grp + 1) * CO + ch * 8 + e]);
            split_store8(x, op_hi + ch * TMR + tid, op_lo + ch * TMR + tid);
        }
        run_layer(aOp_hi, aOp_lo, CO, Cfg::W_MA, CMID, Cfg::T_ACCM, true);
        // attentive feature map E*a as the third K-segment; attentive feature = its column sum over the group
        float* sCol2 = sCol + 4 * CO;
        for (int c0 = 0; c0 < CO; c0 += 32) {
            uint32_t v[32];
            float f[32];
            tmem_ld32(tmem + lane_base + Cfg::T_ACCE + c0, v);
#pragma unroll
            for (int e = 0; e < 32; ++e) f[e] = fmaxf(__uint_as_float(v[e]) + sB[Cfg::B_D3 + c0 + e], 0.f) * a;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                const float x[8] = {f[ch * 8], f[ch * 8 + 1], f[ch * 8 + 2], f[ch * 8 + 3], f[ch * 8 + 4], f[ch * 8 + 5],
                                    f[ch * 8 + 6], f[ch * 8 + 7]};
                split_store8(x, op_hi + (c0 / 8 + ch) * TMR + tid, op_lo + (c0 / 8 + ch) * TMR + tid);
            }
            const float cs = warp_transpose_reduce<false>(f, lane);
            sCol2[warp * CO + c0 + lane] = cs;
        }
        run_layer(aOp_hi, aOp_lo, CO, Cfg::W_MC, CMID, Cfg::T_ACCM, true);
        if ((tid & 63) < CO) {
            const int c = tid & 63;
            out_af[bm * CO + c] = sCol2[(2 * grp) * CO + c] + sCol2[(2 * grp + 1) * CO + c];
        }
        // ---- mlp1 epilogue -> mlp2 -> descriptor = max_k -------------------------------------------------------
        epi_to_operand(Cfg::T_ACCM, CMID, sB + Cfg::B_M1);
        run_layer(aOp_hi, aOp_lo, CMID, Cfg::W_M2, CD, Cfg::T_ACCX, false);
        for (int c0 = 0; c0 < CD; c0 += 32) {
            uint32_t v[32];
            float f[32];
            tmem_ld32(tmem + lane_base + Cfg::T_ACCX + c0, v);
#pragma unroll
            for (int e = 0; e < 32; ++e) f[e] = fmaxf(__uint_as_float(v[e]) + sB[Cfg::B_M2 + c0 + e], 0.f);
            const float cm = warp_transpose_reduce<true>(f, lane);
            sCol[warp * CO + c0 + lane] = cm;
        }
        __syncthreads();
        if ((tid & 63) < CD) {
            const int c = tid & 63;
            out_desc[bm * CD + c] = fmaxf(sCol[(2 * grp) * CO + c], sCol[(2 * grp + 1) * CO + c]);
        }
        // the next tile's first barrier (inside run_layer) orders these reads before sCol / s_red are rewritten
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(Cfg::T_COLS) : "memory");
}

using Cfg1 = LevelCfg<64, 32, 32, 64, 32, 64>;

}  // namespace

// Level-1 fused detector + descriptor.  q [B*M,3] sampled keypoint coordinates, xyz [B,N,3], idx [B*M*64] int32
// neighbour indices; Wpack / biases from engine_tc.pack_level1 (layout = LevelCfg offsets).  Outputs per keypoint:
// out_xyz [B*M,3], out_af [B*M,64], out_desc [B*M,64].
HRN_API int hrn_level1_fused(const float* q, const float* xyz, const int32_t* idx, const void* Wpack, const float* biases,
                             float* out_xyz, float* out_af, float* out_desc, int B, int M, int N, int k, void* stream) {
    if (!q || !xyz || !idx || !Wpack || !biases || !out_xyz || !out_af || !out_desc || B < 0 || M <= 0 || N <= 0) return HRN_ERR_BAD_ARG;
    if (k != 64 || ((long long)B * M * k) % TMR != 0) return HRN_ERR_UNSUPPORTED;
    if (B == 0) return HRN_OK;
    const int n_tiles = (int)((long long)B * M * k / TMR);
    auto kern = level1_fused_kernel<Cfg1, 64, 32, 32, 64, 32, 64>;
    static bool attr_set = false;
    if (!attr_set) {
        HRN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg1::SMEM));
        attr_set = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = n_tiles < 2 * sms ? n_tiles : 2 * sms;
    kern<<<grid, TMR, Cfg1::SMEM, (cudaStream_t)stream>>>(q, xyz, (const int32_t*)idx, (const uint8_t*)Wpack, biases, out_xyz,
                                                          out_af, out_desc, M, N, n_tiles);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}

HRN_API int hrn_level1_pack_bytes(void) { return Cfg1::W_BYTES; }
HRN_API int hrn_level1_bias_count(void) { return Cfg1::B_COUNT; }
