// Shared-MLP layer on the 5th-generation tensor cores (tcgen05 / TMEM), fp32-class accuracy:
//
//     Y[r, n] = act( sum_k X[r, k] * W[n, k] + bias[n] ),     X = virtual "rows" concatenation (rows.cuh)
//
// Replaces the same reference triples as mlp_simt.cu (Conv 1x1 + BatchNorm(eval) + ReLU, layers.py:118-121,
// 186-198,249-268,420-431).  Precision: both operands are split into bf16 hi + bf16 lo (x = hi + lo carries 16
// mantissa bits) and the product is evaluated as  hi*hi + lo*hi + hi*lo  with fp32 accumulation in TMEM
// ("bf16x3"): relative error ~2^-16 per product, far inside the 1e-3 feature tolerance, at 3 bf16 MMAs per
// useful MAC (a single-pass TF32/BF16 product, 2^-11 / 2^-9, is not).
//
// CTA = 128 rows x all Cout (<= 512) columns; 128 threads, thread t <-> row t <-> TMEM lane t.
//   stage loop over K in chunks of 32:
//     * A operand: every thread reads 32 fp32 of ITS row straight from the source matrices (gather /
//       broadcast / per-row scale resolved here -- the grouped tensor never exists), splits to bf16 hi/lo and
//       writes the UMMA no-swizzle K-major core-matrix layout  [k/8][row][8 x bf16]  (16-byte, conflict-free).
//     * B operand: weights are pre-split and pre-tiled on the host (engine_tc.pack_weights) so that a stage is
//       ONE contiguous block -> a single cp.async.bulk (TMA bulk copy, UBLKCP) completing on an mbarrier.
//     * one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N<=256, K=16), 3 per k-step per
//       N-half, accumulating in TMEM; tcgen05.commit -> mbarrier releases the stage (2-stage ring, so the
//       staging of chunk i+1 overlaps the MMAs of chunk i; several CTAs per SM overlap epilogues).
//   epilogue: tcgen05.ld 32 columns at a time -> +bias -> activation -> 128-byte row segments to HBM.
#include "common.cuh"
#include "rows.cuh"
#include <cuda_bf16.h>

namespace {

constexpr int TM = 128;       // rows per CTA (UMMA M)
constexpr int KC = 32;        // K per pipeline stage (two K=16 MMA steps)
constexpr int A_STAGE_BYTES = 2 /*hi,lo*/ * (KC / 8) * TM * 16;   // 16 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// UMMA shared-memory descriptor, K-major, no swizzle: core matrix = 8 rows x 16 bytes (128 contiguous bytes);
// LBO = byte distance between the two K-halves (core matrices adjacent in K), SBO = between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
        "%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float act_fn(float v, int act) {
    if (act == HRN_ACT_RELU) return fmaxf(v, 0.f);
    if (act == HRN_ACT_SOFTPLUS_EPS) return (v > 20.f ? v : log1pf(expf(v))) + 0.001f;
    if (act == HRN_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
    return v;
}

// split 8 floats into bf16 hi / lo and store both 16-byte core-matrix rows
__device__ __forceinline__ void split_store8(const float (&x)[8], uint4* dst_hi, uint4* dst_lo) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
        const float2 hf = __bfloat1622float2(h);
        const __nv_bfloat162 l = __floats2bfloat162_rn(x[2 * i] - hf.x, x[2 * i + 1] - hf.y);
        hi[i] = *reinterpret_cast<const uint32_t*>(&h);
        lo[i] = *reinterpret_cast<const uint32_t*>(&l);
    }
    *dst_hi = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *dst_lo = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// Wp: packed weights [n_stage][2 (hi,lo)][KC/8][NP][8] bf16;  NP = padded Cout (multiple of 16, <= 512)
__global__ void __launch_bounds__(TM)
layer_tc_kernel(const hrn_rows_t in, const __nv_bfloat16* __restrict__ Wp, const float* __restrict__ bias, int act,
                float* __restrict__ Y, int ldy, long long rows, int Cout, int NP, int n_stage, int tmem_cols) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_bar[5];     // [0,1] W landed, [2,3] stage consumed by MMA, [4] accumulator ready
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t w_stage_bytes = (uint32_t)NP * 128u;                 // 2 * (KC/8) * NP * 16
    uint8_t* sA[2] = {smem, smem + A_STAGE_BYTES};
    uint8_t* sW[2] = {smem + 2 * A_STAGE_BYTES, smem + 2 * A_STAGE_BYTES + w_stage_bytes};
    const uint32_t bar_w[2] = {smem_u32(&s_bar[0]), smem_u32(&s_bar[1])};
    const uint32_t bar_m[2] = {smem_u32(&s_bar[2]), smem_u32(&s_bar[3])};
    const uint32_t bar_done = smem_u32(&s_bar[4]);

    if (tid == 0) {
        for (int i = 0; i < 5; ++i) mbar_init(smem_u32(&s_bar[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    // ---- per-thread row bookkeeping ------------------------------------------------------------------------
    const long long r = (long long)blockIdx.x * TM + tid;
    const bool rvalid = r < rows;
    const float* sp[4];
    float rscale[4];
    int cstart[5];                     // first 8-wide K chunk of each segment
    cstart[0] = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        sp[s] = nullptr; rscale[s] = 1.f;
        int nch = 0;
        if (s < in.n_seg) {
            const hrn_seg_t sg = in.seg[s];
            nch = (sg.channels + 7) >> 3;
            if (rvalid) {
                sp[s] = sg.ptr + hrn_src_row(in, sg.mode, r) * sg.ld + sg.col0;
                if (sg.row_scale) rscale[s] = __ldg(sg.row_scale + r);
            }
        }
        cstart[s + 1] = cstart[s] + nch;
    }

    // instruction descriptor: D=f32, A=B=bf16, K-major both, N, M=128
    const int NH = NP > 256 ? 256 : NP;                       // columns per MMA
    const int n_half = NP > 256 ? 2 : 1;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NH >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

    for (int i = 0; i < n_stage; ++i) {
        const int s = i & 1;
        if (i >= 2) mbar_wait(bar_m[s], ((i >> 1) - 1) & 1);          // MMAs of chunk i-2 have drained this stage
        if (tid == 0) {
            mbar_expect_tx(bar_w[s], w_stage_bytes);
            bulk_g2s(smem_u32(sW[s]), Wp + (size_t)i * (w_stage_bytes / 2), w_stage_bytes, bar_w[s]);
        }
        // ---- A operand: this thread's row, KC/8 chunks of 8 -------------------------------------------------
        uint4* a_hi = reinterpret_cast<uint4*>(sA[s]);
        uint4* a_lo = a_hi + (KC / 8) * TM;
#pragma unroll
        for (int c = 0; c < KC / 8; ++c) {
            const int cg = i * (KC / 8) + c;                          // global 8-chunk index
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = 0.f;
            int sgi = 0;
#pragma unroll
            for (int q = 1; q < 4; ++q) if (cg >= cstart[q]) sgi = q;
            if (rvalid && cg < cstart[4] && sgi < in.n_seg) {
                const int ch0 = (cg - cstart[sgi]) << 3;
                const int nvalid = min(8, in.seg[sgi].channels - ch0);
                const float* p = sp[0];
                float sc = rscale[0];
#pragma unroll
                for (int q = 1; q < 4; ++q) if (sgi == q) { p = sp[q]; sc = rscale[q]; }
                p += ch0;
                if (nvalid == 8 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
                    const float4 v0 = __ldg(reinterpret_cast<const float4*>(p));
                    const float4 v1 = __ldg(reinterpret_cast<const float4*>(p) + 1);
                    x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w; x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) if (e < nvalid) x[e] = __ldg(p + e);
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) x[e] *= sc;
            }
            split_store8(x, a_hi + c * TM + tid, a_lo + c * TM + tid);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            mbar_wait(bar_w[s], (i >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_base = smem_u32(sA[s]);
            const uint32_t w_base = smem_u32(sW[s]);
            const uint32_t a_lbo = TM * 16, w_lbo = (uint32_t)NP * 16;
            const uint32_t a_lo_off = (KC / 8) * TM * 16, w_lo_off = (KC / 8) * (uint32_t)NP * 16;
#pragma unroll
            for (int k = 0; k < KC / 16; ++k) {
                const uint64_t ah = umma_desc(a_base + k * 2 * a_lbo, a_lbo, 128);
                const uint64_t al = umma_desc(a_base + a_lo_off + k * 2 * a_lbo, a_lbo, 128);
                for (int h = 0; h < n_half; ++h) {
                    const uint64_t wh = umma_desc(w_base + k * 2 * w_lbo + h * 256 * 16, w_lbo, 128);
                    const uint64_t wl = umma_desc(w_base + w_lo_off + k * 2 * w_lbo + h * 256 * 16, w_lbo, 128);
                    const uint32_t d = tmem + h * 256;
                    umma_bf16(d, ah, wh, idesc, (i > 0 || k > 0) ? 1u : 0u);
                    umma_bf16(d, al, wh, idesc, 1u);
                    umma_bf16(d, ah, wl, idesc, 1u);
                }
            }
            umma_commit(bar_m[s]);
            if (i == n_stage - 1) umma_commit(bar_done);
        }
    }
    // ---- epilogue ---------------------------------------------------------------------------------------------
    mbar_wait(bar_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* yrow = Y + r * ldy;
    for (int c0 = 0; c0 < Cout; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        if (rvalid) {
            if (c0 + 32 <= Cout && ((reinterpret_cast<uintptr_t>(yrow + c0) & 15) == 0)) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 o;
                    o.x = act_fn(__uint_as_float(v[j + 0]) + __ldg(bias + c0 + j + 0), act);
                    o.y = act_fn(__uint_as_float(v[j + 1]) + __ldg(bias + c0 + j + 1), act);
                    o.z = act_fn(__uint_as_float(v[j + 2]) + __ldg(bias + c0 + j + 2), act);
                    o.w = act_fn(__uint_as_float(v[j + 3]) + __ldg(bias + c0 + j + 3), act);
                    *reinterpret_cast<float4*>(yrow + c0 + j) = o;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (c0 + j < Cout) yrow[c0 + j] = act_fn(__uint_as_float(v[j]) + __ldg(bias + c0 + j), act);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
}

}  // namespace

// Wp = weights packed by hrn_pack_weights_layout (see engine_tc.pack_weights): [n_stage][2][4][NP][8] bf16.
// K_pad = 32 * n_stage must equal sum over segments of ceil8(channels), rounded up to 32.
HRN_API int hrn_layer_tc(const hrn_rows_t* in, const void* Wp, const float* bias, int act, float* Y, int ldy,
                         long long rows, int Cout, int NP, int n_stage, void* stream) {
    if (!in || !Wp || !bias || !Y || rows < 0 || Cout <= 0 || in->n_seg < 1 || in->n_seg > 4) return HRN_ERR_BAD_ARG;
    if (NP % 16 != 0 || NP < 16 || NP > 512 || NP < Cout || (NP > 256 && NP != 512)) return HRN_ERR_UNSUPPORTED;
    int chunks = 0;
    for (int s = 0; s < in->n_seg; ++s) {
        if (!in->seg[s].ptr || in->seg[s].channels <= 0) return HRN_ERR_BAD_ARG;
        if (in->seg[s].mode == HRN_SEG_GATHER && !in->gather_idx) return HRN_ERR_BAD_ARG;
        if (in->seg[s].mode == HRN_SEG_BROADCAST && in->group <= 0) return HRN_ERR_BAD_ARG;
        chunks += (in->seg[s].channels + 7) / 8;
    }
    if (n_stage != (chunks * 8 + KC - 1) / KC) return HRN_ERR_BAD_ARG;
    if (rows == 0) return HRN_OK;
    int tmem_cols = 32;
    while (tmem_cols < NP) tmem_cols <<= 1;
    const size_t smem = 2 * (size_t)A_STAGE_BYTES + 2 * (size_t)NP * 128;
    static bool attr_set = false;
    if (!attr_set) {
        HRN_CUDA(cudaFuncSetAttribute(layer_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * A_STAGE_BYTES + 2 * 512 * 128));
        attr_set = true;
    }
    layer_tc_kernel<<<hrn_divup(rows, TM), TM, smem, (cudaStream_t)stream>>>(
        *in, (const __nv_bfloat16*)Wp, bias, act, Y, ldy, rows, Cout, NP, n_stage, tmem_cols);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}
