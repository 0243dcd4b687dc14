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
#include "tc_common.cuh"

namespace {

constexpr int TM = 128;       // rows per CTA (UMMA M)
constexpr int KC = 32;        // K per pipeline stage (two K=16 MMA steps)
constexpr int A_STAGE_BYTES = 2 /*hi,lo*/ * (KC / 8) * TM * 16;   // 16 KB

// Per-thread description of the virtual row: up to 4 segments, resolved once per CTA.
struct RowSrc {
    const float* p[4];     // row base pointer per segment (nullptr: row out of range)
    float sc[4];           // per-row scale per segment
    int c0[5];             // first 8-wide K chunk of each segment; c0[4] = total chunks
    int ch[4];             // channels per segment
};

// Fetch the 8 floats of global chunk `cg` of this thread's row as two predicated 16-byte loads (FAST: every
// segment has channels % 4 == 0 and 16-byte aligned rows, checked on the host) -- no divergent control flow, so
// the compiler batches the loads of a whole stage.
template <bool FAST>
__device__ __forceinline__ void load_chunk(const RowSrc& rs, int cg, float4& v0, float4& v1, float& sc) {
    int sgi = 0;
#pragma unroll
    for (int q = 1; q < 4; ++q) if (cg >= rs.c0[q]) sgi = q;
    const float* p = rs.p[0]; int cs = rs.c0[0], chn = rs.ch[0]; sc = rs.sc[0];
#pragma unroll
    for (int q = 1; q < 4; ++q) if (sgi == q) { p = rs.p[q]; cs = rs.c0[q]; chn = rs.ch[q]; sc = rs.sc[q]; }
    const int ch0 = (cg - cs) << 3;
    const int nvalid = (p != nullptr && cg < rs.c0[4]) ? chn - ch0 : 0;
    v0 = make_float4(0.f, 0.f, 0.f, 0.f); v1 = v0;
    if (FAST) {
        if (nvalid >= 4) v0 = __ldg(reinterpret_cast<const float4*>(p + ch0));
        if (nvalid >= 8) v1 = __ldg(reinterpret_cast<const float4*>(p + ch0) + 1);
    } else {
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = (e < nvalid) ? __ldg(p + ch0 + e) : 0.f;
        v0 = make_float4(x[0], x[1], x[2], x[3]); v1 = make_float4(x[4], x[5], x[6], x[7]);
    }
}

// Wp: packed weights [n_stage][2 (hi,lo)][KC/8][NP][8] bf16;  NP = padded Cout (multiple of 16, <= 512)
template <bool FAST>
__global__ void __launch_bounds__(TM)
layer_tc_kernel(const hrn_rows_t in, const __nv_bfloat16* __restrict__ Wp, const float* __restrict__ bias, int act,
                float* __restrict__ Y, int ldy, long long rows, int Cout, int NPfull, int NP, int n_stage, int tmem_cols, int NW) {
    // NPfull = padded Cout of the packed weights; NP = columns handled by this CTA (slice blockIdx.y of NPfull)
    extern __shared__ __align__(128) uint8_t smem[];
    // [0,1] A stage drained by its MMAs, [2] accumulator ready, [3..3+NW) W slot landed, [3+NW..3+2NW) W slot drained
    __shared__ __align__(8) uint64_t s_bar[3 + 2 * 4];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int n0 = blockIdx.y * NP;
    const uint32_t w_stage_bytes = (uint32_t)NP * 128u;                 // 2 * (KC/8) * NP * 16
    const size_t w_full_stage = (size_t)NPfull * 64;                    // bf16 elements per packed stage
    // one stage of this CTA's weight slice: 1 bulk copy (full width) or 8 (one per hi/lo plane and k-chunk)
    auto fetch_w = [&](int stage, uint32_t dst, uint32_t barw) {
        mbar_expect_tx(barw, w_stage_bytes);
        const __nv_bfloat16* src = Wp + (size_t)stage * w_full_stage;
        if (NP == NPfull) {
            bulk_g2s(dst, src, w_stage_bytes, barw);
        } else {
#pragma unroll
            for (int pc = 0; pc < 2 * (KC / 8); ++pc)
                bulk_g2s(dst + pc * NP * 16, src + ((size_t)pc * NPfull + n0) * 8, (uint32_t)NP * 16, barw);
        }
    };
    uint8_t* sA[2] = {smem, smem + A_STAGE_BYTES};
    uint8_t* sWbase = smem + 2 * A_STAGE_BYTES;          // NW weight slots of w_stage_bytes each
    const uint32_t bar_m[2] = {smem_u32(&s_bar[0]), smem_u32(&s_bar[1])};
    const uint32_t bar_done = smem_u32(&s_bar[2]);
    const uint32_t bar_w0 = smem_u32(&s_bar[3]), bar_f0 = smem_u32(&s_bar[3 + NW]);

    if (tid == 0) {
        for (int i = 0; i < 3 + 2 * 4; ++i) mbar_init(smem_u32(&s_bar[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // first weight stage can fly while TMEM is allocated and the rows are resolved
        for (int j = 0; j < NW - 1 && j < n_stage; ++j)        // NW-1 weight stages of look-ahead
            fetch_w(j, smem_u32(sWbase) + j * w_stage_bytes, bar_w0 + 8 * j);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }

    // ---- per-thread row bookkeeping ------------------------------------------------------------------------
    const long long r = (long long)blockIdx.x * TM + tid;
    const bool rvalid = r < rows;
    RowSrc rs;
    int run = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        rs.p[s] = nullptr; rs.sc[s] = 1.f; rs.ch[s] = 0;
        rs.c0[s] = 0x7fffffff;                      // unused segments never match a chunk index
        if (s < in.n_seg) {
            const hrn_seg_t sg = in.seg[s];
            rs.c0[s] = run;
            run += (sg.channels + 7) >> 3;
            rs.ch[s] = sg.channels;
            if (rvalid) {
                rs.p[s] = sg.ptr + hrn_src_row(in, sg.mode, r) * sg.ld + sg.col0;
                if (sg.row_scale) rs.sc[s] = __ldg(sg.row_scale + r);
            }
        }
    }
    rs.c0[4] = run;                                 // total number of 8-wide chunks

    float4 cur[KC / 4];
    float csc[KC / 8];
#pragma unroll
    for (int c = 0; c < KC / 8; ++c) load_chunk<FAST>(rs, c, cur[2 * c], cur[2 * c + 1], csc[c]);

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    // instruction descriptor: D=f32, A=B=bf16, K-major both, N, M=128
    const int NH = NP > 256 ? 256 : NP;                       // columns per MMA
    const int n_half = NP > 256 ? 2 : 1;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NH >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

    for (int i = 0; i < n_stage; ++i) {
        const int s = i & 1;
        if (i >= 2) mbar_wait(bar_m[s], ((i >> 1) - 1) & 1);          // MMAs of chunk i-2 have drained this stage
        // ---- A operand: split this thread's 32 fp32 into bf16 hi/lo core-matrix rows ------------------------
        uint4* a_hi = reinterpret_cast<uint4*>(sA[s]);
        uint4* a_lo = a_hi + (KC / 8) * TM;
#pragma unroll
        for (int c = 0; c < KC / 8; ++c) {
            const float sc = csc[c];
            const float x[8] = {cur[2 * c].x * sc, cur[2 * c].y * sc, cur[2 * c].z * sc, cur[2 * c].w * sc,
                                cur[2 * c + 1].x * sc, cur[2 * c + 1].y * sc, cur[2 * c + 1].z * sc, cur[2 * c + 1].w * sc};
            split_store8(x, a_hi + c * TM + tid, a_lo + c * TM + tid);
        }
        // ---- prefetch the next stage's row data into registers (in flight behind the barrier + MMA issue) ---
        if (i + 1 < n_stage) {
#pragma unroll
            for (int c = 0; c < KC / 8; ++c)
                load_chunk<FAST>(rs, (i + 1) * (KC / 8) + c, cur[2 * c], cur[2 * c + 1], csc[c]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            const int ws = i % NW;
            mbar_wait(bar_w0 + 8 * ws, (i / NW) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_base = smem_u32(sA[s]);
            const uint32_t w_base = smem_u32(sWbase) + ws * w_stage_bytes;
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
            umma_commit(bar_f0 + 8 * ws);
            if (i == n_stage - 1) umma_commit(bar_done);
            // weights of the next stage: its buffer is free once the MMAs of stage i-1 have drained
            // weights of stage i+NW-1 go into the slot of stage i-1, free once that stage's MMAs have drained
            const int nx = i + NW - 1;
            if (nx < n_stage) {
                const int fs = nx % NW;
                if (i >= 1) mbar_wait(bar_f0 + 8 * fs, ((i - 1) / NW) & 1);
                fetch_w(nx, smem_u32(sWbase) + fs * w_stage_bytes, bar_w0 + 8 * fs);
            }
        }
    }
    // ---- epilogue ---------------------------------------------------------------------------------------------
    mbar_wait(bar_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* yrow = Y + r * ldy + n0;
    const float* brow = bias + n0;
    const int ncols = min(NP, Cout - n0);                               // valid output columns of this slice
    for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        if (rvalid) {
            if (c0 + 32 <= ncols && ((reinterpret_cast<uintptr_t>(yrow + c0) & 15) == 0)) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 o;
                    o.x = act_fn(__uint_as_float(v[j + 0]) + __ldg(brow + c0 + j + 0), act);
                    o.y = act_fn(__uint_as_float(v[j + 1]) + __ldg(brow + c0 + j + 1), act);
                    o.z = act_fn(__uint_as_float(v[j + 2]) + __ldg(brow + c0 + j + 2), act);
                    o.w = act_fn(__uint_as_float(v[j + 3]) + __ldg(brow + c0 + j + 3), act);
                    *reinterpret_cast<float4*>(yrow + c0 + j) = o;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (c0 + j < ncols) yrow[c0 + j] = act_fn(__uint_as_float(v[j]) + __ldg(brow + c0 + j), act);
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
    // few row tiles (per-keypoint heads, coarse level): split the output columns over blockIdx.y so that the grid
    // still covers the 148 SMs
    const int tiles = hrn_divup(rows, TM);
    int NS = NP;
    while (NS > 64 && (NS / 2) % 16 == 0 && tiles * (NP / NS) < 2 * 148) NS >>= 1;
    const int n_split = NP / NS;
    int tmem_cols = 32;
    while (tmem_cols < NS) tmem_cols <<= 1;
    int NW = (int)((226 * 1024 - 2 * A_STAGE_BYTES) / ((size_t)NS * 128));   // weight ring depth
    NW = 2;   // measured: a 3- or 4-slot weight ring costs more in co-resident CTAs than it hides in copy latency
    const size_t smem = 2 * (size_t)A_STAGE_BYTES + (size_t)NW * NS * 128;
    static bool attr_set = false;
    if (!attr_set) {
        HRN_CUDA(cudaFuncSetAttribute(layer_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * A_STAGE_BYTES + 3 * 512 * 128));
        HRN_CUDA(cudaFuncSetAttribute(layer_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * A_STAGE_BYTES + 3 * 512 * 128));
        attr_set = true;
    }
    bool fast = true;   // 16-byte vector path: every segment 4-float aligned
    for (int s = 0; s < in->n_seg; ++s) {
        const hrn_seg_t& g = in->seg[s];
        if ((g.channels & 3) || (g.ld & 3) || (g.col0 & 3) || ((uintptr_t)g.ptr & 15)) fast = false;
    }
    dim3 grid(tiles, n_split);
    if (fast)
        layer_tc_kernel<true><<<grid, TM, smem, (cudaStream_t)stream>>>(
            *in, (const __nv_bfloat16*)Wp, bias, act, Y, ldy, rows, Cout, NP, NS, n_stage, tmem_cols, NW);
    else
        layer_tc_kernel<false><<<grid, TM, smem, (cudaStream_t)stream>>>(
            *in, (const __nv_bfloat16*)Wp, bias, act, Y, ldy, rows, Cout, NP, NS, n_stage, tmem_cols, NW);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}
