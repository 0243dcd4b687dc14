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

// ---------------------------------------------------------------------------------------------------------------
// Warp-specialised persistent variant (one CTA per SM, all 512 TMEM columns, ~200 KB of shared memory):
//
//   warps 0-3   epilogue   TMEM lane quadrant = warp; drain accumulator b while the MMAs fill accumulator b^1
//   warps 4-11  producers  thread <-> (row, k-half): gather 16 fp32 of the row per stage (two stages of loads in
//                          flight in registers), split to bf16 hi/lo, write the UMMA core-matrix layout
//   warp 12     MMA        one thread issues tcgen05.mma for every stage as soon as its operands have landed
//   warp 13     weights    one thread streams the packed weight slab of every stage with cp.async.bulk
//
// A work item is (128-row tile, <=256-column block); CTA c takes items c, c + grid, ...  Each ring stage holds the
// A slab (16 KB) and the weight slab (NS * 128 B) of one 32-wide K chunk:  full[s] collects 8 producer-warp arrivals
// + the bulk-copy transaction bytes, empty[s] is armed by tcgen05.commit.  Nothing in the loop is a CTA-wide barrier:
// the gather latency, the weight stream, the MMAs and the previous item's epilogue all overlap.
constexpr int WS_EPI_WARPS = 4, WS_PROD_WARPS = 8;
constexpr int WS_THREADS = (WS_EPI_WARPS + WS_PROD_WARPS + 2) * 32;   // 448
constexpr int WS_MAX_STAGES = 8;
constexpr int WS_TP = 36;                                              // epilogue transpose tile pitch (floats)
constexpr int WS_TILE_BYTES = WS_EPI_WARPS * 32 * WS_TP * 4;        // 18 KB
constexpr int WS_RAW = 4;                                             // raw fp32 stages (cp.async) in flight per CTA

// SIMPLE: the input is one DIRECT segment without a row scale (every layer that reads the previous layer's output):
// no segment table, no per-row source resolution -- the producers' instruction stream is what bounds these launches.
// PREC 3: bf16 hi/lo operands (six MMAs per 32-wide stage); PREC 1: one fp16 plane (two MMAs per stage; the stage's hi
// planes are the only ones written and read, the weight slab of a stage is NS * 64 bytes).
template <bool SIMPLE, int PREC>
__global__ void __launch_bounds__(WS_THREADS, 1)
layer_ws_kernel(const hrn_rows_t in, const __nv_bfloat16* __restrict__ Wp, const float* __restrict__ bias, int act,
                float* __restrict__ Y, int ldy, long long rows, int Cout, int NPfull, int NS, int n_split, int n_stage,
                int S, int n_items, int gk) {
    // gk = 0: Y [rows, ldy] = act(W x + b).  gk = 8 / 16 / 32: Y [rows / gk, ldy] = max over each group of gk consecutive
    // rows of act(W x + b) -- the reference's x.max(dim=3) after the last descriptor layer (layers.py:208) taken in the
    // epilogue, so the per-row result never reaches HBM (launcher: act monotone, Cout % 32 == 0, aligned Y).
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_full[WS_MAX_STAGES], s_empty[WS_MAX_STAGES], s_accf[2], s_acce[2];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int WPL = PREC == 1 ? 1 : 2;                               // weight / operand planes
    const uint32_t w_bytes = (uint32_t)NS * 64u * WPL;
    const uint32_t stage_bytes = A_STAGE_BYTES + w_bytes;
    const uint32_t smem0 = smem_u32(smem);
    const uint32_t full0 = smem_u32(&s_full[0]), empty0 = smem_u32(&s_empty[0]);
    const uint32_t accf0 = smem_u32(&s_accf[0]), acce0 = smem_u32(&s_acce[0]);

    if (tid == 0) {
        for (int i = 0; i < WS_MAX_STAGES; ++i) { mbar_init(full0 + 8 * i, WS_PROD_WARPS + 1); mbar_init(empty0 + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(accf0 + 8 * i, 1); mbar_init(acce0 + 8 * i, WS_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    if (warp < WS_EPI_WARPS) {
        // ================= epilogue: accumulator -> +bias -> activation -> HBM ==================================
        int ph = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++ph) {
            const int m = it / n_split, n0 = (it - m * n_split) * NS, b = ph & 1;
            const long long r = (long long)m * TM + warp * 32 + lane;
            const bool rvalid = r < rows;
            float* yrow = Y + r * ldy + n0;
            const float* brow = bias + n0;
            const int ncols = min(NS, Cout - n0);
            mbar_wait(accf0 + 8 * b, (ph >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const bool vec_ok = ((reinterpret_cast<uintptr_t>(Y + n0) | reinterpret_cast<uintptr_t>(brow)) & 15) == 0 && (ldy & 3) == 0;
            float* tile = reinterpret_cast<float*>(smem + (size_t)S * stage_bytes) + warp * (32 * WS_TP);
            for (int c0 = 0; c0 < ncols; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + b * 256 + c0, v);
                if (c0 + 32 <= ncols && vec_ok) {
                    // transpose through shared memory so that a store instruction covers 4 rows x 128 contiguous bytes
                    // (4 cache lines per request instead of 32 with one row per lane)
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<uint4*>(tile + lane * WS_TP + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    __syncwarp();
                    const int piece = lane & 7;
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(brow + c0) + piece);
                    if (gk) {
                        // lane (piece, rsub = lane >> 3) holds rows rsub + 4 i of the warp's 32: the group of row rl is
                        // rl / gk = i / (gk / 4).  Local maximum per group, then over the 4 rsub lanes (shuffles 8, 16);
                        // max_i act(x_i + b) = act(max_i x_i + b) exactly (rounding and act are monotone).
                        const int ipg = gk >> 2;                            // i's per group: 2, 4 or 8
                        const long long g0 = ((long long)m * TM + warp * 32) / gk;
                        float4 mx[4];                                      // up to 4 groups per warp (gk = 8)
#pragma unroll
                        for (int g = 0; g < 4; ++g) mx[g] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int rl = (lane >> 3) + 4 * i;
                            const float4 t = *reinterpret_cast<const float4*>(tile + rl * WS_TP + piece * 4);
                            const bool ok = (long long)m * TM + warp * 32 + rl < rows;
                            const int g = i / ipg;
#pragma unroll
                            for (int gg = 0; gg < 4; ++gg)
                                if (gg == g && ok) {
                                    mx[gg].x = fmaxf(mx[gg].x, t.x); mx[gg].y = fmaxf(mx[gg].y, t.y);
                                    mx[gg].z = fmaxf(mx[gg].z, t.z); mx[gg].w = fmaxf(mx[gg].w, t.w);
                                }
                        }
                        const int ng = 32 / gk;                             // groups in this warp's 32 rows
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (g < ng) {
#pragma unroll
                                for (int o = 8; o <= 16; o <<= 1) {
                                    mx[g].x = fmaxf(mx[g].x, __shfl_xor_sync(0xffffffffu, mx[g].x, o));
                                    mx[g].y = fmaxf(mx[g].y, __shfl_xor_sync(0xffffffffu, mx[g].y, o));
                                    mx[g].z = fmaxf(mx[g].z, __shfl_xor_sync(0xffffffffu, mx[g].z, o));
                                    mx[g].w = fmaxf(mx[g].w, __shfl_xor_sync(0xffffffffu, mx[g].w, o));
                                }
                                if ((lane >> 3) == (g & 3) && (g0 + g) * gk < rows) {   // spread the stores over the rsub lanes
                                    float4 o4;
                                    o4.x = act_fn(mx[g].x + bb.x, act); o4.y = act_fn(mx[g].y + bb.y, act);
                                    o4.z = act_fn(mx[g].z + bb.z, act); o4.w = act_fn(mx[g].w + bb.w, act);
                                    *reinterpret_cast<float4*>(Y + (g0 + g) * ldy + n0 + c0 + piece * 4) = o4;
                                }
                            }
                        }
                        __syncwarp();
                        continue;
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int rl = (lane >> 3) + 4 * i;
                        const float4 t = *reinterpret_cast<const float4*>(tile + rl * WS_TP + piece * 4);
                        const long long rr = (long long)m * TM + warp * 32 + rl;
                        float4 o;
                        o.x = act_fn(t.x + bb.x, act); o.y = act_fn(t.y + bb.y, act);
                        o.z = act_fn(t.z + bb.z, act); o.w = act_fn(t.w + bb.w, act);
                        if (rr < rows) *reinterpret_cast<float4*>(Y + rr * ldy + n0 + c0 + piece * 4) = o;
                    }
                    __syncwarp();
                } else if (rvalid) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c0 + j < ncols) yrow[c0 + j] = act_fn(__uint_as_float(v[j]) + __ldg(brow + c0 + j), act);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(acce0 + 8 * b);
        }
    } else if (warp < WS_EPI_WARPS + WS_PROD_WARPS) {
        // ================= producers: gather + split the A operand ==============================================
        // Coalesced gather: a warp owns 16 rows of the tile as two groups of 8.  Per copy instruction lane l moves
        // 16 bytes of row (l & 7) -> 8 rows x 64 contiguous bytes per request (8 cache lines; one row per lane would
        // be 32 and saturates the L1 tag pipeline).  The copies are cp.async (LDGSTS, L2 -> shared memory, no
        // registers, no scoreboard): WS_RAW stages of them stay in flight per thread across item boundaries, which is
        // what covers the HBM latency.  (Register prefetching cannot: ptxas puts every stage's loads on one scoreboard,
        // so waiting for the oldest stage waits for the newest.)  Each lane later reads back exactly the 16-byte slots
        // it copied, so cp.async.wait_group is the only synchronisation the raw ring needs.
        // Lanes l and l ^ 8 hold the two halves of one 8-wide core-matrix row: they swap bf16 halves so that one
        // stores the hi row, the other the lo row (16-byte stores, 8 consecutive rows per quarter warp: conflict-free).
        const int pw = warp - WS_EPI_WARPS;
        const int rsub = lane & 7, hf = (lane >> 3) & 1, cl = lane >> 4;
        int c0s[5], chs[4];                                    // segment table: first 8-wide chunk, channels
        {
            int run = 0;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                chs[s] = 0; c0s[s] = 0x7fffffff;
                if (s < in.n_seg) { c0s[s] = run; run += (in.seg[s].channels + 7) >> 3; chs[s] = in.seg[s].channels; }
            }
            c0s[4] = run;
        }
        int direct_mask = 0;
#pragma unroll
        for (int s = 0; s < 4; ++s) if (s < in.n_seg && in.seg[s].mode == HRN_SEG_DIRECT) direct_mask |= 1 << s;
        const float* rp[2][4];                                 // row base pointers of this lane's two rows
        float rsc[2][4];
        int st = 0; uint32_t par = 0;                          // ring position of the next stage to fill
        float sc[WS_RAW][4];                                   // per-piece row scale of the stages in flight
        int lit = blockIdx.x, li = 0;                          // copy cursor: item, stage within the item
        bool need_resolve = false;                             // the rows of item `lit` have to be resolved before its first copy
        uint8_t* raw0 = smem + (size_t)S * stage_bytes + WS_TILE_BYTES + (size_t)(pw * 4) * 512 + lane * 16;
        const bool small = rows < 0x7fffffffLL;
        auto resolve = [&](int it) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const long long r = (long long)(it / n_split) * TM + pw * 16 + g * 8 + rsub;
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    rp[g][s] = nullptr; rsc[g][s] = 1.f;
                    if (SIMPLE) {
                        if (s == 0 && r < rows) rp[g][0] = in.seg[0].ptr + r * in.seg[0].ld + in.seg[0].col0;
                    } else if (s < in.n_seg && r < rows) {
                        const hrn_seg_t sg = in.seg[s];
                        long long sr;
                        if (small) {                               // 32-bit divisions
                            const unsigned ru = (unsigned)r;
                            sr = sg.mode == HRN_SEG_DIRECT ? (long long)ru
                               : sg.mode == HRN_SEG_BROADCAST ? (long long)(ru / (unsigned)in.group)
                               : (long long)(ru / (unsigned)in.rows_per_batch) * in.src_rows_per_batch + in.gather_idx[r];
                        } else {
                            sr = hrn_src_row(in, sg.mode, r);
                        }
                        rp[g][s] = sg.ptr + sr * sg.ld + sg.col0;
                        if (sg.row_scale) rsc[g][s] = __ldg(sg.row_scale + r);
                    }
                }
            }
        };
        // async copy of the 4 floats [4 hf, 4 hf + 4) of 8-wide chunk cg of one row (zero-filled out of range)
        auto copy_piece = [&](const float* const (&pp)[4], const float (&ps)[4], int cg, uint8_t* dst, float& osc) {
            if (SIMPLE) {
                const int ch0 = (cg << 3) + 4 * hf;
                const bool ok = pp[0] != nullptr && ch0 < chs[0];
                const void* src = ok ? (const void*)(pp[0] + ch0) : (const void*)Wp;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(ok ? 16 : 0) : "memory");
                osc = 1.f;
                return;
            }
            int sgi = 0;
#pragma unroll
            for (int q = 1; q < 4; ++q) if (cg >= c0s[q]) sgi = q;
            const float* p = pp[0]; int cs = c0s[0], chn = chs[0]; osc = ps[0];
#pragma unroll
            for (int q = 1; q < 4; ++q) if (sgi == q) { p = pp[q]; cs = c0s[q]; chn = chs[q]; osc = ps[q]; }
            const int ch0 = ((cg - cs) << 3) + 4 * hf;
            const bool ok = p != nullptr && cg < c0s[4] && chn - ch0 >= 4;
            const void* src = ok ? (const void*)(p + ch0) : (const void*)Wp;
            // streamed rows bypass L1; broadcast / gathered rows are shared by the rows of a group -> keep them in L1
            if (direct_mask >> sgi & 1)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(ok ? 16 : 0) : "memory");
            else
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(ok ? 16 : 0) : "memory");
        };
        auto issue = [&](int slot, float (&ss)[4]) {
            if (lit < n_items) {
#pragma unroll
                for (int g = 0; g < 2; ++g)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        copy_piece(rp[g], rsc[g], li * 4 + 2 * j + cl, raw0 + (size_t)slot * A_STAGE_BYTES + (g * 2 + j) * 512, ss[2 * g + j]);
                if (++li == n_stage) {
                    li = 0; lit += gridDim.x;
                    need_resolve = lit < n_items;        // done at the top of the stage loop: ONE inlined copy of resolve()
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto fill = [&](int slot, const float (&ss)[4]) {
            asm volatile("cp.async.wait_group %0;" ::"n"(WS_RAW - 1) : "memory");
            float4 vv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) vv[e] = *reinterpret_cast<const float4*>(raw0 + (size_t)slot * A_STAGE_BYTES + e * 512);
            mbar_wait(empty0 + 8 * st, par ^ 1);
            uint4* a_hi = reinterpret_cast<uint4*>(smem + (size_t)st * stage_bytes);
            uint4* a_lo = a_hi + (KC / 8) * TM;
#pragma unroll
            for (int g = 0; g < 2; ++g)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float4 t = vv[2 * g + j];
                    const float s_ = SIMPLE ? 1.f : ss[2 * g + j];
                    const float x0 = SIMPLE ? t.x : t.x * s_, x1 = SIMPLE ? t.y : t.y * s_;
                    const float x2 = SIMPLE ? t.z : t.z * s_, x3 = SIMPLE ? t.w : t.w * s_;
                    const int slot_a = (2 * j + cl) * TM + pw * 16 + g * 8 + rsub;
                    if (PREC == 1) {           // each lane stores its own 8 bytes of the 16-byte core-matrix row
                        *reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(a_hi + slot_a) + 8 * hf) =
                            make_uint2(pack_f16x2(x0, x1), pack_f16x2(x2, x3));
                        continue;
                    }
                    uint32_t H0, H1, L0, L1;
                    split_pair(x0, x1, H0, L0);
                    split_pair(x2, x3, H1, L1);
                    const uint32_t r0 = __shfl_xor_sync(0xffffffffu, hf ? H0 : L0, 8);
                    const uint32_t r1 = __shfl_xor_sync(0xffffffffu, hf ? H1 : L1, 8);
                    if (hf == 0) a_hi[slot_a] = make_uint4(H0, H1, r0, r1);
                    else         a_lo[slot_a] = make_uint4(r0, r1, L0, L1);
                }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(full0 + 8 * st);
            if (++st == S) { st = 0; par ^= 1; }
        };
        int my_items = 0;
        if ((int)blockIdx.x < n_items) my_items = (n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1;
        const int total = my_items * n_stage;
        need_resolve = total > 0;
        // one copy of the stage code (run-time slot; unrolled over the raw ring, prologue + main loop, it dominated the
        // kernel's instruction footprint -- see chain_tc.cu)
#pragma unroll 1
        for (int d = -WS_RAW; d < total; ++d) {
            const int slot = (d + WS_RAW) % WS_RAW;
            if (need_resolve) { resolve(lit); need_resolve = false; }      // index / scale loads in front of the wait inside fill()
            if (d >= 0) fill(slot, sc[slot]);
            issue(slot, sc[slot]);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else if (warp == WS_EPI_WARPS + WS_PROD_WARPS) {
        // ================= MMA issue ==============================================================================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_m128<PREC>(NS);
            const uint32_t w_lbo = (uint32_t)NS * 16;
            const uint64_t a_fix = umma_desc_fixed(TM * 16, 128), w_fix = umma_desc_fixed(w_lbo, 128);
            // offsets inside a stage, 16-byte units: lo planes, second k-step
            constexpr uint32_t A_LO = ((KC / 8) * TM * 16) >> 4, A_K1 = (2 * TM * 16) >> 4;
            const uint32_t w_lo = ((KC / 8) * w_lbo) >> 4, w_k1 = (2 * w_lbo) >> 4;
            const uint32_t stage16 = stage_bytes >> 4;
            int st = 0; uint32_t par = 0; int ph = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++ph) {
                const int b = ph & 1;
                mbar_wait(acce0 + 8 * b, ((ph >> 1) & 1) ^ 1);             // epilogue has drained this accumulator
                const uint32_t d = tmem + b * 256;
                for (int i = 0; i < n_stage; ++i) {
                    mbar_wait(full0 + 8 * st, par);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a16 = (smem0 >> 4) + st * stage16, w16 = a16 + (A_STAGE_BYTES >> 4);
                    umma_bf16(d, a_fix | a16, w_fix | w16, idesc, i > 0 ? 1u : 0u);
                    if (PREC == 3) {
                        umma_bf16(d, a_fix | (a16 + A_LO), w_fix | w16, idesc, 1u);
                        umma_bf16(d, a_fix | a16, w_fix | (w16 + w_lo), idesc, 1u);
                    }
                    umma_bf16(d, a_fix | (a16 + A_K1), w_fix | (w16 + w_k1), idesc, 1u);
                    if (PREC == 3) {
                        umma_bf16(d, a_fix | (a16 + A_K1 + A_LO), w_fix | (w16 + w_k1), idesc, 1u);
                        umma_bf16(d, a_fix | (a16 + A_K1), w_fix | (w16 + w_k1 + w_lo), idesc, 1u);
                    }
                    umma_commit(empty0 + 8 * st);
                    if (++st == S) { st = 0; par ^= 1; }
                }
                umma_commit(accf0 + 8 * b);
            }
        }
    } else {
        // ================= weight stream ==========================================================================
        if (lane == 0) {
            const size_t w_full_stage = (size_t)NPfull * 32 * WPL;          // 16-bit elements per packed stage
            int st = 0; uint32_t par = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
                const int n0 = (it % n_split) * NS;
                for (int i = 0; i < n_stage; ++i) {
                    mbar_wait(empty0 + 8 * st, par ^ 1);
                    const uint32_t barw = full0 + 8 * st;
                    const uint32_t dst = smem0 + st * stage_bytes + A_STAGE_BYTES;
                    const __nv_bfloat16* src = Wp + (size_t)i * w_full_stage;
                    mbar_expect_tx(barw, w_bytes);
                    if (NS == NPfull) {
                        bulk_g2s(dst, src, w_bytes, barw);
                    } else {
#pragma unroll
                        for (int pc = 0; pc < WPL * (KC / 8); ++pc)
                            bulk_g2s(dst + pc * NS * 16, src + ((size_t)pc * NPfull + n0) * 8, (uint32_t)NS * 16, barw);
                    }
                    if (++st == S) { st = 0; par ^= 1; }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <bool SIMPLE, int PREC>
int layer_ws_launch_one(const hrn_rows_t* in, const void* Wp, const float* bias, int act, float* Y, int ldy, long long rows,
                        int Cout, int NP, int NS, int n_split, int n_stage, int S, int n_items, int gk, int grid, size_t smem,
                        size_t budget, cudaStream_t stream) {
    static hrn_once_per_device attr_set;
    if (attr_set.need())
        HRN_CUDA(cudaFuncSetAttribute(layer_ws_kernel<SIMPLE, PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    layer_ws_kernel<SIMPLE, PREC><<<grid, WS_THREADS, smem, stream>>>(*in, (const __nv_bfloat16*)Wp, bias, act, Y, ldy, rows, Cout,
                                                                    NP, NS, n_split, n_stage, S, n_items, gk);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}

int layer_ws_launch(const hrn_rows_t* in, const void* Wp, const float* bias, int act, float* Y, int ldy, long long rows,
                    int Cout, int NP, int n_stage, cudaStream_t stream, int gk = 0, int prec = 3) {
    const int tiles = hrn_divup(rows, TM);
    int NS = NP > 256 ? 256 : NP;
    while (NS > 32 && (NS / 2) % 16 == 0 && tiles * (NP / NS) < 148) NS >>= 1;
    const int n_split = NP / NS;
    const int n_items = tiles * n_split;
    const size_t stage_bytes = (size_t)A_STAGE_BYTES + (size_t)NS * (prec == 1 ? 64 : 128);
    const size_t fixed = WS_TILE_BYTES + (size_t)WS_RAW * A_STAGE_BYTES;
    const size_t budget = 226 * 1024;
    int S = (int)((budget - fixed) / stage_bytes);
    if (S > WS_MAX_STAGES) S = WS_MAX_STAGES;
    const size_t smem = (size_t)S * stage_bytes + fixed;
    const int grid = n_items < 148 ? n_items : 148;
    const bool simple = in->n_seg == 1 && in->seg[0].mode == HRN_SEG_DIRECT && !in->seg[0].row_scale;
#define HRN_LWS(SIMPLE_, PREC_) \
    layer_ws_launch_one<SIMPLE_, PREC_>(in, Wp, bias, act, Y, ldy, rows, Cout, NP, NS, n_split, n_stage, S, n_items, gk, grid, smem, budget, stream)
    if (prec == 1) return simple ? HRN_LWS(true, 1) : HRN_LWS(false, 1);
    return simple ? HRN_LWS(true, 3) : HRN_LWS(false, 3);
#undef HRN_LWS
}

}  // namespace

// Wp = weights packed by hrn_pack_weights_layout (see engine_tc.pack_weights): [n_stage][2][4][NP][8] bf16.
// K_pad = 32 * n_stage must equal sum over segments of ceil8(channels), rounded up to 32.
HRN_API int hrn_layer_tc(const hrn_rows_t* in, const void* Wp, const float* bias, int act, float* Y, int ldy,
                         long long rows, int Cout, int NP, int n_stage, int prec, void* stream) {
    if (!in || !Wp || !bias || !Y || rows < 0 || Cout <= 0 || in->n_seg < 1 || in->n_seg > 4) return HRN_ERR_BAD_ARG;
    if (prec != 1 && prec != 3) return HRN_ERR_BAD_ARG;
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
    bool fast = true;   // 16-byte vector path: every segment 4-float aligned
    for (int s = 0; s < in->n_seg; ++s) {
        const hrn_seg_t& g = in->seg[s];
        if ((g.channels & 3) || (g.ld & 3) || (g.col0 & 3) || ((uintptr_t)g.ptr & 15)) fast = false;
    }
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
    static hrn_once_per_device attr_set;
    if (attr_set.need()) {
        HRN_CUDA(cudaFuncSetAttribute(layer_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * A_STAGE_BYTES + 3 * 512 * 128));
        HRN_CUDA(cudaFuncSetAttribute(layer_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * A_STAGE_BYTES + 3 * 512 * 128));
    }
    // 16-byte aligned segments (every call of the registration path): persistent warp-specialised kernel; the
    // one-CTA-per-tile kernel below remains as the general fallback (odd channel counts, unaligned views)
    if (fast)
        return layer_ws_launch(in, Wp, bias, act, Y, ldy, rows, Cout, NP, n_stage, (cudaStream_t)stream, 0, prec);
    if (prec != 3) return HRN_ERR_UNSUPPORTED;        // the general fallback kernel exists for the bf16 hi/lo operands only
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

// Same layer with the group maximum taken in the epilogue: G [rows / k, ldg] = max over each k consecutive rows of
// act(W x + b)  (k = 8, 16 or 32; act = none / ReLU; Cout a multiple of 32; 16-byte aligned segments, G and bias).
// Replaces hrn_layer_tc + hrn_group_max for the last DescExtractor layer (reference layers.py:207-208).
HRN_API int hrn_layer_tc_groupmax(const hrn_rows_t* in, const void* Wp, const float* bias, int act, float* G, int ldg,
                                  long long rows, int Cout, int NP, int n_stage, int k, int prec, void* stream) {
    if (!in || !Wp || !bias || !G || rows < 0 || Cout <= 0 || in->n_seg < 1 || in->n_seg > 4) return HRN_ERR_BAD_ARG;
    if (prec != 1 && prec != 3) return HRN_ERR_BAD_ARG;
    if (k != 8 && k != 16 && k != 32) return HRN_ERR_UNSUPPORTED;
    if (act != HRN_ACT_NONE && act != HRN_ACT_RELU) return HRN_ERR_UNSUPPORTED;
    if (rows % k != 0 || Cout % 32 != 0 || NP != Cout || NP > 512 || (NP > 256 && NP != 512)) return HRN_ERR_UNSUPPORTED;
    if ((ldg & 3) || ((uintptr_t)G & 15) || ((uintptr_t)bias & 15)) return HRN_ERR_UNSUPPORTED;
    int chunks = 0;
    for (int s = 0; s < in->n_seg; ++s) {
        const hrn_seg_t& g = in->seg[s];
        if (!g.ptr || g.channels <= 0) return HRN_ERR_BAD_ARG;
        if (g.mode == HRN_SEG_GATHER && !in->gather_idx) return HRN_ERR_BAD_ARG;
        if (g.mode == HRN_SEG_BROADCAST && in->group <= 0) return HRN_ERR_BAD_ARG;
        if ((g.channels & 3) || (g.ld & 3) || (g.col0 & 3) || ((uintptr_t)g.ptr & 15)) return HRN_ERR_UNSUPPORTED;
        chunks += (g.channels + 7) / 8;
    }
    if (n_stage != (chunks * 8 + KC - 1) / KC) return HRN_ERR_BAD_ARG;
    if (rows == 0) return HRN_OK;
    return layer_ws_launch(in, Wp, bias, act, G, ldg, rows, Cout, NP, n_stage, (cudaStream_t)stream, k, prec);
}
