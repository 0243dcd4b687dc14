// Three-layer shared-MLP chain on tcgen05 with the activations kept on chip.
//
//     X (virtual rows, rows.cuh)  ->  relu(W1 . + b1)  ->  relu(W2 . + b2)  ->  relu(W3 . + b3)  ->  epilogue
//
// Replaces the conv stacks `convs` / `convs_1` / `convs_2` of the reference (models/HRegNet/layers.py:118-121,
// 249-260, 420-423: three  Conv2d(1x1, bias=False) + BatchNorm2d + ReLU  triples) where the widths are <= 256, plus
// the reductions that consume them:
//   EPI_STORE     Y rows                                                    (descriptor stack, layers.py:201)
//   EPI_GROUPMAX  Y rows and max over the k rows of each group              (layers.py:202)
//   EPI_ATTN      a = softmax_k(max_c Y), AF[g,:] = sum_k a Y, optionally the rows Y*a
//                 (layers.py:150-159, 329-332, 384-390, 446-450)
// Per-layer kernels write and re-read every [rows, C] intermediate (3 launches, 6 HBM passes); here a tile of 128
// rows goes through all layers on chip: the first operand is gathered from the virtual rows, each layer's result goes
// TMEM -> registers -> bias+ReLU -> bf16 hi/lo split -> shared memory as the next layer's A operand, and the weights
// stream from L2 in K=16 pieces through a cp.async.bulk ring.  bf16x3 products, fp32 accumulation, same numerics as
// mlp_tc.cu.  The kernel is persistent and warp-specialised (see chain_ws_kernel below).
#include "common.cuh"
#include "tc_common.cuh"
#include <math_constants.h>

namespace {

constexpr int CTM = 128;                 // rows per tile (UMMA M)

enum { EPI_STORE = 0, EPI_GROUPMAX = 1, EPI_ATTN = 2 };

struct ChainArgs {
    hrn_rows_t in;
    const uint8_t* W;        // packed pieces of the three layers, execution order
    const float* bias;       // b1 | b2 | b3
    float* Y;                // rows x N3 (nullable for EPI_ATTN)
    float* G;                // groups x N3: group max (EPI_GROUPMAX) or attentive feature (EPI_ATTN)
    float* a;                // rows: attention weights (EPI_ATTN)
    long long rows;
    int ldy;
    int n[3];                // layer widths as issued (multiples of 16, <= 256); n[nl-1] = padded last width
    int nl;                  // number of layers (2 or 3)
    int cout;                // real output columns of the last layer (<= n[nl-1])
    int act;                 // activation of the last layer (HRN_ACT_*); hidden layers are ReLU
    int slot_bytes;          // weight ring slot = widest layer x 64 B (one K=16 piece)
    int ring;                // weight ring slots
    int chunks0;             // 8-wide K chunks of the virtual input (segments padded to 8, total padded to even)
    int mode;
    int kseg;                // rows per group (8, 16 or 32)
    // optional fp32 rows ADDED to the first layer's pre-activations -- the part of that (linear) layer which is constant
    // inside a group (Zb [rows / kseg, ldz], nullable) / depends on the gathered source row only (Zg [src rows, ldz]),
    // applied once per point by the caller (engine_tc._split_first_layer); see chain_wide.cu
    const float* Zb;
    const float* Zg;
    int ldz;
};

__device__ __forceinline__ uint32_t ch_idesc(int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(CTM >> 4) << 24);
}

// segmented transpose-reduce: lanes form groups of KSEG consecutive lanes; after the call, the lane at position p of
// its group holds in v[0 .. 32/KSEG) the reduction over the group of columns p*(32/KSEG) + i.
template <int KSEG, bool IS_MAX>
__device__ __forceinline__ void seg_transpose_reduce(float (&v)[32], int lane) {
    int cnt = 32;
#pragma unroll
    for (int off = KSEG / 2; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
        const int half = cnt / 2;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (i < half) {
                const float send = up ? v[i] : v[i + half];
                const float keep = up ? v[i + half] : v[i];
                const float recv = __shfl_xor_sync(0xffffffffu, send, off);
                v[i] = IS_MAX ? fmaxf(keep, recv) : keep + recv;
            }
        }
        cnt = half;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent warp-specialised chain (one CTA per SM, all 512 TMEM columns = two accumulators used alternately by
// consecutive layers).  No tile-sized operand buffer: both kinds of A operand travel through small rings of 16 KB
// stages (128 rows x 32 K, bf16 hi + lo), which leaves the shared memory to a deep weight ring.
//
//   warps 0-7   epilogue   (warp & 3 = TMEM lane quadrant, warp >> 2 = group; the groups take alternate 32-column
//                          blocks).  Hidden layers: accumulator -> relu(x + b) -> bf16 hi/lo -> ring H, one block (two
//                          K=16 pieces) at a time, so the NEXT layer's MMAs run while this layer is still being
//                          drained; last layer: the STORE / GROUPMAX / ATTN epilogue, overlapped with the next tile
//   warps 8-15  producers  coalesced cp.async gather of the virtual rows into a raw fp32 ring, then split to bf16
//                          hi/lo into ring G; they run ahead of the MMAs across tile boundaries
//   warp 16     MMA        issues every K=16 piece as soon as its operand stage and its weights have landed
//   warp 17     weights    streams the packed K=16 weight pieces of all layers through a cp.async.bulk ring
constexpr int CW_EPI_WARPS = 8, CW_PROD_WARPS = 8;                    // epilogue: two warps per TMEM lane quadrant
constexpr int CW_THREADS = (CW_EPI_WARPS + CW_PROD_WARPS + 2) * 32;   // 576
constexpr int CW_STAGE_BYTES = 4 * CTM * 32;                          // 4 chunks x 128 rows x (8 fp32 | 8 bf16 hi + lo) = 16 KB
constexpr int CW_RING_MAX = 8;                                        // weight slots
constexpr int CW_GS_MAX = 4, CW_HS_MAX = 4;                           // operand ring depths
constexpr int CW_ZS_MAX = 4;
constexpr int CW_ZG_BYTES = CTM * 32 * 4;                             // gathered rows of a 32-column block: 128 x 128 B
constexpr int CW_ZSTAGE_BYTES = CW_ZG_BYTES + (CTM / 8) * 32 * 4;     // + up to 16 group rows = 18 KB
constexpr int CW_TP = 36;
constexpr int CW_TILE_BYTES = CW_EPI_WARPS * 32 * CW_TP * 4;

// -DHRN_CHAIN_PROF: cycles the MMA thread of CTA 0 spends waiting for [0] input stages, [1] hidden-layer blocks, [2] weights,
// [3] a free accumulator, [4] total; read back with hrn_chain_prof (tools/chain_probe.py)
#ifdef HRN_CHAIN_PROF
__device__ long long g_chain_prof[8];
#define CPROF_BEGIN() const long long cp_t0 = clock64()
#define CPROF_END(i) cp[i] += clock64() - cp_t0
#else
#define CPROF_BEGIN() do { } while (0)
#define CPROF_END(i) do { } while (0)
#endif

struct ChainWsArgs {
    ChainArgs c;
    int n_tiles;
    int use_tile;            // transpose tile for coalesced Y stores present
    int gs, hs;              // depths of ring G (gathered input stages) and ring H (hidden-layer blocks)
    int acc_stride, nbuf;    // TMEM accumulators: nbuf = 512 / acc_stride buffers (4 x 128 or 2 x 256 columns), used round-robin
    int zs;                  // Z ring stages (0: no Z rows)
};

// PREC 3: bf16 hi/lo operands, three MMAs per K=16 piece; PREC 1: single fp16 plane, one MMA per piece (the hi plane of
// every stage / slot is the only one written and read).
template <int KSEG, int RAW, int PREC>
__global__ void __launch_bounds__(CW_THREADS, 1) chain_ws_kernel(const ChainWsArgs AW) {
    const ChainArgs& A = AW.c;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_wfull[CW_RING_MAX], s_wempty[CW_RING_MAX], s_gfull[CW_GS_MAX], s_gempty[CW_GS_MAX],
        s_hfull[CW_HS_MAX], s_hempty[CW_HS_MAX], s_accf[4], s_fin[4], s_zfull[CW_ZS_MAX], s_zempty[CW_ZS_MAX];
    __shared__ uint32_t s_tmem;

    const int RING = A.ring, GS = AW.gs, HS = AW.hs, ZS = AW.zs;
    const bool has_z = A.Zg != nullptr;
    const int ACC = AW.acc_stride, BM = AW.nbuf - 1;       // accumulator of layer L: columns (L & BM) * ACC
    const uint32_t SLOT_BYTES = (uint32_t)A.slot_bytes;
    uint8_t* sG = smem;
    uint8_t* sH = sG + GS * CW_STAGE_BYTES;
    uint8_t* sRing = sH + HS * CW_STAGE_BYTES;
    uint8_t* sRaw = sRing + RING * SLOT_BYTES;
    float* sBias = reinterpret_cast<float*>(sRaw + RAW * CW_STAGE_BYTES);
    float* sX = sBias + 3 * 256;                            // [2][128] row maxima exchanged by the epilogue groups
    float* sTile = sX + 2 * CTM;
    // Z ring behind everything else: [zs][128 gathered rows x 128 B (16-byte chunks XOR-swizzled by row) | 16 group rows x 128 B]
    uint8_t* sZ = reinterpret_cast<uint8_t*>(sTile) + (AW.use_tile ? CW_TILE_BYTES : 0);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nl = A.nl;
    const int n_tiles = AW.n_tiles;
    const int n_st0 = (A.chunks0 + 3) >> 2;                 // 32-wide input stages per tile
    const int cout = A.cout;

    if (tid == 0) {
        for (int i = 0; i < CW_RING_MAX; ++i) { mbar_init(smem_u32(&s_wfull[i]), 1); mbar_init(smem_u32(&s_wempty[i]), 1); }
        for (int i = 0; i < CW_GS_MAX; ++i) { mbar_init(smem_u32(&s_gfull[i]), CW_PROD_WARPS); mbar_init(smem_u32(&s_gempty[i]), 1); }
        for (int i = 0; i < CW_HS_MAX; ++i) { mbar_init(smem_u32(&s_hfull[i]), 4); mbar_init(smem_u32(&s_hempty[i]), 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(smem_u32(&s_accf[i]), 1); mbar_init(smem_u32(&s_fin[i]), CW_EPI_WARPS); }
        for (int i = 0; i < CW_ZS_MAX; ++i) { mbar_init(smem_u32(&s_zfull[i]), CW_PROD_WARPS * 32); mbar_init(smem_u32(&s_zempty[i]), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    const int nbias = A.n[0] + (nl == 3 ? A.n[1] : 0) + cout;
    for (int i = tid; i < nbias; i += CW_THREADS) sBias[i] = __ldg(A.bias + i);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t ring_a = smem_u32(sRing);
    const uint32_t pieces[3] = {(uint32_t)(A.chunks0 / 2), (uint32_t)(A.n[0] / 16), nl == 3 ? (uint32_t)(A.n[1] / 16) : 0u};

    if (warp < CW_EPI_WARPS) {
        // ================= epilogue warps ======================================================================
        const int eg = warp >> 2, wq = warp & 3;             // group (alternate blocks), TMEM lane quadrant
        const int rt = wq * 32 + lane;                       // row inside the tile = TMEM lane
        const uint32_t lane_base = ((uint32_t)(wq * 32) << 16);
        uint32_t accph = 0;                                  // phase bits of s_accf[4]
        int L = 0;                                           // layers completed by this CTA -> accumulator L & BM
        int hs = 0; uint32_t hpar = 0; int hq = 0;           // ring H position, global block counter
        int zs = 0; uint32_t zpar = 0;                       // Z ring position
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const long long r = (long long)tile * CTM + rt;
            for (int l = 0; l + 1 < nl; ++l, ++L) {
                const int b = L & BM, N = A.n[l];
                const float* bb = sBias + (l == 0 ? 0 : A.n[0]);
                mbar_wait(smem_u32(&s_accf[b]), (accph >> b) & 1); accph ^= 1u << b;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int c0 = 0; c0 < N; c0 += 32, ++hq) {
                    if ((hq & 1) == eg) {
                        uint32_t v[32];
                        tmem_ld32(tmem + lane_base + b * ACC + c0, v);
                        const bool zblk = has_z && l == 0;
                        const float* zg = reinterpret_cast<const float*>(sZ + (size_t)zs * CW_ZSTAGE_BYTES) + rt * 32;
                        const float* zb = reinterpret_cast<const float*>(sZ + (size_t)zs * CW_ZSTAGE_BYTES + CW_ZG_BYTES) + (rt / KSEG) * 32;
                        if (zblk) mbar_wait(smem_u32(&s_zfull[zs]), zpar);      // this block's Z rows have landed
                        mbar_wait(smem_u32(&s_hempty[hs]), hpar ^ 1);
                        uint4* h_hi = reinterpret_cast<uint4*>(sH + (size_t)hs * CW_STAGE_BYTES);
                        uint4* h_lo = h_hi + 4 * CTM;
#pragma unroll
                        for (int ch = 0; ch < 4; ++ch) {
                            if (c0 + ch * 8 < N) {
                                const float4 b0 = *reinterpret_cast<const float4*>(bb + c0 + ch * 8);
                                const float4 b1 = *reinterpret_cast<const float4*>(bb + c0 + ch * 8 + 4);
                                float s[8];                       // bias additions as packed fp32x2 adds
                                f2_unpack(f2_add(f2_pack(__uint_as_float(v[ch * 8 + 0]), __uint_as_float(v[ch * 8 + 1])), f2_pack(b0.x, b0.y)), s[0], s[1]);
                                f2_unpack(f2_add(f2_pack(__uint_as_float(v[ch * 8 + 2]), __uint_as_float(v[ch * 8 + 3])), f2_pack(b0.z, b0.w)), s[2], s[3]);
                                f2_unpack(f2_add(f2_pack(__uint_as_float(v[ch * 8 + 4]), __uint_as_float(v[ch * 8 + 5])), f2_pack(b1.x, b1.y)), s[4], s[5]);
                                f2_unpack(f2_add(f2_pack(__uint_as_float(v[ch * 8 + 6]), __uint_as_float(v[ch * 8 + 7])), f2_pack(b1.z, b1.w)), s[6], s[7]);
                                if (zblk) {       // + gathered row (+ group row), fp32
                                    const float4 g0 = *reinterpret_cast<const float4*>(zg + (((2 * ch) ^ (rt & 7)) << 2));
                                    const float4 g1 = *reinterpret_cast<const float4*>(zg + (((2 * ch + 1) ^ (rt & 7)) << 2));
                                    float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0;
                                    if (A.Zb) {
                                        q0 = *reinterpret_cast<const float4*>(zb + ch * 8);
                                        q1 = *reinterpret_cast<const float4*>(zb + ch * 8 + 4);
                                    }
                                    f2_unpack(f2_add(f2_add(f2_pack(s[0], s[1]), f2_pack(q0.x, q0.y)), f2_pack(g0.x, g0.y)), s[0], s[1]);
                                    f2_unpack(f2_add(f2_add(f2_pack(s[2], s[3]), f2_pack(q0.z, q0.w)), f2_pack(g0.z, g0.w)), s[2], s[3]);
                                    f2_unpack(f2_add(f2_add(f2_pack(s[4], s[5]), f2_pack(q1.x, q1.y)), f2_pack(g1.x, g1.y)), s[4], s[5]);
                                    f2_unpack(f2_add(f2_add(f2_pack(s[6], s[7]), f2_pack(q1.z, q1.w)), f2_pack(g1.z, g1.w)), s[6], s[7]);
                                }
                                if (PREC == 1) {
                                    h_hi[ch * CTM + rt] = make_uint4(pack_f16x2_relu(s[0], s[1]), pack_f16x2_relu(s[2], s[3]),
                                                                     pack_f16x2_relu(s[4], s[5]), pack_f16x2_relu(s[6], s[7]));
                                } else {
                                    const float x[8] = {fmaxf(s[0], 0.f), fmaxf(s[1], 0.f), fmaxf(s[2], 0.f), fmaxf(s[3], 0.f),
                                                        fmaxf(s[4], 0.f), fmaxf(s[5], 0.f), fmaxf(s[6], 0.f), fmaxf(s[7], 0.f)};
                                    split_store8(x, h_hi + ch * CTM + rt, h_lo + ch * CTM + rt);
                                }
                            }
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) {
                            mbar_arrive(smem_u32(&s_hfull[hs]));
                            if (zblk) mbar_arrive(smem_u32(&s_zempty[zs]));
                        }
                    }
                    if (++hs == HS) { hs = 0; hpar ^= 1; }
                    if (has_z && l == 0 && ++zs == ZS) { zs = 0; zpar ^= 1; }
                }
            }
            // ---- last layer: the groups take alternate 32-column chunks --------------------------------------------
            const int b = L & BM;
            ++L;
            mbar_wait(smem_u32(&s_accf[b]), (accph >> b) & 1); accph ^= 1u << b;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc = tmem + lane_base + b * ACC;
            const float* b3 = sBias + A.n[0] + (nl == 3 ? A.n[1] : 0);
            const int act = A.act;
            const int pos = lane % KSEG;
            const long long grp = r / KSEG;
            constexpr int PER = 32 / KSEG;
            float a_w = 1.f;
            if (A.mode == EPI_ATTN) {
                float x1 = -CUDART_INF_F;
                for (int c0 = eg * 32; c0 < cout; c0 += 64) {
                    uint32_t v[32];
                    tmem_ld32(acc + c0, v);
                    if (c0 + 32 <= cout) {
#pragma unroll
                        for (int e = 0; e < 32; e += 4) {
                            const float4 b4 = *reinterpret_cast<const float4*>(b3 + c0 + e);
                            float s0, s1, s2, s3;
                            f2_unpack(f2_add(f2_pack(__uint_as_float(v[e]), __uint_as_float(v[e + 1])), f2_pack(b4.x, b4.y)), s0, s1);
                            f2_unpack(f2_add(f2_pack(__uint_as_float(v[e + 2]), __uint_as_float(v[e + 3])), f2_pack(b4.z, b4.w)), s2, s3);
                            x1 = fmaxf(x1, fmaxf(fmaxf(s0, s1), fmaxf(s2, s3)));
                        }
                        x1 = fmaxf(x1, 0.f);
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; ++e)
                            if (c0 + e < cout) x1 = fmaxf(x1, fmaxf(__uint_as_float(v[e]) + b3[c0 + e], 0.f));
                    }
                }
                sX[eg * CTM + rt] = x1;
                asm volatile("bar.sync 1, 256;" ::: "memory");             // the 8 epilogue warps
                x1 = fmaxf(x1, sX[(eg ^ 1) * CTM + rt]);
                float gm = x1;
#pragma unroll
                for (int o = KSEG / 2; o > 0; o >>= 1) gm = fmaxf(gm, __shfl_xor_sync(0xffffffffu, gm, o));
                const float ex = expf(x1 - gm);
                float sm = ex;
#pragma unroll
                for (int o = KSEG / 2; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
                a_w = ex / sm;
                if (A.a && eg == 0) A.a[r] = a_w;
            }
            const bool vec_ok = A.Y && ((cout & 3) == 0);
            float* tile_w = sTile + warp * (32 * CW_TP);
            for (int c0 = eg * 32; c0 < cout; c0 += 64) {
                uint32_t v[32];
                float f[32];
                tmem_ld32(acc + c0, v);
                if (act == HRN_ACT_RELU && c0 + 32 <= cout) {
                    const f32x2_t a2 = f2_pack(a_w, a_w);             // packed fp32x2 bias add / attention scaling
#pragma unroll
                    for (int e = 0; e < 32; e += 4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(b3 + c0 + e);
                        float s0, s1, s2, s3;
                        f2_unpack(f2_add(f2_pack(__uint_as_float(v[e]), __uint_as_float(v[e + 1])), f2_pack(b4.x, b4.y)), s0, s1);
                        f2_unpack(f2_add(f2_pack(__uint_as_float(v[e + 2]), __uint_as_float(v[e + 3])), f2_pack(b4.z, b4.w)), s2, s3);
                        f2_unpack(f2_mul(f2_pack(fmaxf(s0, 0.f), fmaxf(s1, 0.f)), a2), f[e], f[e + 1]);
                        f2_unpack(f2_mul(f2_pack(fmaxf(s2, 0.f), fmaxf(s3, 0.f)), a2), f[e + 2], f[e + 3]);
                    }
                } else if (act == HRN_ACT_RELU) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) f[e] = (c0 + e < cout) ? fmaxf(__uint_as_float(v[e]) + b3[c0 + e], 0.f) * a_w : 0.f;
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e) f[e] = (c0 + e < cout) ? act_fn(__uint_as_float(v[e]) + b3[c0 + e], act) : 0.f;
                }
                if (A.Y) {
                    if (vec_ok && AW.use_tile && c0 + 32 <= cout) {
                        // transpose through shared memory: a store instruction covers 4 rows x 128 contiguous bytes
#pragma unroll
                        for (int e = 0; e < 32; e += 4)
                            *reinterpret_cast<float4*>(tile_w + lane * CW_TP + e) = make_float4(f[e], f[e + 1], f[e + 2], f[e + 3]);
                        __syncwarp();
                        const int piece = lane & 7;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int rl = (lane >> 3) + 4 * i;
                            const float4 t = *reinterpret_cast<const float4*>(tile_w + rl * CW_TP + piece * 4);
                            *reinterpret_cast<float4*>(A.Y + ((long long)tile * CTM + wq * 32 + rl) * A.ldy + c0 + piece * 4) = t;
                        }
                        __syncwarp();
                    } else {
                        float* yr = A.Y + r * A.ldy + c0;
                        if (vec_ok) {
#pragma unroll
                            for (int e = 0; e < 32; e += 4)
                                if (c0 + e < cout) *reinterpret_cast<float4*>(yr + e) = make_float4(f[e], f[e + 1], f[e + 2], f[e + 3]);
                        } else {
#pragma unroll
                            for (int e = 0; e < 32; ++e) if (c0 + e < cout) yr[e] = f[e];
                        }
                    }
                }
                if (A.mode != EPI_STORE && A.G) {
                    if (A.mode == EPI_GROUPMAX) {
                        if (act != HRN_ACT_RELU) {
#pragma unroll
                            for (int e = 0; e < 32; ++e) if (c0 + e >= cout) f[e] = -CUDART_INF_F;
                        }
                        seg_transpose_reduce<KSEG, true>(f, lane);
                    } else {
                        seg_transpose_reduce<KSEG, false>(f, lane);
                    }
#pragma unroll
                    for (int i = 0; i < PER; ++i) {
                        const int c = c0 + pos * PER + i;
                        if (c < cout) A.G[grp * cout + c] = f[i];
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s_fin[b]));
        }
    } else if (warp < CW_EPI_WARPS + CW_PROD_WARPS) {
        // ================= producers (same lane mapping as layer_ws_kernel, mlp_tc.cu) ==========================
        const hrn_rows_t& in = A.in;
        const int pw = warp - CW_EPI_WARPS;
        const int rsub = lane & 7, hf = (lane >> 3) & 1, cl = lane >> 4;
        int c0s[5], chs[4];
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
        const float* rp[2][4];
        float rsc[2][4];
        float sc[RAW][4];
        int ltile = blockIdx.x, li = 0;                      // copy cursor: tile, stage within the tile
        bool need_resolve = false;                           // the rows of tile `ltile` have to be resolved before its first copy
        int gs = 0; uint32_t gpar = 0;                       // ring G position
        uint8_t* raw0 = sRaw + (size_t)(pw * 4) * 512 + lane * 16;
        auto resolve = [&](int tile) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const unsigned ru = (unsigned)tile * CTM + pw * 16 + g * 8 + rsub;     // rows < 2^31 (checked on the host)
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    rp[g][s] = nullptr; rsc[g][s] = 1.f;
                    if (s < in.n_seg) {
                        const hrn_seg_t sg = in.seg[s];
                        const long long sr = sg.mode == HRN_SEG_DIRECT ? (long long)ru
                                           : sg.mode == HRN_SEG_BROADCAST ? (long long)(ru / (unsigned)in.group)
                                           : (long long)(ru / (unsigned)in.rows_per_batch) * in.src_rows_per_batch + in.gather_idx[ru];
                        rp[g][s] = sg.ptr + sr * sg.ld + sg.col0;
                        if (sg.row_scale) rsc[g][s] = __ldg(sg.row_scale + ru);
                    }
                }
            }
        };
        auto copy_piece = [&](const float* const (&pp)[4], const float (&ps)[4], int cg, uint8_t* dst, float& osc) {
            int sgi = 0;
#pragma unroll
            for (int q = 1; q < 4; ++q) if (cg >= c0s[q]) sgi = q;
            const float* p = pp[0]; int cs = c0s[0], chn = chs[0]; osc = ps[0];
#pragma unroll
            for (int q = 1; q < 4; ++q) if (sgi == q) { p = pp[q]; cs = c0s[q]; chn = chs[q]; osc = ps[q]; }
            const int ch0 = ((cg - cs) << 3) + 4 * hf;
            const bool ok = p != nullptr && cg < c0s[4] && chn - ch0 >= 4;
            const void* src = ok ? (const void*)(p + ch0) : (const void*)A.W;
            if (direct_mask >> sgi & 1)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(ok ? 16 : 0) : "memory");
            else
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(ok ? 16 : 0) : "memory");
        };
        auto issue = [&](int slot, float (&ss)[4]) {
            if (ltile < n_tiles) {
#pragma unroll
                for (int g = 0; g < 2; ++g)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        copy_piece(rp[g], rsc[g], li * 4 + 2 * j + cl, raw0 + (size_t)slot * CW_STAGE_BYTES + (g * 2 + j) * 512, ss[2 * g + j]);
                if (++li == n_st0) {
                    li = 0; ltile += gridDim.x;
                    need_resolve = ltile < n_tiles;      // done at the top of the stage loop: ONE inlined copy of resolve()
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto fill = [&](int slot, const float (&ss)[4]) {
            asm volatile("cp.async.wait_group %0;" ::"n"(RAW - 1) : "memory");
            float4 vv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) vv[e] = *reinterpret_cast<const float4*>(raw0 + (size_t)slot * CW_STAGE_BYTES + e * 512);
            mbar_wait_backoff(smem_u32(&s_gempty[gs]), gpar ^ 1);
            uint4* g_hi = reinterpret_cast<uint4*>(sG + (size_t)gs * CW_STAGE_BYTES);
            uint4* g_lo = g_hi + 4 * CTM;
#pragma unroll
            for (int g = 0; g < 2; ++g)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float4 t = vv[2 * g + j];
                    const float s_ = ss[2 * g + j];
                    const float x0 = t.x * s_, x1 = t.y * s_, x2 = t.z * s_, x3 = t.w * s_;
                    const int slot_a = (2 * j + cl) * CTM + pw * 16 + g * 8 + rsub;
                    if (PREC == 1) {           // each lane stores its own 8 bytes of the 16-byte core-matrix row
                        *reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(g_hi + slot_a) + 8 * hf) =
                            make_uint2(pack_f16x2(x0, x1), pack_f16x2(x2, x3));
                        continue;
                    }
                    uint32_t H0, H1, L0, L1;
                    split_pair(x0, x1, H0, L0);
                    split_pair(x2, x3, H1, L1);
                    const uint32_t r0 = __shfl_xor_sync(0xffffffffu, hf ? H0 : L0, 8);
                    const uint32_t r1 = __shfl_xor_sync(0xffffffffu, hf ? H1 : L1, 8);
                    if (hf == 0) g_hi[slot_a] = make_uint4(H0, H1, r0, r1);
                    else         g_lo[slot_a] = make_uint4(r0, r1, L0, L1);
                }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s_gfull[gs]));
            if (++gs == GS) { gs = 0; gpar ^= 1; }
        };
        // Z blocks of a tile (first-layer bias rows), one ring stage per 32-column block: warp pw copies the gathered rows
        // pw*16 .. +16 (4 rows x 128 B per instruction) and, for groups of 8 rows, the group rows 2pw, 2pw+1
        int zs = 0; uint32_t zpar = 0;
        auto z_blocks = [&](int tile) {
            const int gl = 2 * pw + (lane >> 3), chunk = lane & 7;
            const float* zsrc[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const unsigned ru = (unsigned)tile * CTM + pw * 16 + i * 4 + (lane >> 3);
                const long long zr = (long long)(ru / (unsigned)in.rows_per_batch) * in.src_rows_per_batch + in.gather_idx[ru];
                zsrc[i] = A.Zg + zr * A.ldz + chunk * 4;
            }
            const float* zbsrc = A.Zb ? A.Zb + ((long long)tile * (CTM / 8) + gl) * A.ldz + chunk * 4 : nullptr;
            const int nb = A.n[0] / 32;
            for (int j = 0; j < nb; ++j) {
                mbar_wait_backoff(smem_u32(&s_zempty[zs]), zpar ^ 1);
                const uint32_t base = smem_u32(sZ + (size_t)zs * CW_ZSTAGE_BYTES);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rl = pw * 16 + i * 4 + (lane >> 3);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(base + rl * 128 + ((chunk ^ (rl & 7)) << 4)), "l"(zsrc[i] + 32 * j) : "memory");
                }
                if (zbsrc && lane < 16)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(base + CW_ZG_BYTES + gl * 128 + (chunk << 4)), "l"(zbsrc + 32 * j) : "memory");
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&s_zfull[zs])) : "memory");
                if (++zs == ZS) { zs = 0; zpar ^= 1; }
            }
        };
        int my_tiles = 0;
        if ((int)blockIdx.x < n_tiles) my_tiles = (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1;
        const int total = my_tiles * n_st0;
        int filled = 0;                                      // stages delivered: a tile's Z blocks follow its last stage
        need_resolve = total > 0;
        // One copy of the stage code: unrolled over the RAW slots (prologue + main loop) it was 8 K of the kernel's 14 K
        // instructions, the warps of the four roles run different regions of it at the same time, and ncu showed 12 % of
        // the big chains' stall samples (27 % of the 50 us heads') on instruction fetches.  A slot's four row scales go
        // through local memory instead of registers (sc[d] with a run-time d).
#pragma unroll 1
        for (int d = -RAW; d < total; ++d) {
            const int slot = (d + RAW) % RAW;
            // the next tile's index / scale loads are issued here, in front of the wait inside fill()
            if (need_resolve) { resolve(ltile); need_resolve = false; }
            if (d >= 0) {
                fill(slot, sc[slot]);
                if (++filled % n_st0 == 0 && has_z) z_blocks((int)blockIdx.x + (filled / n_st0 - 1) * (int)gridDim.x);
            }
            issue(slot, sc[slot]);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else if (warp == CW_EPI_WARPS + CW_PROD_WARPS) {
        // ================= MMA issue ==============================================================================
        // One thread feeds the tensor core; its instruction stream per K=16 piece is kept minimal (no divisions, shared
        // memory descriptors = hoisted constant + 14-bit address field): at ~80 instructions per piece this thread, not
        // the tensor pipe, was the limiter of the whole kernel.
        if (lane == 0) {
            constexpr uint64_t DESC_FIXED = ((uint64_t)(128 >> 4) << 32) | (1ull << 46);          // SBO = 128 B, version bit
            const uint64_t a_desc0 = DESC_FIXED | ((uint64_t)((CTM * 16) >> 4) << 16);             // LBO = 2048 B
            const uint32_t g_a = smem_u32(sG) >> 4, h_a = smem_u32(sH) >> 4;                        // ring bases, 16-byte units
            constexpr uint32_t ST16 = CW_STAGE_BYTES >> 4, LO16 = (4 * CTM * 16) >> 4, P16 = (2 * CTM * 16) >> 4;
            const uint32_t slot16 = SLOT_BYTES >> 4;
            const uint32_t wfull0 = smem_u32(&s_wfull[0]), wempty0 = smem_u32(&s_wempty[0]);
            uint32_t ws = 0, wpar = 0;                        // weight ring position / parity
            int gs = 0, hs = 0; uint32_t gpar = 0, hpar = 0;  // operand ring positions
            int L = 0, ti = 0;
            uint32_t fin_pending = 0, finph = 0;              // per accumulator: last result not yet drained by the epilogue / phase
#ifdef HRN_CHAIN_PROF
            long long cp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            const long long cp_start = clock64();
#endif
            auto claim = [&](int b) {                         // before the first MMA of a layer into accumulator b
                if (fin_pending >> b & 1) {
                    CPROF_BEGIN();
                    mbar_wait(smem_u32(&s_fin[b]), (finph >> b) & 1);
                    CPROF_END(3);
                    finph ^= 1u << b; fin_pending &= ~(1u << b);
                }
            };
            // one K=16 piece: A = chunks [2p, 2p+2) of the operand stage whose hi plane starts at a16 (16-byte units)
            auto piece_mma = [&](uint32_t a16, uint64_t w_desc0, uint32_t wlo16, uint32_t idesc, uint32_t d, uint32_t accumulate) {
                {
                    CPROF_BEGIN();
                    mbar_wait(wfull0 + 8 * ws, wpar);
                    CPROF_END(2);
                }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t w16 = (ring_a >> 4) + ws * slot16;
                const uint64_t ah = a_desc0 | a16, al = a_desc0 | (a16 + LO16);
                const uint64_t wh = w_desc0 | w16, wl = w_desc0 | (w16 + wlo16);
                umma_bf16(d, ah, wh, idesc, accumulate);
                if (PREC == 3) {
                    umma_bf16(d, al, wh, idesc, 1u);
                    umma_bf16(d, ah, wl, idesc, 1u);
                }
                umma_commit(wempty0 + 8 * ws);
                if (++ws == (uint32_t)RING) { ws = 0; wpar ^= 1; }
            };
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
                // ---- layer 1: input stages as the producers deliver them ----
                {
                    const int b = L & BM;
                    const uint32_t d = tmem + b * ACC;
                    claim(b);
                    if (nl == 1) fin_pending |= 1u << b;
                    const uint32_t N = (uint32_t)A.n[0], idesc = umma_idesc_m128<PREC>(A.n[0]);
                    const uint64_t w_desc0 = DESC_FIXED | ((uint64_t)N << 16);                      // LBO = N * 16 B
                    const uint32_t wlo16 = 2 * N;                                                   // lo plane: + 2 * N * 16 B
                    int left = (int)pieces[0];
                    for (int s = 0; s < n_st0; ++s, left -= 2) {
                        {
                            CPROF_BEGIN();
                            mbar_wait(smem_u32(&s_gfull[gs]), gpar);
                            CPROF_END(0);
                        }
                        const uint32_t a16 = g_a + gs * ST16;
                        piece_mma(a16, w_desc0, wlo16, idesc, d, s > 0 ? 1u : 0u);
                        if (left > 1) piece_mma(a16 + P16, w_desc0, wlo16, idesc, d, 1u);
                        umma_commit(smem_u32(&s_gempty[gs]));
                        if (++gs == GS) { gs = 0; gpar ^= 1; }
                    }
                    umma_commit(smem_u32(&s_accf[b]));
                    ++L;
                }
                // ---- later layers: 32-column blocks as the epilogue warps convert them ----
                for (int l = 1; l < nl; ++l, ++L) {
                    const int b = L & BM;
                    const uint32_t d = tmem + b * ACC;
                    claim(b);
                    if (l == nl - 1) fin_pending |= 1u << b;
                    const uint32_t N = (uint32_t)A.n[l], idesc = umma_idesc_m128<PREC>(A.n[l]);
                    const uint64_t w_desc0 = DESC_FIXED | ((uint64_t)N << 16);
                    const uint32_t wlo16 = 2 * N;
                    int left = A.n[l - 1] / 16;
                    for (int j = 0; left > 0; ++j, left -= 2) {
                        {
                            CPROF_BEGIN();
                            mbar_wait(smem_u32(&s_hfull[hs]), hpar);
                            CPROF_END(1);
                        }
                        const uint32_t a16 = h_a + hs * ST16;
                        piece_mma(a16, w_desc0, wlo16, idesc, d, j > 0 ? 1u : 0u);
                        if (left > 1) piece_mma(a16 + P16, w_desc0, wlo16, idesc, d, 1u);
                        umma_commit(smem_u32(&s_hempty[hs]));
                        if (++hs == HS) { hs = 0; hpar ^= 1; }
                    }
                    umma_commit(smem_u32(&s_accf[b]));
                }
            }
#ifdef HRN_CHAIN_PROF
            if (blockIdx.x == 0) {
                cp[4] = clock64() - cp_start;
                for (int i = 0; i < 8; ++i) g_chain_prof[i] = cp[i];
            }
#endif
        }
    } else {
        // ================= weight stream ==========================================================================
        if (lane == 0) {
            uint32_t ws = 0, wpar = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const uint8_t* src = A.W;
                for (int l = 0; l < nl; ++l) {
                    const uint32_t bytes = (uint32_t)A.n[l] * (PREC == 1 ? 32u : 64u);     // hi (+ lo) plane of a K=16 piece
                    for (uint32_t p = 0; p < pieces[l]; ++p) {
                        mbar_wait_backoff(smem_u32(&s_wempty[ws]), wpar ^ 1);
                        mbar_expect_tx(smem_u32(&s_wfull[ws]), bytes);
                        bulk_g2s(ring_a + ws * SLOT_BYTES, src, bytes, smem_u32(&s_wfull[ws]));
                        src += bytes;
                        if (++ws == (uint32_t)RING) { ws = 0; wpar ^= 1; }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

inline int grid_ws_of(int n_tiles) { return n_tiles < 148 ? n_tiles : 148; }

template <int KSEG, int RAW, int PREC>
cudaError_t launch_chain_one(const ChainWsArgs& AW, int grid, int smem, int budget, cudaStream_t st) {
    static hrn_once_per_device attr;
    if (attr.need()) {
        cudaError_t e = cudaFuncSetAttribute(chain_ws_kernel<KSEG, RAW, PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, budget);
        if (e != cudaSuccess) return e;
    }
    chain_ws_kernel<KSEG, RAW, PREC><<<grid, CW_THREADS, smem, st>>>(AW);
    return cudaSuccess;
}

template <int PREC>
cudaError_t launch_chain_prec(const ChainWsArgs& AW, int kseg, bool narrow, int grid, int smem, int budget, cudaStream_t st) {
    if (narrow) {
        if (kseg == 8) return launch_chain_one<8, 4, PREC>(AW, grid, smem, budget, st);
        if (kseg == 16) return launch_chain_one<16, 4, PREC>(AW, grid, smem, budget, st);
        return launch_chain_one<32, 4, PREC>(AW, grid, smem, budget, st);
    }
    if (kseg == 8) return launch_chain_one<8, 2, PREC>(AW, grid, smem, budget, st);
    if (kseg == 16) return launch_chain_one<16, 2, PREC>(AW, grid, smem, budget, st);
    return launch_chain_one<32, 2, PREC>(AW, grid, smem, budget, st);
}

cudaError_t launch_chain(const ChainWsArgs& AW, int kseg, bool narrow, int prec, int grid, int smem, int budget, cudaStream_t st) {
    return prec == 1 ? launch_chain_prec<1>(AW, kseg, narrow, grid, smem, budget, st)
                     : launch_chain_prec<3>(AW, kseg, narrow, grid, smem, budget, st);
}

}  // namespace

// Two or three fused layers on a virtual rows matrix.  W: packed K=16 pieces of the layers in execution order
// (engine_tc.pack_chain; prec = 3: bf16 hi + lo planes, prec = 1: one fp16 plane), bias = b1|b2|b3 (last: `cout` entries),
// issued widths n1,n2,(n3) multiples of 16 and <= 256
// (the last one = cout padded to 16), hidden activations ReLU, last activation `act`.  mode / outputs as in the file
// header; `kseg` = rows per group (8, 16 or 32; ignored for mode 0); rows must be a multiple of 128.
HRN_API int hrn_chain_tc(const hrn_rows_t* in, const void* W, const float* bias, int nl, int n1, int n2, int n3, int cout,
                         int act, int chunks0, int mode, int kseg, float* Y, int ldy, float* G, float* a, long long rows,
                         int prec, const float* Zb, const float* Zg, int ldz, void* stream) {
    if (!in || !W || !bias || rows < 0 || in->n_seg < 1 || in->n_seg > 4 || (nl != 2 && nl != 3)) return HRN_ERR_BAD_ARG;
    if (prec != 1 && prec != 3) return HRN_ERR_BAD_ARG;
    const int nn[3] = {n1, n2, nl == 3 ? n3 : 16};
    for (int l = 0; l < 3; ++l) if (nn[l] % 16 || nn[l] > 256 || nn[l] < 16) return HRN_ERR_UNSUPPORTED;
    const int nlast = nl == 3 ? n3 : n2;
    if (cout <= 0 || cout > nlast) return HRN_ERR_BAD_ARG;
    if (kseg != 8 && kseg != 16 && kseg != 32) return HRN_ERR_UNSUPPORTED;
    if (rows % CTM != 0 || (chunks0 & 1)) return HRN_ERR_UNSUPPORTED;
    int chunks = 0;
    for (int s = 0; s < in->n_seg; ++s) {
        const hrn_seg_t& g = in->seg[s];
        if (!g.ptr || g.channels <= 0) return HRN_ERR_BAD_ARG;
        if ((g.channels & 3) || (g.ld & 3) || (g.col0 & 3) || ((uintptr_t)g.ptr & 15)) return HRN_ERR_UNSUPPORTED;
        if (g.mode == HRN_SEG_GATHER && !in->gather_idx) return HRN_ERR_BAD_ARG;
        if (g.mode == HRN_SEG_BROADCAST && in->group <= 0) return HRN_ERR_BAD_ARG;
        chunks += (g.channels + 7) / 8;
    }
    if (chunks0 != ((chunks + 1) & ~1)) return HRN_ERR_BAD_ARG;
    if (mode == EPI_STORE && !Y) return HRN_ERR_BAD_ARG;
    if (mode == EPI_ATTN && act != HRN_ACT_RELU) return HRN_ERR_UNSUPPORTED;
    if (Y && (cout & 3) == 0 && ((ldy & 3) || ((uintptr_t)Y & 15))) return HRN_ERR_UNSUPPORTED;
    if (Zb && !Zg) return HRN_ERR_BAD_ARG;
    if (Zg && (!in->gather_idx || in->rows_per_batch <= 0 || ldz < n1 || (ldz & 3) || ((uintptr_t)Zg & 15) || (n1 & 31))) return HRN_ERR_BAD_ARG;
    if (Zb && (kseg != 8 || in->group != 8 || ((uintptr_t)Zb & 15))) return HRN_ERR_UNSUPPORTED;
    if (rows == 0) return HRN_OK;
    ChainArgs A;
    A.in = *in; A.W = (const uint8_t*)W; A.bias = bias; A.Y = Y; A.G = G; A.a = a; A.rows = rows; A.ldy = ldy;
    A.n[0] = n1; A.n[1] = n2; A.n[2] = nl == 3 ? n3 : 16; A.nl = nl; A.cout = cout; A.act = act;
    A.chunks0 = chunks0; A.mode = mode; A.kseg = kseg;
    A.Zb = Zb; A.Zg = Zg; A.ldz = ldz;
    int maxn = n1 > n2 ? n1 : n2;
    if (nl == 3 && n3 > maxn) maxn = n3;
    A.slot_bytes = maxn * (prec == 1 ? 32 : 64);
    cudaStream_t st = (cudaStream_t)stream;
    if (rows >= 0x7fffffffLL) return HRN_ERR_UNSUPPORTED;        // 32-bit row arithmetic in the producers
    {
        // shared memory: ring G + ring H (16 KB stages) | weight ring | raw fp32 ring | biases | optional store tile
        const int budget = 227 * 1024 - 1024;
        const bool narrow = maxn <= 128;
        const int raw = narrow ? 4 : 2;
        const int tile_b = Y ? CW_TILE_BYTES : 0;
        int gs = 3, hs = 3, fixed = 0, ring = 0, use_tile = Y ? 1 : 0;
        const int zs = Zg ? 2 : 0;
        const bool one_stage = Zg && chunks0 <= 4;                 // one input stage per tile: no need to run stages ahead
        for (int attempt = 0; attempt < 4; ++attempt) {            // keep >= 4 weight slots: shrink the operand rings, then drop the tile
            gs = one_stage ? 1 : (attempt >= 1 ? 2 : 3); hs = attempt >= 2 ? 2 : 3;
            use_tile = (Y && attempt < 3) ? 1 : 0;
            fixed = (gs + hs + raw) * CW_STAGE_BYTES + 3 * 256 * 4 + 2 * CTM * 4 + zs * CW_ZSTAGE_BYTES;
            ring = (budget - fixed - (use_tile ? tile_b : 0)) / A.slot_bytes;
            if (ring >= 4) break;
        }
        if (ring > CW_RING_MAX) ring = CW_RING_MAX;
        A.ring = ring;
        ChainWsArgs AW;
        AW.c = A;
        AW.n_tiles = (int)(rows / CTM);
        AW.use_tile = use_tile;
        AW.gs = gs; AW.hs = hs;
        AW.zs = zs;
        AW.acc_stride = narrow ? 128 : 256;
        AW.nbuf = 512 / AW.acc_stride;
        const int smem_ws = fixed + ring * A.slot_bytes + (use_tile ? CW_TILE_BYTES : 0);
        HRN_CUDA((launch_chain(AW, kseg, narrow, prec, grid_ws_of(AW.n_tiles), smem_ws, budget, st)));
        HRN_LAUNCH_CHECK();
        return HRN_OK;
    }
}

#ifdef HRN_CHAIN_PROF
HRN_API int hrn_chain_prof(long long* host8) { return (int)cudaMemcpyFromSymbol(host8, g_chain_prof, 8 * sizeof(long long)); }
#endif
