// Three-layer shared-MLP chain for layers up to 512 wide, on a CLUSTER OF TWO CTAs, with the activations kept on chip.
//
//     X (virtual rows, rows.cuh) -> relu(W1 . + b1) -> relu(W2 . + b2) -> relu(W3 . + b3)
//       -> a = softmax_k(max_c Y),  AF[g,:] = sum_k a Y
//
// Replaces CoarseReg's  convs_1  (528 -> 512 -> 512 -> 512, reference models/HRegNet/layers.py:364-375) and the attention
// tail that consumes it (layers.py:384-390).  A 128-row tile of 512-wide hidden activations is 256 KB as bf16 hi/lo
// operands and its accumulator is all of one SM's tensor memory, so the single-CTA chain kernel (chain_tc.cu) stops at 256
// columns and this stage ran layer by layer: three launches that wrote and re-read 134 MB each (0.57 GB per 32-pair
// step) and gathered the 528-channel input once per 256-column block.
//
// Here a cluster of two CTAs owns a tile: BOTH see the same 128 rows, EACH computes one half of every layer's output
// columns (N = 256 per MMA, two 256-column accumulators in tensor memory used alternately by consecutive layers).  A
// hidden layer's result leaves the accumulator 32 columns at a time (TMEM -> bias / ReLU -> bf16 hi/lo), lands in the
// producing CTA's ring HL and is forwarded to the peer's ring HR by one bulk shared->shared::cluster copy
// (cp.async.bulk + complete_tx on the PEER's mbarrier); both CTAs consume the blocks in the same order
// (rank 0 block j, rank 1 block j, ...), so layer l+1's MMAs run in both CTAs while layer l is still being drained.
// A ring slot is reused when BOTH CTAs have consumed the block pair: their MMA threads signal it with ONE multicast
// tcgen05.commit on the slot's mbarrier in both CTAs (count 2).  Nothing per-row reaches HBM; the row maxima of the two
// column halves meet through distributed shared memory.
//
// Roles per CTA (608 threads):
//   warps 0-7   epilogue   (warp & 3 = TMEM lane quadrant, warp >> 2 = group; the groups take alternate 32-column blocks)
//   warps 8-15  producers  coalesced cp.async gather of the virtual rows -> raw fp32 ring -> bf16 hi/lo ring G (each CTA
//                          gathers the whole input: both need all K columns of the first layer)
//   warp 16     MMA        one thread issues every K=16 piece as its operand block and its weights land
//   warp 17     weights    streams this CTA's half of the packed K=16 weight pieces through a cp.async.bulk ring
//   warp 18     sender     forwards every finished block of ring HL to the peer's ring HR
// bf16x3 products (or one fp16 plane: PREC 1), fp32 accumulation -- same numerics as chain_tc.cu.
#include "common.cuh"
#include "tc_common.cuh"
#include <math_constants.h>
#include <stdio.h>
#include <stdlib.h>

namespace {

constexpr int WTM = 128;                                               // rows per tile (UMMA M)
constexpr int WW_EPI_WARPS = 8, WW_PROD_WARPS = 8;
constexpr int WW_THREADS = (WW_EPI_WARPS + WW_PROD_WARPS + 3) * 32;    // 608
constexpr int WW_STAGE_BYTES = 4 * WTM * 32;                           // 128 rows x 32 K: fp32, or bf16 hi + lo = 16 KB
constexpr int WW_RING_MAX = 8, WW_GS_MAX = 4, WW_HS_MAX = 4;
constexpr int WW_ACC = 256;                                            // accumulator stride (TMEM columns)
constexpr int WW_ZS_MAX = 4;
constexpr int WW_ZG_BYTES = WTM * 32 * 4;                              // gathered rows of a 32-column block: 128 x 128 B
constexpr int WW_ZSTAGE_BYTES = WW_ZG_BYTES + (WTM / 8) * 32 * 4;      // + the 16 group rows = 18 KB

struct WideArgs {
    hrn_rows_t in;
    const uint8_t* W;            // per rank: pieces of layer 1 | 2 | 3 (this rank's output columns), execution order
    long long w_rank_bytes;
    const float* bias;           // per rank: b1 | b2 | b3 halves
    float* G;                    // groups x (2 nh[2]): attentive feature
    float* a;                    // rows: attention weights
    int n_tiles;
    int nh[3];                   // columns per CTA of the three layers (multiples of 32, <= 256)
    int chunks0;                 // 8-wide K chunks of the virtual input (even)
    int slot_bytes, ring, gs, hs;
    // optional fp32 rows ADDED to the first layer's pre-activations (the part of that layer which is the same for the
    // rows of a group / which only depends on the gathered source row, applied once per point by the caller):
    const float* Zb;             // [rows / group, ldz]: row r takes Zb[r / group]
    const float* Zg;             // [src rows, ldz]:     row r takes Zg[b * src_rows_per_batch + gather_idx[r]]
    int ldz, zs;                 // leading dimension (floats), Z ring stages
    // head mode (per-keypoint confidence heads, layers.py:391-394): no attention tail, Yh[r] = act(column 0 of the last
    // layer + its bias); the last layer is zero-padded to 64 columns by the caller
    float* Yh;
    int head_act;
#ifdef HRN_WIDE_DEBUG
    float* dbg;                  // [2 layers][rows][512] pre-activations of the hidden layers as the epilogue sees them
#endif
};
#ifdef HRN_WIDE_DEBUG
float* g_wide_dbg = nullptr;
#endif
// -DHRN_WIDE_PROF: cycles the MMA thread of CTA 0 spends waiting for [0] input stages, [1] own blocks, [2] the peer's
// blocks, [3] weights, [4] a free accumulator, [5] total; read back with hrn_chain_wide_prof (tools/wide_probe.py)
#ifdef HRN_WIDE_PROF
__device__ long long g_wide_prof[8];
#define WPROF_BEGIN() const long long wp_t0 = clock64()
#define WPROF_END(i) wp[i] += clock64() - wp_t0
#else
#define WPROF_BEGIN() do { } while (0)
#define WPROF_END(i) do { } while (0)
#endif

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t map_peer(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one commit, arrivals on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}

template <int KSEG>
__device__ __forceinline__ void wide_seg_sum(float (&v)[32], int lane) {
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
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        }
        cnt = half;
    }
}

template <int KSEG, int RAW, int PREC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(WW_THREADS, 1) chain_wide_kernel(const WideArgs A) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_wfull[WW_RING_MAX], s_wempty[WW_RING_MAX], s_gfull[WW_GS_MAX], s_gempty[WW_GS_MAX],
        s_hlfull[WW_HS_MAX], s_hrfull[WW_HS_MAX], s_hfree[WW_HS_MAX], s_accf[2], s_fin[2], s_xbar, s_zfull[WW_ZS_MAX],
        s_zempty[WW_ZS_MAX];
    __shared__ uint32_t s_tmem;

    const uint32_t rank = cluster_rank(), peer = rank ^ 1u;
    const int RING = A.ring, GS = A.gs, HS = A.hs;
    const uint32_t SLOT_BYTES = (uint32_t)A.slot_bytes;
    constexpr uint32_t BLOCK_BYTES = PREC == 1 ? WW_STAGE_BYTES / 2 : WW_STAGE_BYTES;   // hi (+ lo) plane of a 32-column block
    uint8_t* sG = smem;
    uint8_t* sHL = sG + GS * WW_STAGE_BYTES;
    uint8_t* sHR = sHL + HS * WW_STAGE_BYTES;
    uint8_t* sRing = sHR + HS * WW_STAGE_BYTES;
    uint8_t* sRaw = sRing + RING * SLOT_BYTES;
    float* sBias = reinterpret_cast<float*>(sRaw + RAW * WW_STAGE_BYTES);
    float* sX = sBias + 3 * 256;                            // [2][128] row maxima of the two epilogue groups
    float* sXr = sX + 2 * WTM;                              // [2][128] the peer's row maxima (by tile parity)
    uint8_t* sZ = reinterpret_cast<uint8_t*>(sXr + 2 * WTM);  // Z ring: [zs][128 gathered rows x 128 B (16-byte chunks XOR-swizzled by row) | 16 group rows x 128 B]
    const bool has_z = A.Zg != nullptr;
    const int ZS = A.zs;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tiles = A.n_tiles;
    const int tile0 = (int)(blockIdx.x >> 1), tstride = (int)(gridDim.x >> 1);
    const int n_st0 = (A.chunks0 + 3) >> 2;                 // 32-wide input stages per tile

    if (tid == 0) {
        for (int i = 0; i < WW_RING_MAX; ++i) { mbar_init(smem_u32(&s_wfull[i]), 1); mbar_init(smem_u32(&s_wempty[i]), 1); }
        for (int i = 0; i < WW_GS_MAX; ++i) { mbar_init(smem_u32(&s_gfull[i]), WW_PROD_WARPS); mbar_init(smem_u32(&s_gempty[i]), 1); }
        for (int i = 0; i < WW_HS_MAX; ++i) {
            mbar_init(smem_u32(&s_hlfull[i]), 4);           // the four quadrant warps of the producing epilogue group
            mbar_init(smem_u32(&s_hrfull[i]), 1);           // the peer's sender: arrive.expect_tx + the copy's complete_tx
            mbar_init(smem_u32(&s_hfree[i]), 2);            // both CTAs' MMA threads (multicast commit)
        }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&s_accf[i]), 1); mbar_init(smem_u32(&s_fin[i]), WW_EPI_WARPS); }
        mbar_init(smem_u32(&s_xbar), WTM);                  // one remote arrival per row
        for (int i = 0; i < WW_ZS_MAX; ++i) {
            mbar_init(smem_u32(&s_zfull[i]), WW_PROD_WARPS * 32);   // every producer thread: cp.async.mbarrier.arrive.noinc
            mbar_init(smem_u32(&s_zempty[i]), 4);                   // the four quadrant warps that consumed the block
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {
        const int nbias = A.nh[0] + A.nh[1] + A.nh[2];
        const float* bsrc = A.bias + (size_t)rank * nbias;
        for (int i = tid; i < nbias; i += WW_THREADS) sBias[i] = __ldg(bsrc + i);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                                     // the peer's barriers exist before anything is sent to them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t ring_a = smem_u32(sRing);

    if (warp < WW_EPI_WARPS) {
        // ================= epilogue warps ======================================================================
        const int eg = warp >> 2, wq = warp & 3;
        const int rt = wq * 32 + lane;                       // row inside the tile = TMEM lane
        const uint32_t lane_base = ((uint32_t)(wq * 32) << 16);
        uint32_t accph = 0, xph = 0;
        int L = 0;
        int hs = 0; uint32_t hpar = 0; int hq = 0;
        int zs = 0; uint32_t zpar = 0;
        int ti = 0;
        const uint32_t xbar_peer = map_peer(smem_u32(&s_xbar), peer);
        for (int tile = tile0; tile < n_tiles; tile += tstride, ++ti) {
            const long long r = (long long)tile * WTM + rt;
            for (int l = 0; l < 2; ++l, ++L) {
                const int b = L & 1, N = A.nh[l];
                const float* bb = sBias + (l == 0 ? 0 : A.nh[0]);
                mbar_wait(smem_u32(&s_accf[b]), (accph >> b) & 1); accph ^= 1u << b;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int c0 = 0; c0 < N; c0 += 32, ++hq) {
                    if ((hq & 1) == eg) {
                        uint32_t v[32];
                        tmem_ld32(tmem + lane_base + b * WW_ACC + c0, v);
                        const bool zblk = has_z && l == 0;
                        const float* zg = reinterpret_cast<const float*>(sZ + (size_t)zs * WW_ZSTAGE_BYTES) + rt * 32;
                        const float* zb = reinterpret_cast<const float*>(sZ + (size_t)zs * WW_ZSTAGE_BYTES + WW_ZG_BYTES) + (rt >> 3) * 32;
                        if (zblk) mbar_wait(smem_u32(&s_zfull[zs]), zpar);           // this block's Z rows have landed
                        mbar_wait(smem_u32(&s_hfree[hs]), hpar ^ 1);          // both CTAs are done with the slot's last pair
                        uint4* h_hi = reinterpret_cast<uint4*>(sHL + (size_t)hs * WW_STAGE_BYTES);
                        uint4* h_lo = h_hi + 4 * WTM;
#pragma unroll
                        for (int ch = 0; ch < 4; ++ch) {
                            const float4 b0 = *reinterpret_cast<const float4*>(bb + c0 + ch * 8);
                            const float4 b1 = *reinterpret_cast<const float4*>(bb + c0 + ch * 8 + 4);
                            float s[8];
                            f2_unpack(f2_add(f2_pack(__uint_as_float(v[ch * 8 + 0]), __uint_as_float(v[ch * 8 + 1])), f2_pack(b0.x, b0.y)), s[0], s[1]);
                            f2_unpack(f2_add(f2_pack(__uint_as_float(v[ch * 8 + 2]), __uint_as_float(v[ch * 8 + 3])), f2_pack(b0.z, b0.w)), s[2], s[3]);
                            f2_unpack(f2_add(f2_pack(__uint_as_float(v[ch * 8 + 4]), __uint_as_float(v[ch * 8 + 5])), f2_pack(b1.x, b1.y)), s[4], s[5]);
                            f2_unpack(f2_add(f2_pack(__uint_as_float(v[ch * 8 + 6]), __uint_as_float(v[ch * 8 + 7])), f2_pack(b1.z, b1.w)), s[6], s[7]);
                            if (zblk) {     // + group row + gathered row (fp32)
                                const float4 g0 = *reinterpret_cast<const float4*>(zg + (((2 * ch) ^ (rt & 7)) << 2));
                                const float4 g1 = *reinterpret_cast<const float4*>(zg + (((2 * ch + 1) ^ (rt & 7)) << 2));
                                const float4 q0 = *reinterpret_cast<const float4*>(zb + ch * 8);
                                const float4 q1 = *reinterpret_cast<const float4*>(zb + ch * 8 + 4);
                                f2_unpack(f2_add(f2_add(f2_pack(s[0], s[1]), f2_pack(q0.x, q0.y)), f2_pack(g0.x, g0.y)), s[0], s[1]);
                                f2_unpack(f2_add(f2_add(f2_pack(s[2], s[3]), f2_pack(q0.z, q0.w)), f2_pack(g0.z, g0.w)), s[2], s[3]);
                                f2_unpack(f2_add(f2_add(f2_pack(s[4], s[5]), f2_pack(q1.x, q1.y)), f2_pack(g1.x, g1.y)), s[4], s[5]);
                                f2_unpack(f2_add(f2_add(f2_pack(s[6], s[7]), f2_pack(q1.z, q1.w)), f2_pack(g1.z, g1.w)), s[6], s[7]);
                            }
#ifdef HRN_WIDE_DEBUG
                            if (A.dbg)
                                for (int e = 0; e < 8; ++e)
                                    A.dbg[((size_t)l * A.n_tiles * WTM + r) * 512 + rank * N + c0 + ch * 8 + e] = s[e];
#endif
                            if (PREC == 1) {
                                h_hi[ch * WTM + rt] = make_uint4(pack_f16x2_relu(s[0], s[1]), pack_f16x2_relu(s[2], s[3]),
                                                                 pack_f16x2_relu(s[4], s[5]), pack_f16x2_relu(s[6], s[7]));
                            } else {
                                const float x[8] = {fmaxf(s[0], 0.f), fmaxf(s[1], 0.f), fmaxf(s[2], 0.f), fmaxf(s[3], 0.f),
                                                    fmaxf(s[4], 0.f), fmaxf(s[5], 0.f), fmaxf(s[6], 0.f), fmaxf(s[7], 0.f)};
                                split_store8(x, h_hi + ch * WTM + rt, h_lo + ch * WTM + rt);
                            }
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) {
                            mbar_arrive(smem_u32(&s_hlfull[hs]));
                            if (zblk) mbar_arrive(smem_u32(&s_zempty[zs]));
                        }
                    }
                    if (++hs == HS) { hs = 0; hpar ^= 1; }
                    if (has_z && l == 0 && ++zs == ZS) { zs = 0; zpar ^= 1; }
                }
            }
            // ---- last layer: attention over this CTA's column half, the row maxima of the halves meet through DSMEM ----
            const int b = L & 1;
            ++L;
            mbar_wait(smem_u32(&s_accf[b]), (accph >> b) & 1); accph ^= 1u << b;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc = tmem + lane_base + b * WW_ACC;
            const float* b3 = sBias + A.nh[0] + A.nh[1];
            const int cout = A.nh[2];
            if (A.Yh) {                                       // head mode: one output per row, from rank 0's first column
                if (rank == 0 && eg == 0) {
                    uint32_t v[32];
                    tmem_ld32(acc, v);
                    A.Yh[r] = act_fn(__uint_as_float(v[0]) + b3[0], A.head_act);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&s_fin[b]));
                continue;
            }
            const int pos = lane % KSEG;
            const long long grp = r / KSEG;
            constexpr int PER = 32 / KSEG;
            float x1 = 0.f;                                   // post-ReLU values are >= 0
            for (int c0 = eg * 32; c0 < cout; c0 += 64) {
                uint32_t v[32];
                tmem_ld32(acc + c0, v);
#pragma unroll
                for (int e = 0; e < 32; e += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(b3 + c0 + e);
                    float s0, s1, s2, s3;
                    f2_unpack(f2_add(f2_pack(__uint_as_float(v[e]), __uint_as_float(v[e + 1])), f2_pack(b4.x, b4.y)), s0, s1);
                    f2_unpack(f2_add(f2_pack(__uint_as_float(v[e + 2]), __uint_as_float(v[e + 3])), f2_pack(b4.z, b4.w)), s2, s3);
                    x1 = fmaxf(x1, fmaxf(fmaxf(s0, s1), fmaxf(s2, s3)));
                }
            }
            sX[eg * WTM + rt] = x1;
            asm volatile("bar.sync 1, 256;" ::: "memory");                 // the 8 epilogue warps
            x1 = fmaxf(x1, sX[(eg ^ 1) * WTM + rt]);
            float* xr = sXr + (ti & 1) * WTM;
            if (eg == 0) {
                const uint32_t dst = map_peer(smem_u32(xr + rt), peer);
                asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(dst), "f"(x1) : "memory");
                asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(xbar_peer) : "memory");
            }
            mbar_wait_cluster(smem_u32(&s_xbar), xph);
            xph ^= 1;
            x1 = fmaxf(x1, xr[rt]);
            float gm = x1;
#pragma unroll
            for (int o = KSEG / 2; o > 0; o >>= 1) gm = fmaxf(gm, __shfl_xor_sync(0xffffffffu, gm, o));
            const float ex = expf(x1 - gm);
            float sm = ex;
#pragma unroll
            for (int o = KSEG / 2; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
            const float a_w = ex / sm;
            if (A.a && rank == 0 && eg == 0) A.a[r] = a_w;
            const int ctot = 2 * cout;
            for (int c0 = eg * 32; c0 < cout; c0 += 64) {
                uint32_t v[32];
                float f[32];
                tmem_ld32(acc + c0, v);
                const f32x2_t a2 = f2_pack(a_w, a_w);
#pragma unroll
                for (int e = 0; e < 32; e += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(b3 + c0 + e);
                    float s0, s1, s2, s3;
                    f2_unpack(f2_add(f2_pack(__uint_as_float(v[e]), __uint_as_float(v[e + 1])), f2_pack(b4.x, b4.y)), s0, s1);
                    f2_unpack(f2_add(f2_pack(__uint_as_float(v[e + 2]), __uint_as_float(v[e + 3])), f2_pack(b4.z, b4.w)), s2, s3);
                    f2_unpack(f2_mul(f2_pack(fmaxf(s0, 0.f), fmaxf(s1, 0.f)), a2), f[e], f[e + 1]);
                    f2_unpack(f2_mul(f2_pack(fmaxf(s2, 0.f), fmaxf(s3, 0.f)), a2), f[e + 2], f[e + 3]);
                }
                wide_seg_sum<KSEG>(f, lane);
#pragma unroll
                for (int i = 0; i < PER; ++i) A.G[grp * ctot + rank * cout + c0 + pos * PER + i] = f[i];
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s_fin[b]));
        }
    } else if (warp < WW_EPI_WARPS + WW_PROD_WARPS) {
        // ================= producers (same lane mapping as chain_ws_kernel, chain_tc.cu) ==========================
        const hrn_rows_t& in = A.in;
        const int pw = warp - WW_EPI_WARPS;
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
        int ltile = tile0, li = 0;                           // copy cursor: tile, stage within the tile
        bool need_resolve = false;                           // the rows of tile `ltile` have to be resolved before its first copy
        int gs = 0; uint32_t gpar = 0;                       // ring G position
        uint8_t* raw0 = sRaw + (size_t)(pw * 4) * 512 + lane * 16;
        auto resolve = [&](int tile) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const unsigned ru = (unsigned)tile * WTM + pw * 16 + g * 8 + rsub;     // rows < 2^31 (checked on the host)
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
                        copy_piece(rp[g], rsc[g], li * 4 + 2 * j + cl, raw0 + (size_t)slot * WW_STAGE_BYTES + (g * 2 + j) * 512, ss[2 * g + j]);
                if (++li == n_st0) {
                    li = 0; ltile += tstride;
                    need_resolve = ltile < n_tiles;      // done at the top of the stage loop: ONE inlined copy of resolve()
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto fill = [&](int slot, const float (&ss)[4]) {
            asm volatile("cp.async.wait_group %0;" ::"n"(RAW - 1) : "memory");
            float4 vv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) vv[e] = *reinterpret_cast<const float4*>(raw0 + (size_t)slot * WW_STAGE_BYTES + e * 512);
            mbar_wait_backoff(smem_u32(&s_gempty[gs]), gpar ^ 1);
            uint4* g_hi = reinterpret_cast<uint4*>(sG + (size_t)gs * WW_STAGE_BYTES);
            uint4* g_lo = g_hi + 4 * WTM;
#pragma unroll
            for (int g = 0; g < 2; ++g)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float4 t = vv[2 * g + j];
                    const float s_ = ss[2 * g + j];
                    const float x0 = t.x * s_, x1 = t.y * s_, x2 = t.z * s_, x3 = t.w * s_;
                    const int slot_a = (2 * j + cl) * WTM + pw * 16 + g * 8 + rsub;
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
        // Z blocks of a tile (first-layer bias rows of this CTA's column half), one ring stage per 32-column block:
        // warp pw copies the gathered rows pw*16 .. +16 (4 rows x 128 B per instruction) and the group rows 2pw, 2pw+1
        int zs = 0; uint32_t zpar = 0;
        auto z_blocks = [&](int tile) {
            const int gl = 2 * pw + (lane >> 3), chunk = lane & 7;
            const float* zsrc[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const unsigned ru = (unsigned)tile * WTM + pw * 16 + i * 4 + (lane >> 3);
                const long long zr = (long long)(ru / (unsigned)in.rows_per_batch) * in.src_rows_per_batch + in.gather_idx[ru];
                zsrc[i] = A.Zg + zr * A.ldz + rank * A.nh[0] + chunk * 4;
            }
            const float* zbsrc = A.Zb + ((long long)tile * (WTM / 8) + gl) * A.ldz + rank * A.nh[0] + chunk * 4;
            const int nb = A.nh[0] / 32;
            for (int j = 0; j < nb; ++j) {
                mbar_wait_backoff(smem_u32(&s_zempty[zs]), zpar ^ 1);
                const uint32_t base = smem_u32(sZ + (size_t)zs * WW_ZSTAGE_BYTES);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rl = pw * 16 + i * 4 + (lane >> 3);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(base + rl * 128 + ((chunk ^ (rl & 7)) << 4)), "l"(zsrc[i] + 32 * j) : "memory");
                }
                if (lane < 16)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(base + WW_ZG_BYTES + gl * 128 + (chunk << 4)), "l"(zbsrc + 32 * j) : "memory");
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&s_zfull[zs])) : "memory");
                if (++zs == ZS) { zs = 0; zpar ^= 1; }
            }
        };
        int my_tiles = 0;
        if (tile0 < n_tiles) my_tiles = (n_tiles - 1 - tile0) / tstride + 1;
        const int total = my_tiles * n_st0;
        need_resolve = total > 0;
        int filled = 0;                                      // stages delivered: a tile's Z blocks follow its last stage
        // one copy of the stage code (run-time slot instead of prologue + main loop unrolled over the raw ring, see chain_tc.cu)
#pragma unroll 1
        for (int d = -RAW; d < total; ++d) {
            const int slot = (d + RAW) % RAW;
            if (need_resolve) { resolve(ltile); need_resolve = false; }    // index / scale loads in front of the wait inside fill()
            if (d >= 0) {
                fill(slot, sc[slot]);
                if (++filled % n_st0 == 0 && has_z) z_blocks(tile0 + (filled / n_st0 - 1) * tstride);
            }
            issue(slot, sc[slot]);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else if (warp == WW_EPI_WARPS + WW_PROD_WARPS) {
        // ================= MMA issue ==============================================================================
        if (lane == 0) {
            constexpr uint64_t DESC_FIXED = ((uint64_t)(128 >> 4) << 32) | (1ull << 46);          // SBO = 128 B, version bit
            const uint64_t a_desc0 = DESC_FIXED | ((uint64_t)((WTM * 16) >> 4) << 16);             // LBO = 2048 B
            // ring bases in 16-byte units, reduced to the descriptor's 14-bit start-address field: inside a cluster the
            // shared-window address of a CTA carries its rank in the high bits, which must not leak into the LBO field
            const uint32_t g_a = (smem_u32(sG) >> 4) & 0x3FFFu;
            const uint32_t h_own = (smem_u32(sHL) >> 4) & 0x3FFFu, h_peer = (smem_u32(sHR) >> 4) & 0x3FFFu;
            constexpr uint32_t ST16 = WW_STAGE_BYTES >> 4, LO16 = (4 * WTM * 16) >> 4, P16 = (2 * WTM * 16) >> 4;
            const uint32_t slot16 = SLOT_BYTES >> 4;
            const uint32_t wfull0 = smem_u32(&s_wfull[0]), wempty0 = smem_u32(&s_wempty[0]);
            uint32_t ws = 0, wpar = 0;
            int gs = 0, hs = 0; uint32_t gpar = 0, hpar = 0;
            int L = 0;
            uint32_t fin_pending = 0, finph = 0;
#ifdef HRN_WIDE_PROF
            long long wp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            const long long wp_start = clock64();
#endif
            auto claim = [&](int b) {
                if (fin_pending >> b & 1) {
                    WPROF_BEGIN();
                    mbar_wait(smem_u32(&s_fin[b]), (finph >> b) & 1);
                    WPROF_END(4);
                    finph ^= 1u << b; fin_pending &= ~(1u << b);
                }
            };
            auto piece_mma = [&](uint32_t a16, uint64_t w_desc0, uint32_t wlo16, uint32_t idesc, uint32_t d, uint32_t accumulate) {
                {
                    WPROF_BEGIN();
                    mbar_wait(wfull0 + 8 * ws, wpar);
                    WPROF_END(3);
                }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t w16 = ((ring_a >> 4) & 0x3FFFu) + ws * slot16;
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
            for (int tile = tile0; tile < n_tiles; tile += tstride) {
                // ---- layer 1: input stages as the producers deliver them ----
                {
                    const int b = L & 1;
                    const uint32_t d = tmem + b * WW_ACC;
                    claim(b);
                    const uint32_t N = (uint32_t)A.nh[0], idesc = umma_idesc_m128<PREC>(A.nh[0]);
                    const uint64_t w_desc0 = DESC_FIXED | ((uint64_t)N << 16);                      // LBO = N * 16 B
                    const uint32_t wlo16 = 2 * N;
                    int left = A.chunks0 / 2;
                    for (int s = 0; s < n_st0; ++s, left -= 2) {
                        {
                            WPROF_BEGIN();
                            mbar_wait(smem_u32(&s_gfull[gs]), gpar);
                            WPROF_END(0);
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
                // ---- layers 2, 3: block pairs (rank 0 block j, rank 1 block j) as both CTAs' epilogues convert them ----
                for (int l = 1; l < 3; ++l, ++L) {
                    const int b = L & 1;
                    const uint32_t d = tmem + b * WW_ACC;
                    claim(b);
                    if (l == 2) fin_pending |= 1u << b;
                    const uint32_t N = (uint32_t)A.nh[l], idesc = umma_idesc_m128<PREC>(A.nh[l]);
                    const uint64_t w_desc0 = DESC_FIXED | ((uint64_t)N << 16);
                    const uint32_t wlo16 = 2 * N;
                    const int nb = A.nh[l - 1] / 32;
                    for (int j = 0; j < nb; ++j) {
#pragma unroll
                        for (uint32_t rk = 0; rk < 2; ++rk) {
                            const bool own = rk == rank;
                            {
                                WPROF_BEGIN();
                                mbar_wait(smem_u32(own ? &s_hlfull[hs] : &s_hrfull[hs]), hpar);
                                WPROF_END(own ? 1 : 2);
                            }
                            const uint32_t a16 = (own ? h_own : h_peer) + hs * ST16;
                            piece_mma(a16, w_desc0, wlo16, idesc, d, (j > 0 || rk > 0) ? 1u : 0u);
                            piece_mma(a16 + P16, w_desc0, wlo16, idesc, d, 1u);
                        }
                        umma_commit_mc(smem_u32(&s_hfree[hs]), (uint16_t)3);     // this pair is consumed: tell both CTAs
                        if (++hs == HS) { hs = 0; hpar ^= 1; }
                    }
                    umma_commit(smem_u32(&s_accf[b]));
                }
            }
#ifdef HRN_WIDE_PROF
            if (blockIdx.x == 0) {
                wp[5] = clock64() - wp_start;
                for (int i = 0; i < 8; ++i) g_wide_prof[i] = wp[i];
            }
#endif
        }
    } else if (warp == WW_EPI_WARPS + WW_PROD_WARPS + 1) {
        // ================= weight stream ==========================================================================
        if (lane == 0) {
            uint32_t ws = 0, wpar = 0;
            const uint8_t* base = A.W + (size_t)rank * (size_t)A.w_rank_bytes;
            const uint32_t pieces[3] = {(uint32_t)(A.chunks0 / 2), (uint32_t)(2 * A.nh[0] / 16), (uint32_t)(2 * A.nh[1] / 16)};
            for (int tile = tile0; tile < n_tiles; tile += tstride) {
                const uint8_t* src = base;
                for (int l = 0; l < 3; ++l) {
                    const uint32_t bytes = (uint32_t)A.nh[l] * (PREC == 1 ? 32u : 64u);     // hi (+ lo) plane of a K=16 piece
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
    } else {
        // ================= sender: finished blocks of ring HL -> the peer's ring HR ===================================
        if (lane == 0) {
            int hs = 0; uint32_t hpar = 0;
            const uint32_t hl_a = smem_u32(sHL);
            const uint32_t hr_peer = map_peer(smem_u32(sHR), peer);
            const uint32_t full_peer0 = map_peer(smem_u32(&s_hrfull[0]), peer);
            for (int tile = tile0; tile < n_tiles; tile += tstride) {
                for (int l = 0; l < 2; ++l) {
                    const int nb = A.nh[l] / 32;
                    for (int j = 0; j < nb; ++j) {
                        mbar_wait(smem_u32(&s_hlfull[hs]), hpar);           // written by the epilogue (fenced for the async proxy)
                        const uint32_t rb = full_peer0 + 8 * hs;
                        asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;"
                                     ::"r"(rb), "r"(BLOCK_BYTES) : "memory");
                        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(hr_peer + hs * WW_STAGE_BYTES), "r"(hl_a + hs * WW_STAGE_BYTES), "r"(BLOCK_BYTES), "r"(rb)
                                     : "memory");
                        if (++hs == HS) { hs = 0; hpar ^= 1; }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                                     // nobody exits while its peer may still write to it
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int KSEG, int RAW, int PREC>
cudaError_t launch_wide_one(const WideArgs& A, int smem, int budget, cudaStream_t st) {
    auto kern = chain_wide_kernel<KSEG, RAW, PREC>;
    static hrn_once_per_device attr;
    static int max_clusters[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (attr.need()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, budget);
        if (e != cudaSuccess) return e;
        // how many 2-CTA clusters of this kernel the device can hold at once (GPCs with an odd SM count leave SMs out)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * 74); cfg.blockDim = dim3(WW_THREADS); cfg.dynamicSmemBytes = (size_t)budget;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 64; }
        if (dev >= 0 && dev < 64) max_clusters[dev] = n;
    }
    int clusters = (dev >= 0 && dev < 64 && max_clusters[dev] > 0) ? max_clusters[dev] : 64;
    if (clusters > A.n_tiles) clusters = A.n_tiles;
    kern<<<2 * clusters, WW_THREADS, smem, st>>>(A);
    return cudaSuccess;
}

}  // namespace

// Three fused layers up to 512 wide + attention tail on a virtual rows matrix, one 2-CTA cluster per 128-row tile.
// W: per column half (rank 0, then rank 1, w_rank_bytes each) the packed K=16 pieces of the three layers in execution
// order (engine_tc.pack_chain_wide: the K order of layers 2 / 3 interleaves 32-column blocks of the two halves);
// bias: per half b1 | b2 | b3; n1..n3 = full widths (multiples of 64, <= 512); kseg = rows per group (8).
// Outputs: G [rows / kseg, n3] attentive feature, a [rows] attention weights (layers.py:384-390).  rows % 128 == 0.
static int chain_wide_impl(const hrn_rows_t* in, const void* W, long long w_rank_bytes, const float* bias, int n1, int n2,
                           int n3, int chunks0, int kseg, const float* Zb, const float* Zg, int ldz, float* G, float* a,
                           float* Yh, int head_act, long long rows, int prec, void* stream) {
    if (!in || !W || !bias || (!G && !Yh) || rows < 0 || in->n_seg < 1 || in->n_seg > 4 || w_rank_bytes <= 0) return HRN_ERR_BAD_ARG;
    if (prec != 1 && prec != 3) return HRN_ERR_BAD_ARG;
    const int nn[3] = {n1, n2, n3};
    for (int l = 0; l < 3; ++l) if (nn[l] % 64 || nn[l] > 512 || nn[l] < 64) return HRN_ERR_UNSUPPORTED;
    if (kseg != 8) return HRN_ERR_UNSUPPORTED;
    if (rows % WTM != 0 || (chunks0 & 1) || rows >= 0x7fffffffLL) return HRN_ERR_UNSUPPORTED;
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
    if ((Zb == nullptr) != (Zg == nullptr)) return HRN_ERR_BAD_ARG;
    if (Zg && (!in->gather_idx || in->group != kseg || in->rows_per_batch <= 0 || ldz < n1 || (ldz & 3) || ((uintptr_t)Zb & 15) ||
               ((uintptr_t)Zg & 15))) return HRN_ERR_BAD_ARG;
    if (rows == 0) return HRN_OK;
    WideArgs A;
    A.in = *in; A.W = (const uint8_t*)W; A.w_rank_bytes = w_rank_bytes; A.bias = bias; A.G = G; A.a = a;
    A.n_tiles = (int)(rows / WTM);
    int maxh = 0;
    for (int l = 0; l < 3; ++l) { A.nh[l] = nn[l] / 2; if (A.nh[l] > maxh) maxh = A.nh[l]; }
    A.chunks0 = chunks0;
    A.Zb = Zb; A.Zg = Zg; A.ldz = ldz;
    A.Yh = Yh; A.head_act = head_act;
#ifdef HRN_WIDE_DEBUG
    A.dbg = g_wide_dbg;
#endif
    A.slot_bytes = maxh * (prec == 1 ? 32 : 64);
    // shared memory: ring G | ring HL | ring HR (16 KB stages) | weight ring | raw fp32 ring | biases | row maxima
    const int budget = 227 * 1024 - 1024;
    int raw = 2;
    A.gs = 2; A.hs = 2;
    A.zs = Zg ? 2 : 0;
    // one input stage per tile: no need to run stages ahead; the shared memory goes to a third slot of the block rings
    // instead (a 16 KB block takes ~800 cycles through DSMEM: with two slots the MMA thread waited 27 % of its time for
    // the peer's blocks, with three 8 %)
    if (Zg && chunks0 <= 4) { A.gs = 1; raw = 1; A.hs = 3; }
    int ring_want = WW_RING_MAX;
    if (const char* e = getenv("HRN_WIDE_CFG")) {               // tuning: "gs,hs,raw,ring"
        int g_ = 0, h_ = 0, r_ = 0, w_ = 0;
        if (sscanf(e, "%d,%d,%d,%d", &g_, &h_, &r_, &w_) == 4 && g_ >= 1 && g_ <= WW_GS_MAX && h_ >= 2 && h_ <= WW_HS_MAX &&
            r_ >= 1 && r_ <= 4 && w_ >= 3 && w_ <= WW_RING_MAX) { A.gs = g_; A.hs = h_; raw = r_; ring_want = w_; }
        if (const char* z = getenv("HRN_WIDE_ZS")) { const int zz = atoi(z); if (Zg && zz >= 2 && zz <= WW_ZS_MAX) A.zs = zz; }
    }
    const int fixed = (A.gs + 2 * A.hs + raw) * WW_STAGE_BYTES + 3 * 256 * 4 + 4 * WTM * 4 + A.zs * WW_ZSTAGE_BYTES;
    int ring = (budget - fixed) / A.slot_bytes;
    if (ring > ring_want) ring = ring_want;
    if (ring < 3) return HRN_ERR_UNSUPPORTED;
    A.ring = ring;
    const int smem = fixed + ring * A.slot_bytes;
    cudaStream_t st = (cudaStream_t)stream;
    if (prec == 1) {
        if (raw == 1) HRN_CUDA((launch_wide_one<8, 1, 1>(A, smem, budget, st)));
        else if (raw == 2) HRN_CUDA((launch_wide_one<8, 2, 1>(A, smem, budget, st)));
        else if (raw == 3) HRN_CUDA((launch_wide_one<8, 3, 1>(A, smem, budget, st)));
        else HRN_CUDA((launch_wide_one<8, 4, 1>(A, smem, budget, st)));
    } else {
        if (raw == 1) HRN_CUDA((launch_wide_one<8, 1, 3>(A, smem, budget, st)));
        else if (raw == 2) HRN_CUDA((launch_wide_one<8, 2, 3>(A, smem, budget, st)));
        else if (raw == 3) HRN_CUDA((launch_wide_one<8, 3, 3>(A, smem, budget, st)));
        else HRN_CUDA((launch_wide_one<8, 4, 3>(A, smem, budget, st)));
    }
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}

#ifdef HRN_WIDE_DEBUG
HRN_API int hrn_chain_wide_set_debug(float* p) { g_wide_dbg = p; return 0; }
#endif

#ifdef HRN_WIDE_PROF
HRN_API int hrn_chain_wide_prof(long long* host8) { return (int)cudaMemcpyFromSymbol(host8, g_wide_prof, 8 * sizeof(long long)); }
#endif

HRN_API int hrn_chain_wide(const hrn_rows_t* in, const void* W, long long w_rank_bytes, const float* bias, int n1, int n2,
                           int n3, int chunks0, int kseg, const float* Zb, const float* Zg, int ldz, float* G, float* a,
                           long long rows, int prec, void* stream) {
    if (!G) return HRN_ERR_BAD_ARG;
    return chain_wide_impl(in, W, w_rank_bytes, bias, n1, n2, n3, chunks0, kseg, Zb, Zg, ldz, G, a, nullptr, 0, rows, prec, stream);
}

// Per-keypoint head of up to 512-wide layers (CoarseReg's confidence head 512 -> 512 -> 512 -> 1, layers.py:391-394) on
// the same 2-CTA-cluster kernel: Y[r] = act(column 0 of the third layer).  n3 = the last layer zero-padded to 64 columns.
HRN_API int hrn_chain_wide_head(const hrn_rows_t* in, const void* W, long long w_rank_bytes, const float* bias, int n1,
                                int n2, int n3, int chunks0, int act, float* Y, long long rows, int prec, void* stream) {
    if (!Y) return HRN_ERR_BAD_ARG;
    return chain_wide_impl(in, W, w_rank_bytes, bias, n1, n2, n3, chunks0, 8, nullptr, nullptr, 0, nullptr, nullptr, Y, act, rows, prec,
                           stream);
}
