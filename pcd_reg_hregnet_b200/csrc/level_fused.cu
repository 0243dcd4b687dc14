// Fused hierarchy level of HierFeatureExtraction: grouping + KeypointDetector + DescExtractor in ONE persistent
// kernel -- no per-neighbour tensor ever reaches HBM.
//
// Replaces, per level (reference models/HRegNet/models.py:27-28 / 33-34):
//   knn_group (layers.py:9-27)  ->  KeypointDetector.convs / attention / keypoints / attentive feature
//   (layers.py:150-159)  ->  DescExtractor.convs / max / cat / mlp1 / mlp2 / max (layers.py:200-209).
// Unfused, level 1 alone moves ~11 GB per 32-pair step through HBM (4.2 M rows x 64 channels x a dozen passes);
// here the inputs are the neighbour index + coordinates (+ the gathered 64-channel feature row at level 2) and the
// outputs are per KEYPOINT only (xyz, attentive feature, descriptor).
//
// CTA = 128 rows = 128/k keypoints x k neighbours; thread t <-> row t <-> TMEM lane t; persistent over tiles.
// Every layer: A operand in smem -> tcgen05.mma (bf16 hi/lo split of both operands, 3 MMAs per k-step, fp32
// accumulate in TMEM) -> tcgen05.ld -> bias+ReLU -> bf16 hi/lo split -> written back IN PLACE as the next layer's
// A operand.  Chain per tile (the detector and the descriptor stack do not depend on each other, so their second and
// third layers are issued TOGETHER -- six MMA round trips per tile, each one an issue -> commit -> mbarrier wait):
//   G -[d1;x1]-> C1d | C1x   (the two first layers share their input: ONE step with the weights stacked along N)
//   C1d -d2-> C2 -d3-> E(CO, stays in TMEM)          | max_c, softmax over the k neighbours, keypoint
//   C1x -x2-> C2 -x3-> X1(CO)                        | column max over the group
//   mlp1 = Wb.X1 + Wc.(E*a)  (two K-segments accumulated in TMEM)  + [Wa.max_k(X1) + b]  -> CMID -mlp2-> CD -> max_k
// The reference's third mlp1 segment, max_k(X1) repeated over the k neighbours (layers.py:203-204), is constant inside a
// group: Wa.max_k(X1) + b is evaluated ONCE per keypoint on the CUDA cores in fp32 (while the tensor core runs the X1
// segment) and enters the mlp1 epilogue as a per-keypoint bias (as in level_ws.cu) -- 8 MMA steps per tile instead of 9.
// Four independent 128-thread groups per CTA, each with its own tile, operand buffer and TMEM columns, overlap each
// other's MMA / epilogue phases; the weights (52 KB of bf16 hi/lo UMMA tiles + the 8 KB fp32 Wa) are RESIDENT in shared
// memory and shared by the groups.  (Levels 2 and 3, whose weights have to stream, run on level_ws.cu.)
//
// Epilogue instruction diet (the kernel is bound by the SIMT work between the MMAs, not by the tensor pipe):
//   * ReLU is folded into the operand conversions: hi = cvt.rz.relu.bf16x2(x), lo = cvt.rn.relu.bf16x2(x - hi) -- with
//     a TRUNCATED hi the residual of a non-negative x is non-negative, and for x < 0 both halves clamp to 0;
//   * the biases ride in the MMAs: the grouped input of the merged first layer has 12 spare K columns, one of them is the
//     constant 1 and its weight column is the bias; the later layers get one extra K=16 piece (ones x [b_hi, b_lo]);
//   * max_k relu(x) = relu(max_k x): the column maxima are taken on the pre-activations.
// Biases of d2 d3 x2 x3 m2 as one more K=16 MMA piece per layer (a resident block of ones times the bias split into bf16
// hi + lo) instead of additions in the drains: level 1 775 -> 676 us (the kernel is bound by its SIMT work, the tensor pipe
// has room).  Comment the define out for the additive form.
#define L1_BIAS_MMA 1
#include "common.cuh"
#include "tc_common.cuh"
#include <math_constants.h>
#include <type_traits>

namespace {

constexpr int TMR = 128;
constexpr int cmax(int a, int b) { return a > b ? a : b; }

template <int KNBR_, int CIN_, int C1_, int C2_, int CO_, int CMID_, int CD_>
struct LevelCfg {
    static constexpr int KNBR = KNBR_, CIN = CIN_, C1 = C1_, C2 = C2_, CO = CO_, CMID = CMID_, CD = CD_;
    static constexpr int KG = (CIN + 5 + 15) / 16 * 16;             // grouped input: [feat(CIN) | rel xyz, |rel|, 1 | 0-pad]
    // weight blocks in execution order; block = hi plane [K/8][N][16 B] + lo plane = 4*K*N bytes
    static constexpr int NL = 8;                                    // blocks: [d1;x1] d2 d3 x2 x3 mb mc m2
    __host__ __device__ static constexpr int lk(int l) {            // K of block l
        return l == 0 ? KG : (l == 1 || l == 3) ? C1 : (l == 2 || l == 4) ? C2 : (l == 7) ? CMID : CO;
    }
    __host__ __device__ static constexpr int ln(int l) {
        return l == 0 ? 2 * C1 : (l == 1 || l == 3) ? C2 : (l == 2 || l == 4) ? CO : (l == 7) ? CD : CMID;
    }
    __host__ __device__ static constexpr int woff(int l) { int o = 0; for (int i = 0; i < l; ++i) o += 4 * lk(i) * ln(i); return o; }
    static constexpr int W_BYTES = woff(NL);
    // fp32 Wa (the max_k(X1) block of mlp1) behind the MMA blocks: two half-matrices [CO/2][CMID] -- input channels c with
    // (c >> 2) & 1 == h in half h -- 16 floats apart modulo the banks, so that the two threads of an output read
    // different banks
    static constexpr int WA_HALF = CO / 2 * CMID + 16;
    static constexpr int WA_BYTES = 2 * WA_HALF * 4;
#ifdef L1_BIAS_MMA
    // biases of d2 d3 x2 x3 m2 as one more K=16 MMA piece per layer: A = a resident block of ones (columns 0 and 1), B = the
    // bias split into bf16 hi (K column 0) and lo (K column 1) -- the drains then add nothing
    __host__ __device__ static constexpr int bp_off(int l) {          // byte offset of layer l's bias piece behind the Wa halves
        return l == 1 ? 0 : l == 2 ? 32 * C2 : l == 3 ? 32 * (C2 + CO) : l == 4 ? 32 * (2 * C2 + CO) : 32 * (2 * C2 + 2 * CO);
    }
    static constexpr int BP_BYTES = 32 * (2 * C2 + 2 * CO + CD);
    static constexpr int ONES_BYTES = 2 * TMR * 16;
#else
    static constexpr int BP_BYTES = 0, ONES_BYTES = 0;
#endif
    static constexpr int PACK_BYTES = W_BYTES + WA_BYTES + BP_BYTES;
    // biases (floats): d1,d2,d3,x1,x2,x3,m1,m2 (d1 / x1 ride in the first MMA and are not read here)
    static constexpr int B_D1 = 0, B_D2 = B_D1 + C1, B_D3 = B_D2 + C2, B_X1 = B_D3 + CO, B_X2 = B_X1 + C1,
                         B_X3 = B_X2 + C2, B_M1 = B_X3 + CO, B_M2 = B_M1 + CMID, B_COUNT = B_M2 + CD;
    // in-place operand buffer (chunks of 8 channels); the grouped input is consumed by one step, so it lives here too
    static constexpr int OPC = cmax(cmax(cmax(C1, C2), cmax(CO, CMID)), KG) / 8;
    static constexpr int OP_BYTES = 2 * OPC * TMR * 16;
    static constexpr int CW = cmax(CO, CD);
    // TMEM columns: [E] [work: hidden layers / mlp1 accumulator / X1 / descriptor].  One work region is enough: every
    // result in it is drained (to the operand buffer or to HBM) before the next layer that targets it is issued.
    static constexpr int T_ACCE = 0;
    static constexpr int T_ACC0 = CO;
    static constexpr int T_ACCX = CO;
    static constexpr int T_USED = CO + cmax(cmax(cmax(2 * C1, C2), CMID), cmax(CO, CD));
    static_assert(2 * C1 <= OPC * 8 && 2 * C2 <= OPC * 8, "the operand buffer holds the inputs of both chains side by side");
    static_assert(2 * C2 <= cmax(cmax(cmax(2 * C1, C2), CMID), cmax(CO, CD)), "work region holds both second-layer results");
    static constexpr int WPG = KNBR / 32;                           // warps per keypoint group
    static constexpr int KPT = TMR / KNBR;                          // keypoints per tile
    // NG independent 128-thread groups per CTA, each with its own tile, operand buffers and TMEM columns, all sharing
    // ONE resident copy of the weights (4 tiles in flight per SM)
    static constexpr int NG = 4;
    static constexpr int GRP_SMEM = OP_BYTES + 2 * 4 * CW * 4 + KPT * CMID * 4;
    static constexpr int SMEM_NG = PACK_BYTES + ONES_BYTES + B_COUNT * 4 + NG * GRP_SMEM + 256;
    static constexpr int T_COLS_NG = NG * T_USED <= 256 ? 256 : 512;
    static_assert(NG * T_USED <= 512, "TMEM (groups)");
    static_assert(KNBR == 64 && CMID == 32 && CO % 8 == 0, "the per-keypoint mat-vec maps 64 threads onto 32 outputs x 2 halves");
    static_assert(CIN % 8 == 0, "feature chunks must be 8-aligned");
    static_assert(SMEM_NG <= 227 * 1024, "shared memory");
};

__device__ __forceinline__ uint32_t make_idesc(int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TMR >> 4) << 24);
}

// one thread: D[tmem_col] (+)= A(smem, K, hi/lo planes) x W(smem, [N x K], hi/lo planes)^T, bf16x3
__device__ __forceinline__ void issue_layer(uint32_t a_hi, uint32_t a_lo, int K, uint32_t w_hi, int N, uint32_t tmem_d,
                                            bool accumulate, uint32_t ones = 0, uint32_t bias_piece = 0) {
    const uint32_t idesc = make_idesc(N);
    const uint32_t w_lbo = (uint32_t)N * 16;
    const uint64_t a_fix = umma_desc_fixed(TMR * 16, 128), w_fix = umma_desc_fixed(w_lbo, 128);
    // start addresses in 16-byte units (shared memory < 256 KB: they fit the 14-bit field); one k-step = two chunks
    uint32_t ah = a_hi >> 4, al = a_lo >> 4, wh = w_hi >> 4, wl = (w_hi + (uint32_t)(K / 8) * w_lbo) >> 4;
    const uint32_t a_step = (2 * TMR * 16) >> 4, w_step = (2 * w_lbo) >> 4;
#pragma unroll 4
    for (int k = 0; k < K / 16; ++k) {
        umma_bf16(tmem_d, a_fix | ah, w_fix | wh, idesc, (accumulate || k > 0) ? 1u : 0u);
        umma_bf16(tmem_d, a_fix | al, w_fix | wh, idesc, 1u);
        umma_bf16(tmem_d, a_fix | ah, w_fix | wl, idesc, 1u);
        ah += a_step; al += a_step; wh += w_step; wl += w_step;
    }
    if (bias_piece) umma_bf16(tmem_d, a_fix | (ones >> 4), w_fix | (bias_piece >> 4), idesc, 1u);      // + 1 * b_hi + 1 * b_lo
}

// f[e] = acc[e] (+ bias[e]) for 32 consecutive columns, NO activation; the biases come from shared memory as 8 x 16-byte
// loads and are added two at a time (packed fp32x2)
template <bool HAS_BIAS>
__device__ __forceinline__ void bias32(const uint32_t (&v)[32], const float* __restrict__ bb, float (&f)[32]) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        if (HAS_BIAS) {
            const float4 b4 = *reinterpret_cast<const float4*>(bb + 4 * q);
            f2_unpack(f2_add(f2_pack(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1])), f2_pack(b4.x, b4.y)), f[4 * q], f[4 * q + 1]);
            f2_unpack(f2_add(f2_pack(__uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3])), f2_pack(b4.z, b4.w)), f[4 * q + 2], f[4 * q + 3]);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) f[4 * q + e] = __uint_as_float(v[4 * q + e]);
        }
    }
}

// lane c receives the op-reduction over the 32 lanes of v[c]  (31 shuffles)
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

template <class Cfg>
__global__ void __launch_bounds__(TMR * Cfg::NG, 1)
level_fused_kernel(const float* __restrict__ q, const float* __restrict__ xyz, const float* __restrict__ feat,
                   const int32_t* __restrict__ idx, const uint8_t* __restrict__ Wpack, const float* __restrict__ biases,
                   float* __restrict__ out_xyz, float* __restrict__ out_af, float* __restrict__ out_desc, int M, int N,
                   int n_tiles) {
    constexpr int KNBR = Cfg::KNBR, CIN = Cfg::CIN, C1 = Cfg::C1, C2 = Cfg::C2, CO = Cfg::CO, CMID = Cfg::CMID,
                  CD = Cfg::CD, WPG = Cfg::WPG, CW = Cfg::CW;
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int NG = Cfg::NG;
    __shared__ __align__(8) uint64_t s_bar[NG];                           // MMA done (group g)
    __shared__ uint32_t s_tmem;
    __shared__ float s_red_all[NG][4][8];

    const int cta_tid = threadIdx.x;
    const int grp_id = cta_tid / TMR;                                     // independent 128-thread group
    const int tid = cta_tid % TMR, warp = tid >> 5, lane = tid & 31;      // group-local ids (warp = TMEM lane quarter)
    float (*s_red)[8] = s_red_all[grp_id];
    uint8_t* sW = smem;
    const float* sWa = reinterpret_cast<const float*>(smem + Cfg::W_BYTES);
    uint8_t* sOnes = smem + Cfg::PACK_BYTES;
    float* sB = reinterpret_cast<float*>(smem + Cfg::PACK_BYTES + Cfg::ONES_BYTES);
    uint8_t* sOp = smem + Cfg::PACK_BYTES + Cfg::ONES_BYTES + Cfg::B_COUNT * 4 + grp_id * Cfg::GRP_SMEM;
    float* sCol = reinterpret_cast<float*>(sOp + Cfg::OP_BYTES);          // [4][CW] per-warp column partials
    float* sCol2 = sCol + 4 * CW;
    float* sKpb = sCol2 + 4 * CW;                                         // [KPT][CMID] per-keypoint bias of mlp1
    auto gsync = [&]() {                                                  // barrier of this group only
        asm volatile("bar.sync %0, %1;" ::"r"(grp_id + 1), "n"(TMR) : "memory");
    };

    const int gw0 = (warp / WPG) * WPG;                                   // first warp of this row's keypoint group
    const bool leader = (warp == gw0);
    const uint32_t bar = smem_u32(&s_bar[grp_id]);
    uint32_t phase = 0;
    const int vgrid = (int)gridDim.x * NG, vblock = (int)blockIdx.x * NG + grp_id;   // groups act as virtual CTAs
    const uint32_t wbase = smem_u32(sW);

    // ---- one-time setup ---------------------------------------------------------------------------------------
    if (cta_tid == 0) {
        for (int i = 0; i < NG; ++i) mbar_init(smem_u32(&s_bar[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = cta_tid; i < Cfg::PACK_BYTES / 16; i += TMR * NG)
        reinterpret_cast<uint4*>(sW)[i] = __ldg(reinterpret_cast<const uint4*>(Wpack) + i);
    for (int i = cta_tid; i < Cfg::B_COUNT; i += TMR * NG) sB[i] = __ldg(biases + i);
#ifdef L1_BIAS_MMA
    for (int i = cta_tid; i < 2 * TMR; i += TMR * NG)             // [2 chunks][128 rows][8 bf16]: columns 0 and 1 of chunk 0 = 1.0
        reinterpret_cast<uint4*>(sOnes)[i] = i < TMR ? make_uint4(0x3F803F80u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
#endif
    if (cta_tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(Cfg::T_COLS_NG) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_all = s_tmem;
    const uint32_t tmem = tmem_all + grp_id * Cfg::T_USED;                // this group's accumulator columns
    const uint32_t lane_base = ((uint32_t)(warp * 32) << 16);
    const uint32_t aOp_hi = smem_u32(sOp), aOp_lo = aOp_hi + Cfg::OPC * TMR * 16;
    uint4* op_hi = reinterpret_cast<uint4*>(sOp);
    uint4* op_lo = op_hi + Cfg::OPC * TMR;

    // operand ready in smem -> one thread issues layer `li` (operand = chunks [ch0, ch0 + K/8) of the buffer), optionally a
    // second, independent layer `lj` behind it under the same commit
    auto issue2 = [&](int li, int ch0, int tcol, bool acc, int lj, int ch1, int tcol1) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        gsync();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#ifdef L1_BIAS_MMA
            const uint32_t ones = smem_u32(sOnes), bp0 = wbase + Cfg::W_BYTES + Cfg::WA_BYTES;
            auto bp = [&](int l) -> uint32_t { return (l == 1 || l == 2 || l == 3 || l == 4 || l == 7) ? bp0 + Cfg::bp_off(l) : 0u; };
            issue_layer(aOp_hi + ch0 * TMR * 16, aOp_lo + ch0 * TMR * 16, Cfg::lk(li), wbase + Cfg::woff(li), Cfg::ln(li), tmem + tcol, acc, ones, bp(li));
            if (lj >= 0)
                issue_layer(aOp_hi + ch1 * TMR * 16, aOp_lo + ch1 * TMR * 16, Cfg::lk(lj), wbase + Cfg::woff(lj), Cfg::ln(lj), tmem + tcol1, false, ones, bp(lj));
#else
            issue_layer(aOp_hi + ch0 * TMR * 16, aOp_lo + ch0 * TMR * 16, Cfg::lk(li), wbase + Cfg::woff(li), Cfg::ln(li), tmem + tcol, acc);
            if (lj >= 0)
                issue_layer(aOp_hi + ch1 * TMR * 16, aOp_lo + ch1 * TMR * 16, Cfg::lk(lj), wbase + Cfg::woff(lj), Cfg::ln(lj), tmem + tcol1, false);
#endif
            umma_commit(bar);
        }
    };
    auto issue = [&](int li, int tcol, bool acc) { issue2(li, 0, tcol, acc, -1, 0, 0); };
    // -> everybody waits for the accumulator
    auto wait_layer = [&]() {
        mbar_wait(bar, phase);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    };
    auto run_layer = [&](int li, int tcol, bool acc) { issue(li, tcol, acc); wait_layer(); };
    // accumulator [tcol, tcol+Nn) -> relu(x (+ b)) -> bf16 hi/lo operand (in place)
    auto epi_to_operand = [&](int tcol, int Nn, const float* bb, auto has_bias, int chd = 0) {
        for (int c0 = 0; c0 < Nn; c0 += 32) {
            uint32_t v[32];
            float f[32];
            tmem_ld32(tmem + lane_base + tcol + c0, v);
            bias32<decltype(has_bias)::value>(v, bb + c0, f);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
                split_store8_relu(f + ch * 8, op_hi + (chd + c0 / 8 + ch) * TMR + tid, op_lo + (chd + c0 / 8 + ch) * TMR + tid);
        }
    };
    using yes_t = std::true_type;
    using no_t = std::false_type;
    // value of column c for this row's keypoint group after the per-warp partials were published in `buf`
    auto col_max = [&](const float* buf, int c) {
        float v = buf[gw0 * CW + c];
        if (WPG == 2) v = fmaxf(v, buf[(gw0 + 1) * CW + c]);
        return v;
    };

    for (int tile = vblock; tile < n_tiles; tile += vgrid) {
        // ---- grouped input  [feat[idx] | rel xyz, |rel|, 1] -> operand buffer ------------------------------------
        const long long r = (long long)tile * TMR + tid;
        const long long bm = r / KNBR;
        const long long b = bm / M;
        const int n = __ldg(idx + r);
        const float* pp = xyz + (b * N + n) * 3;
        const float* qq = q + bm * 3;
        const float nx = __ldg(pp), ny = __ldg(pp + 1), nz = __ldg(pp + 2);
        const float rx = nx - __ldg(qq), ry = ny - __ldg(qq + 1), rz = nz - __ldg(qq + 2);
        if (CIN > 0) {
            const float4* fr = reinterpret_cast<const float4*>(feat + (b * N + n) * CIN);
            constexpr int HALF = CIN / 8 > 1 ? CIN / 16 : 1;          // chunks per batch of loads (bounds registers)
#pragma unroll
            for (int h = 0; h < (CIN / 8) / HALF; ++h) {
                float4 fv[2 * HALF];
#pragma unroll
                for (int i = 0; i < 2 * HALF; ++i) fv[i] = __ldg(fr + h * 2 * HALF + i);
#pragma unroll
                for (int c = 0; c < HALF; ++c) {
                    const float x[8] = {fv[2 * c].x, fv[2 * c].y, fv[2 * c].z, fv[2 * c].w,
                                        fv[2 * c + 1].x, fv[2 * c + 1].y, fv[2 * c + 1].z, fv[2 * c + 1].w};
                    split_store8(x, op_hi + (h * HALF + c) * TMR + tid, op_lo + (h * HALF + c) * TMR + tid);
                }
            }
        }
        {
            // the constant 1 multiplies the bias column of the merged first layer
            const float x[8] = {rx, ry, rz, sqrtf(rx * rx + ry * ry + rz * rz), 1.f, 0.f, 0.f, 0.f};
            split_store8(x, op_hi + (CIN / 8) * TMR + tid, op_lo + (CIN / 8) * TMR + tid);
#pragma unroll
            for (int c = CIN / 8 + 1; c < Cfg::KG / 8; ++c) {             // K padding
                op_hi[c * TMR + tid] = make_uint4(0, 0, 0, 0);
                op_lo[c * TMR + tid] = make_uint4(0, 0, 0, 0);
            }
        }
        // ---- first layer of both chains: work columns [0, C1) = detector, [C1, 2 C1) = descriptor ------------------
        run_layer(0, Cfg::T_ACC0, false);
        // ---- second layers d2 | x2 together: operands = chunks [0, C1/8) | [C1/8, 2 C1/8), results in work [0, C2) | [C2, 2 C2)
        epi_to_operand(Cfg::T_ACC0, C1, sB, no_t{});
        epi_to_operand(Cfg::T_ACC0 + C1, C1, sB, no_t{}, C1 / 8);
        issue2(1, 0, Cfg::T_ACC0, false, 3, C1 / 8, Cfg::T_ACC0 + C2);
        wait_layer();
        // ---- third layers d3 | x3 together: E -> its own columns, X1 -> the work region (its inputs are drained by then) ----
#ifdef L1_BIAS_MMA
        epi_to_operand(Cfg::T_ACC0, C2, sB, no_t{});
        epi_to_operand(Cfg::T_ACC0 + C2, C2, sB, no_t{}, C2 / 8);
#else
        epi_to_operand(Cfg::T_ACC0, C2, sB + Cfg::B_D2, yes_t{});
        epi_to_operand(Cfg::T_ACC0 + C2, C2, sB + Cfg::B_X2, yes_t{}, C2 / 8);
#endif
        issue2(2, 0, Cfg::T_ACCE, false, 4, C2 / 8, Cfg::T_ACCX);
        wait_layer();
        // ---- attention: a = softmax_k(max_c E), keypoint = sum_k a * nn -------------------------------------
        float x1 = 0.f;                                                   // post-ReLU values are >= 0
        for (int c0 = 0; c0 < CO; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem + lane_base + Cfg::T_ACCE + c0, v);
#ifdef L1_BIAS_MMA
#pragma unroll
            for (int q = 0; q < 8; ++q)
                x1 = fmaxf(x1, fmaxf(fmaxf(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1])),
                                     fmaxf(__uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]))));
#else
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 b4 = *reinterpret_cast<const float4*>(sB + Cfg::B_D3 + c0 + 4 * q);
                float s0, s1, s2, s3;
                f2_unpack(f2_add(f2_pack(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1])), f2_pack(b4.x, b4.y)), s0, s1);
                f2_unpack(f2_add(f2_pack(__uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3])), f2_pack(b4.z, b4.w)), s2, s3);
                x1 = fmaxf(x1, fmaxf(fmaxf(s0, s1), fmaxf(s2, s3)));
            }
#endif
        }
        float gmax = hrn_warp_max(x1);
        if (WPG == 2) {
            if (lane == 0) s_red[warp][0] = gmax;
            gsync();
            gmax = fmaxf(s_red[gw0][0], s_red[gw0 + 1][0]);
        }
        const float ex = expf(x1 - gmax);
        float s0 = hrn_warp_sum(ex), s1 = hrn_warp_sum(ex * nx), s2 = hrn_warp_sum(ex * ny), s3 = hrn_warp_sum(ex * nz);
        if (WPG == 2) {
            if (lane == 0) { s_red[warp][1] = s0; s_red[warp][2] = s1; s_red[warp][3] = s2; s_red[warp][4] = s3; }
            gsync();
            s0 = s_red[gw0][1] + s_red[gw0 + 1][1]; s1 = s_red[gw0][2] + s_red[gw0 + 1][2];
            s2 = s_red[gw0][3] + s_red[gw0 + 1][3]; s3 = s_red[gw0][4] + s_red[gw0 + 1][4];
        }
        const float a = ex / s0;
        if (leader && lane < 3) out_xyz[bm * 3 + lane] = (lane == 0 ? s1 : (lane == 1 ? s2 : s3)) / s0;
        // ---- descriptor head ------------------------------------------------------------------------------------
        // X1 -> operand, and its column maximum over the rows of the group (taken before the ReLU: max and ReLU commute)
        for (int c0 = 0; c0 < CO; c0 += 32) {
            uint32_t v[32];
            float f[32];
            tmem_ld32(tmem + lane_base + Cfg::T_ACCX + c0, v);
#ifdef L1_BIAS_MMA
            bias32<false>(v, sB, f);
#else
            bias32<true>(v, sB + Cfg::B_X3 + c0, f);
#endif
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
                split_store8_relu(f + ch * 8, op_hi + (c0 / 8 + ch) * TMR + tid, op_lo + (c0 / 8 + ch) * TMR + tid);
            const float cm = warp_transpose_reduce<true>(f, lane);
            sCol[warp * CW + c0 + lane] = fmaxf(cm, 0.f);
        }
        issue(5, Cfg::T_ACC0, false);                                     // mlp1 = Wb.X1 (the barrier inside publishes sCol)
        // per-keypoint bias  Wa . max_k(X1) + b  on the CUDA cores (fp32) while the tensor core runs the X1 segment:
        // thread (output j, half hf) of the keypoint of its own rows sums the input channels c with (c >> 2) & 1 == hf
        {
            const int u = tid & 63, j = u >> 1, hf = u & 1;
            const float* wa = sWa + hf * Cfg::WA_HALF + j;
            const float* m0 = sCol + gw0 * CW + 4 * hf;
            float acc = hf ? 0.f : sB[Cfg::B_M1 + j];
#pragma unroll
            for (int i = 0; i < CO / 8; ++i) {
                const float4 p0 = *reinterpret_cast<const float4*>(m0 + 8 * i);
                const float4 p1 = *reinterpret_cast<const float4*>(m0 + CW + 8 * i);
                acc = fmaf(wa[(4 * i + 0) * CMID], fmaxf(p0.x, p1.x), acc);
                acc = fmaf(wa[(4 * i + 1) * CMID], fmaxf(p0.y, p1.y), acc);
                acc = fmaf(wa[(4 * i + 2) * CMID], fmaxf(p0.z, p1.z), acc);
                acc = fmaf(wa[(4 * i + 3) * CMID], fmaxf(p0.w, p1.w), acc);
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            if (hf == 0) sKpb[(tid >> 6) * CMID + j] = acc;
        }
        wait_layer();
        // attentive feature map E*a as the second K-segment; attentive feature = its column sum over the group
        for (int c0 = 0; c0 < CO; c0 += 32) {
            uint32_t v[32];
            float f[32];
            tmem_ld32(tmem + lane_base + Cfg::T_ACCE + c0, v);
#ifdef L1_BIAS_MMA
            bias32<false>(v, sB, f);
#else
            bias32<true>(v, sB + Cfg::B_D3 + c0, f);
#endif
            const f32x2_t a2 = f2_pack(a, a);
#pragma unroll
            for (int e = 0; e < 32; e += 2) f2_unpack(f2_mul(f2_pack(fmaxf(f[e], 0.f), fmaxf(f[e + 1], 0.f)), a2), f[e], f[e + 1]);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                const float x[8] = {f[ch * 8], f[ch * 8 + 1], f[ch * 8 + 2], f[ch * 8 + 3], f[ch * 8 + 4], f[ch * 8 + 5],
                                    f[ch * 8 + 6], f[ch * 8 + 7]};
                split_store8(x, op_hi + (c0 / 8 + ch) * TMR + tid, op_lo + (c0 / 8 + ch) * TMR + tid);
            }
            const float cs = warp_transpose_reduce<false>(f, lane);
            sCol2[warp * CW + c0 + lane] = cs;
        }
        run_layer(6, Cfg::T_ACC0, true);                                  // mlp1 += Wc.(E*a)  (barrier: sCol2, sKpb published)
        if (leader)
            for (int c = lane; c < CO; c += 32) {
                float v = sCol2[gw0 * CW + c];
                if (WPG == 2) v += sCol2[(gw0 + 1) * CW + c];
                out_af[bm * CO + c] = v;
            }
        // ---- mlp1 epilogue (bias = per-keypoint row) -> mlp2 -> descriptor = max_k ------------------------------
        epi_to_operand(Cfg::T_ACC0, CMID, sKpb + (tid >> 6) * CMID, yes_t{});
        run_layer(7, Cfg::T_ACCX, false);
        for (int c0 = 0; c0 < CD; c0 += 32) {
            uint32_t v[32];
            float f[32];
            tmem_ld32(tmem + lane_base + Cfg::T_ACCX + c0, v);
#ifdef L1_BIAS_MMA
            bias32<false>(v, sB, f);
#else
            bias32<true>(v, sB + Cfg::B_M2 + c0, f);
#endif
            const float cm = warp_transpose_reduce<true>(f, lane);
            sCol[warp * CW + c0 + lane] = fmaxf(cm, 0.f);
        }
        gsync();
        if (leader)
            for (int c = lane; c < CD; c += 32) out_desc[bm * CD + c] = col_max(sCol, c);
        // the next tile's first barrier (inside run_layer) orders these reads before sCol / s_red are rewritten
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (cta_tid < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_all), "n"(Cfg::T_COLS_NG) : "memory");
}

using CfgL1 = LevelCfg<64, 0, 32, 32, 64, 32, 64>;             // detector_1 / desc_extractor_1 (models.py:14,22)

template <class Cfg>
int launch_level(const float* q, const float* xyz, const float* feat, const int32_t* idx, const void* Wpack,
                 const float* biases, float* out_xyz, float* out_af, float* out_desc, int B, int M, int N, cudaStream_t st) {
    const int n_tiles = (int)((long long)B * M * Cfg::KNBR / TMR);
    auto kern = level_fused_kernel<Cfg>;
    static hrn_once_per_device attr;
    if (attr.need()) HRN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_NG));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int want = (n_tiles + Cfg::NG - 1) / Cfg::NG;
    const int grid = want < sms ? want : sms;
    kern<<<grid, TMR * Cfg::NG, Cfg::SMEM_NG, st>>>(q, xyz, feat, idx, (const uint8_t*)Wpack, biases, out_xyz, out_af, out_desc, M, N,
                                       n_tiles);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}

}  // namespace

// Fused detector + descriptor of hierarchy level 1.  q [B*M,3] sampled keypoint coordinates, xyz [B,N,3], feat = NULL
// (level 1 has no input features), idx [B*M*k] int32 neighbour indices; Wpack / biases from engine_tc.pack_level (LevelCfg
// layout, execution order, then the fp32 Wa halves).  Outputs per keypoint: out_xyz [B*M,3], out_af [B*M,CO], out_desc [B*M,CD].
HRN_API int hrn_level_fused(int level, const float* q, const float* xyz, const float* feat, const int32_t* idx,
                            const void* Wpack, const float* biases, float* out_xyz, float* out_af, float* out_desc,
                            int B, int M, int N, int k, void* stream) {
    if (!q || !xyz || !idx || !Wpack || !biases || !out_xyz || !out_af || !out_desc || B < 0 || M <= 0 || N <= 0) return HRN_ERR_BAD_ARG;
    if (((long long)B * M * k) % TMR != 0) return HRN_ERR_UNSUPPORTED;
    if (B == 0) return HRN_OK;
    if (level == 1 && k == CfgL1::KNBR && !feat)
        return launch_level<CfgL1>(q, xyz, nullptr, idx, Wpack, biases, out_xyz, out_af, out_desc, B, M, N, (cudaStream_t)stream);
    return HRN_ERR_UNSUPPORTED;     // levels 2 and 3: hrn_level_ws
}

HRN_API int hrn_level_pack_bytes(int level) { return level == 1 ? CfgL1::PACK_BYTES : -1; }
HRN_API int hrn_level_bias_count(int level) { return level == 1 ? CfgL1::B_COUNT : -1; }
