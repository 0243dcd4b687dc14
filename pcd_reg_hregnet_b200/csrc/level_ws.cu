// Fused hierarchy level of HierFeatureExtraction, warp-specialised and pipelined inside the tile: grouping +
// KeypointDetector + DescExtractor of ONE level in one persistent kernel, for the WIDE levels (level 3: widths
// 128/128/256, k = 16; level 2: 64/64/128, k = 32) whose weights do not fit in shared memory.
//
// Replaces, per level (reference models/HRegNet/models.py:33-34 / 39-40):
//   knn_group (layers.py:9-27) -> KeypointDetector.convs / attention / keypoints / attentive feature (layers.py:150-159)
//   -> DescExtractor.convs / max / cat / mlp1 / mlp2 / max (layers.py:200-209).
// Before this kernel level 3 ran as two chain launches + two per-layer launches + group kernels and moved ~1.3 GB per
// 32-pair step through HBM (the [rows, 256] tensors E*a, X1 and the mlp1 output); here nothing per-neighbour leaves the SM.
//
// Tile = 128 rows = 128/k keypoints x k neighbours.  Per tile and group (a group = 8 epilogue warps + 1 MMA warp +
// 1 weight-stream warp; NH = 2 threads per row, each draining alternate 32-column blocks of an accumulator):
//
//   G -[d1;x1]-> C1d | C1x            both drained at once: the two conv stacks do not depend on each other
//   C1d -d2-> C2d -d3-> E             attention a = softmax_k(max_c E), keypoint = sum_k a nn
//   C1x -x2-> C2x -x3-> X1            issued BESIDE the detector chain (x2 right behind d2, x3 under the attention phase)
//   E*a -Wc-> M1  (first K-segment of mlp1; its column sums over the group are the attentive feature)
//   X1 -Wb-> M1 += (second K-segment); column max of X1 over the group
//   M1 + Wa.max_k(X1) + b -> relu -m2-> D -> max_k = descriptor
//
// The reference's third mlp1 segment, max_k(X1) repeated over the k neighbours (layers.py:203-204), is constant inside a
// group: Wa.max_k(X1) is evaluated ONCE per keypoint on the CUDA cores in fp32 (a 2C x C mat-vec per keypoint, while the
// tensor core runs the X1 segment) and enters the mlp1 epilogue as a per-keypoint bias -- one K-segment (a third of
// mlp1's MMAs and weight stream) less, same sum up to fp32 summation order.
//
// Pipelining: an accumulator is drained in 32-column blocks (TMEM -> bias/ReLU -> bf16 hi/lo -> operand buffer); every
// block raises its own mbarrier, and the MMA warp issues the next layer's K=16 pieces as their blocks land, so the tensor
// core works on layer l+1 while layer l is still being drained.  The weights of all layers stream from L2 as K=16 pieces
// through a cp.async.bulk ring, in issue order: L0 d2 x2 d3 x3 m1c m1b m2.  TMEM columns of a group (C = first width):
// L0 [0,2C), C2d [2C,3C), C2x [3C,4C), E [0,2C), X1 [2C,4C), M1 [0,C), D [2C,4C).  x3 overwrites the columns of C2d / C2x
// (it waits for the complete drain of C2x) and m1c the first half of E (it waits for that half of the E*a drain); every
// other layer starts on its first block.
// The operand buffer holds the inputs of both stacks side by side (detector: chunks [0,C/8), descriptor: [C/8,2C/8));
// the E*a drain may only enter the descriptor half when x3 has read it.
//
// Drain instruction diet (as in level_fused.cu; the drains and the MMAs of a tile alternate, so every instruction of a drain
// is on the tile's dependency chain):
//   * the biases ride in the MMAs: d1 / x1 as the weight column of a constant-1 channel in the K padding of the grouped
//     input, d2 x2 d3 x3 m2 as ONE extra MMA per layer -- a resident block of ones (K columns 0 and 1) times the bias split
//     into bf16 hi (K column 0) + lo (K column 1), issued in front of the layer's K pieces; mlp1's bias enters the
//     per-keypoint mat-vec;
//   * ReLU is folded into the operand conversions (split_store8_relu: hi = cvt.rz.relu, lo = cvt.rn.relu of the residual);
//   * max_k relu(x) = relu(max_k x): the group maxima of X1 and of the descriptor are taken on the pre-activations.
#include "common.cuh"
#include "tc_common.cuh"
#include <math_constants.h>

// Phase profile (cycle stamps of CTA 0 / group 0: epilogue warp 0 and the MMA thread), compiled in with -DLW_PROF and read
// back by tools/level_probe.py through hrn_level_ws_prof -- how the per-tile dependency chain spends its time.
#ifdef LW_PROF
__device__ unsigned long long g_lw_prof[64];
#define LW_STAMP(slot)                                                                   \
    do {                                                                                 \
        if (prof_on) { const long long t__ = clock64(); atomicAdd(&g_lw_prof[slot], (unsigned long long)(t__ - prof_t)); prof_t = t__; } \
    } while (0)
#else
#define LW_STAMP(slot) do { } while (0)
#endif

#ifndef LW3_EW
#define LW3_EW 8         // epilogue warps of the level-3 kernel.  16 (four threads per row) measured SLOWER in-box: step 4.585 vs
                         // 4.558 ms -- the tile is bound by the MMA <-> drain hand-offs of its dependency chain, not by SIMT throughput
#endif

namespace {

constexpr int LTM = 128;
constexpr int LW_NL = 8;                                  // MMA layers per tile
constexpr int LW_MAXBLK = 8;                              // operand blocks (32 columns) per layer, at most

// EW_: epilogue warps per group, 8 or 16 -- NH = EW / 4 threads per row, each draining every NH-th 32-column block of an
// accumulator (a warp can only read its own TMEM lane quadrant, so more warps means more threads per row, not fewer rows)
template <int C_, int KNBR_, int CIN_, int NG_, int RING_, int EW_>
struct LwCfg {
    static constexpr int C = C_, CO = 2 * C_, KNBR = KNBR_, CIN = CIN_, NG = NG_, RING = RING_, EW = EW_, NH = EW_ / 4;
    static constexpr int KG = (CIN + 5 + 15) / 16 * 16;                 // grouped input [feat | rel xyz, |rel|, 1 | 0-pad]
    static constexpr int OPC = (CO > KG ? CO : KG) / 8;                 // 8-channel chunks of the operand buffer
    static constexpr int OP_PLANE = OPC * LTM * 16;
    static constexpr int OP_BYTES = 2 * OP_PLANE;
    static constexpr int SLOT = CO * 64;                                // widest K=16 weight piece (hi + lo)
    static constexpr int KPT = LTM / KNBR;                              // keypoints per tile
    static constexpr int KSEG = KNBR;                                   // lanes per keypoint inside a warp
    static constexpr int PER = 32 / KSEG;
    // layers in issue order: L0=[d1;x1]  d2  x2  d3  x3  m1c(E*a)  m1b(X1)  m2
    enum { L_0 = 0, L_D2 = 1, L_X2 = 2, L_D3 = 3, L_X3 = 4, L_M1C = 5, L_M1B = 6, L_M2 = 7 };
    __host__ __device__ static constexpr int lk(int l) { return l == 0 ? KG : (l == L_M1C || l == L_M1B) ? CO : C; }
    __host__ __device__ static constexpr int ln(int l) { return (l == 0 || l == L_D3 || l == L_X3 || l == L_M2) ? CO : C; }
    __host__ __device__ static constexpr int lacc(int l) {              // first TMEM column of the accumulator
        return l == L_D2 ? 2 * C : l == L_X2 ? 3 * C : (l == L_X3 || l == L_M2) ? 2 * C : 0;
    }
    // descriptor-stack layers read the second half of the operand buffer: first K=16 piece / first 32-column block
    __host__ __device__ static constexpr int lpiece0(int l) { return (l == L_X2 || l == L_X3) ? C / 16 : 0; }
    __host__ __device__ static constexpr int lblock0(int l) { return (l == L_X2 || l == L_X3) ? C / 32 : 0; }
    // layers that overwrite TMEM columns their own input is still being drained from: that many leading input blocks must
    // have left TMEM before the first MMA (x3 -> [2C,4C) over C2d | C2x: all of C2x; m1c -> M1 [0,C) over the first half
    // of E: the blocks of that half, the second half of the E*a drain then overlaps the MMAs)
    __host__ __device__ static constexpr int lprewait(int l) { return (l == L_X3 || l == L_M1C) ? C / 32 : 0; }
    // layers whose bias is one more MMA in front of their K pieces: ones x [b_hi, b_lo], a piece of ln * 32 bytes (hi plane only)
    __host__ __device__ static constexpr bool lbias(int l) { return l == L_D2 || l == L_X2 || l == L_D3 || l == L_X3 || l == L_M2; }
    __host__ __device__ static constexpr int woff(int l) {
        int o = 0;
        for (int i = 0; i < l; ++i) o += lk(i) * ln(i) * 4 + (lbias(i) ? ln(i) * 32 : 0);
        return o;
    }
    static constexpr int W_BYTES = woff(LW_NL);
    static constexpr int ONES_BYTES = 2 * LTM * 16;                     // resident A operand of the bias MMAs, shared by the groups
    // biases (floats): d1 d2 d3 x1 x2 x3 m1 m2 (only m1 is read by the kernel: the others ride in the MMAs)
    // at d1 0, d2 C, d3 2C, x1 4C, x2 5C, x3 6C, m1 8C, m2 9C
    static constexpr int B_M1 = 8 * C, B_COUNT = 11 * C;
    static constexpr int T_GROUP = 4 * C;
    static constexpr int T_COLS = NG * T_GROUP <= 256 ? 256 : 512;
    // shared memory of a group: operand buffer | weight ring | sCol [KPT][CO] | kpb [KPT][C] | sX [2][128]
    static constexpr int G_SCOL = OP_BYTES + RING * SLOT;
    static constexpr int G_KPB = G_SCOL + KPT * CO * 4;
    static constexpr int G_SX = G_KPB + KPT * C * 4;
    static constexpr int G_BYTES = G_SX + NH * LTM * 4;
    static constexpr int SMEM = NG * G_BYTES + ONES_BYTES;
    static constexpr int THREADS = NG * (EW + 2) * 32;
    static_assert(NG * T_GROUP <= 512, "TMEM");
    static_assert(KNBR == 16 || KNBR == 32, "group reductions are written for 16 or 32 neighbours (one warp holds whole groups)");
    static_assert(C % 64 == 0 && CO / 32 <= LW_MAXBLK, "two epilogue halves take alternate 32-column blocks");
    static_assert(KG == (CIN + 4 + 15) / 16 * 16, "the constant-1 channel must fit the K padding");
    static_assert(CIN % 64 == 0 && KG - CIN <= 32, "each thread of a row converts whole 32-channel blocks; the geometry block is one block");
    static_assert(SMEM + 512 <= 227 * 1024, "shared memory (dynamic + the static barriers)");
    static_assert((KPT * C) % (EW * 32) == 0 || (EW * 32) % (KPT * C) == 0, "mat-vec mapping");
    static_assert((EW == 8 || EW == 16) && (C / 32) % NH == 0 && (CIN / 32) % NH == 0, "every thread of a row takes whole 32-column blocks");
};

__device__ __forceinline__ uint32_t lw_idesc(int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(LTM >> 4) << 24);
}

// segmented transpose-reduce: lanes form groups of KSEG consecutive lanes; afterwards the lane at position p of its group
// holds in v[0 .. 32/KSEG) the reduction over the group of columns p*(32/KSEG) + i
template <int KSEG, bool IS_MAX>
__device__ __forceinline__ void lw_seg_reduce(float (&v)[32], int lane) {
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

template <int KSEG>
__device__ __forceinline__ float lw_seg_max(float v) {
#pragma unroll
    for (int o = KSEG / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <int KSEG>
__device__ __forceinline__ float lw_seg_sum(float v) {
#pragma unroll
    for (int o = KSEG / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
level_ws_kernel(const float* __restrict__ q, const float* __restrict__ xyz, const float* __restrict__ feat,
                const int32_t* __restrict__ idx, const uint8_t* __restrict__ Wpack, const float* __restrict__ WaT,
                const float* __restrict__ biases, float* __restrict__ out_xyz, float* __restrict__ out_af,
                float* __restrict__ out_desc, int M, int N, int n_tiles) {
    constexpr int C = Cfg::C, CO = Cfg::CO, KNBR = Cfg::KNBR, CIN = Cfg::CIN, NG = Cfg::NG, RING = Cfg::RING,
                  KSEG = Cfg::KSEG, PER = Cfg::PER, KPT = Cfg::KPT, OPC = Cfg::OPC;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_wfull[NG][RING], s_wempty[NG][RING], s_opb[NG][LW_MAXBLK], s_accf[NG][2];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp_all = tid >> 5, lane = tid & 31;
    // Warp roles: the epilogue warps of all groups first (a warp reads the TMEM lane quadrant warp_id % 4 of the CTA, so
    // every group's eight epilogue warps must start at a multiple of four), then per group one MMA and one weight warp.
    constexpr int EW = Cfg::EW, NH = Cfg::NH;
    const bool is_epi = warp_all < NG * EW;
    const int grp = is_epi ? warp_all / EW : (warp_all - NG * EW) / 2;
    const int warp = is_epi ? warp_all % EW : EW + ((warp_all - NG * EW) & 1);
    uint8_t* gsm = smem + (size_t)grp * Cfg::G_BYTES;
    uint4* sOnes = reinterpret_cast<uint4*>(smem + (size_t)NG * Cfg::G_BYTES);     // [2 chunks][128 rows][16 B]
    float* sCol = reinterpret_cast<float*>(gsm + Cfg::G_SCOL);
    float* sKpb = reinterpret_cast<float*>(gsm + Cfg::G_KPB);
    float* sX = reinterpret_cast<float*>(gsm + Cfg::G_SX);

    if (tid == 0) {
        for (int g = 0; g < NG; ++g) {
            for (int i = 0; i < RING; ++i) { mbar_init(smem_u32(&s_wfull[g][i]), 1); mbar_init(smem_u32(&s_wempty[g][i]), 1); }
            for (int i = 0; i < LW_MAXBLK; ++i) mbar_init(smem_u32(&s_opb[g][i]), 4);
            mbar_init(smem_u32(&s_accf[g][0]), 1);
            mbar_init(smem_u32(&s_accf[g][1]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp_all == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(Cfg::T_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // A operand of the bias MMAs: bf16 ones in K columns 0 and 1 of every row, zeros elsewhere
    for (int i = tid; i < 2 * LTM; i += Cfg::THREADS) sOnes[i] = i < LTM ? make_uint4(0x3f803f80u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_all = s_tmem;
    const uint32_t tmem = tmem_all + grp * Cfg::T_GROUP;
    const int vgrid = (int)gridDim.x * NG, vblock = (int)blockIdx.x * NG + grp;        // groups act as virtual CTAs
    const uint32_t op_a = smem_u32(gsm), ring_a = op_a + Cfg::OP_BYTES;
    const uint32_t wfull0 = smem_u32(&s_wfull[grp][0]), wempty0 = smem_u32(&s_wempty[grp][0]);
    const uint32_t opb0 = smem_u32(&s_opb[grp][0]), accf = smem_u32(&s_accf[grp][0]);

    if (warp < EW) {
        // ================= epilogue warps: gather, drains, attention, reductions ================================
        const int h = warp >> 2, wq = warp & 3;                  // which of the row's NH threads (every NH-th 32-column block), TMEM lane quadrant
        const int rt = wq * 32 + lane;                           // row inside the tile = TMEM lane
        const int et = warp * 32 + lane;                         // epilogue thread id inside the group, 0..255
        const uint32_t lane_base = ((uint32_t)(wq * 32) << 16);
        const int pos = lane % KSEG, seg = lane / KSEG;
        const int kp_local = wq * (32 / KSEG) + seg;             // keypoint of this row inside the tile
        uint4* op_hi = reinterpret_cast<uint4*>(gsm);
        uint4* op_lo = op_hi + OPC * LTM;
        // Layer l completes on barrier l & 1 (eight layers per tile: the assignment repeats).  With ONE barrier the two
        // stacks' layers, which no longer wait for each other, could complete two phases before the epilogue looks -- a
        // parity wait cannot see that; on either barrier a layer's MMAs depend on a drain that follows the wait for the
        // barrier's previous layer.
        uint32_t accph = 0, acc_l = 0;
        auto ebar = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(EW * 32) : "memory"); };
        auto wait_acc = [&]() {                                  // waits are made in layer order
            const uint32_t s_ = acc_l & 1u;
            mbar_wait(accf + 8 * s_, (accph >> s_) & 1u);
            accph ^= 1u << s_;
            ++acc_l;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        };
        auto publish = [&](int b) {                              // operand block b is in shared memory
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(opb0 + 8 * b);
        };
        auto store_block = [&](const float (&f)[32], int b) {    // 32 non-negative columns -> bf16 hi/lo operand chunks 4b .. 4b+3
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                const float x[8] = {f[ch * 8], f[ch * 8 + 1], f[ch * 8 + 2], f[ch * 8 + 3], f[ch * 8 + 4], f[ch * 8 + 5],
                                    f[ch * 8 + 6], f[ch * 8 + 7]};
                split_store8(x, op_hi + (4 * b + ch) * LTM + rt, op_lo + (4 * b + ch) * LTM + rt);
            }
        };
        auto store_block_relu = [&](const float (&f)[32], int b) {       // relu(f) -> operand chunks, ReLU inside the conversions
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
                split_store8_relu(f + ch * 8, op_hi + (4 * b + ch) * LTM + rt, op_lo + (4 * b + ch) * LTM + rt);
        };
        // accumulator [tcol, tcol + ncols) (bias included by the MMAs) -> relu -> operand
        auto drain_plain = [&](int tcol, int ncols, int b0 = 0) {                        // -> operand blocks b0 ..
            for (int b = h; b < ncols / 32; b += NH) {
                uint32_t v[32];
                float f[32];
                tmem_ld32(tmem + lane_base + tcol + 32 * b, v);
#pragma unroll
                for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(v[e]);
                store_block_relu(f, b0 + b);
                publish(b0 + b);
            }
        };

#ifdef LW_PROF
        const bool prof_on = blockIdx.x == 0 && grp == 0 && warp == 0 && lane == 0;
        long long prof_t = clock64();
#endif
        // Grouped input [feat[idx] | rel xyz, |rel|, 1 | 0] -> operand buffer (all NH threads of a row), software-pipelined across
        // tiles: the loads of tile t+1 are issued before the wait for tile t's last layer (mlp2) and converted right after it --
        // the operand buffer is free then, and the first layer's MMAs of tile t+1 run under tile t's descriptor epilogue
        // (L0 accumulates in TMEM [0,2C), the descriptor is read from [2C,4C)).
        constexpr int CH = CIN / 8 / NH;                                      // feature chunks of this thread
        float4 fv[2 * CH];                                                    // the whole half row in flight at once
        float gx = 0.f, gy = 0.f, gz = 0.f, gqx = 0.f, gqy = 0.f, gqz = 0.f;
        int n_next = vblock < n_tiles ? __ldg(idx + (long long)vblock * LTM + rt) : 0;
        auto gather_load = [&](int t) {
            const long long bm_ = ((long long)t * LTM + rt) / KNBR;
            const long long b_ = bm_ / M;
            const int n = n_next;                                             // loaded one tile ahead
            if (t + vgrid < n_tiles) n_next = __ldg(idx + (long long)(t + vgrid) * LTM + rt);
            const float* pp = xyz + (b_ * N + n) * 3;
            const float4* fr = reinterpret_cast<const float4*>(feat + (b_ * N + n) * CIN) + h * CH * 2;
#pragma unroll
            for (int i = 0; i < 2 * CH; ++i) fv[i] = __ldg(fr + i);
            gx = __ldg(pp); gy = __ldg(pp + 1); gz = __ldg(pp + 2);
            if (h == NH - 1) { const float* qq = q + bm_ * 3; gqx = __ldg(qq); gqy = __ldg(qq + 1); gqz = __ldg(qq + 2); }
        };
        auto gather_store = [&]() {
            // every 32-channel block is published on its own: the first layer's MMAs start on the first block while
            // the rest of the row is still being converted (feature blocks h*CH/4 .. by thread h of the row, the geometry block by the last one)
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const float x[8] = {fv[2 * c].x, fv[2 * c].y, fv[2 * c].z, fv[2 * c].w,
                                    fv[2 * c + 1].x, fv[2 * c + 1].y, fv[2 * c + 1].z, fv[2 * c + 1].w};
                split_store8(x, op_hi + (h * CH + c) * LTM + rt, op_lo + (h * CH + c) * LTM + rt);
                if ((c & 3) == 3) publish(h * (CH / 4) + (c >> 2));
            }
            if (h == NH - 1) {
                const float rx = gx - gqx, ry = gy - gqy, rz = gz - gqz;
                const float x[8] = {rx, ry, rz, sqrtf(rx * rx + ry * ry + rz * rz), 1.f, 0.f, 0.f, 0.f};   // 1: bias column of d1 / x1
                split_store8(x, op_hi + (CIN / 8) * LTM + rt, op_lo + (CIN / 8) * LTM + rt);
#pragma unroll
                for (int c = CIN / 8 + 1; c < Cfg::KG / 8; ++c) {         // K padding
                    op_hi[c * LTM + rt] = make_uint4(0, 0, 0, 0);
                    op_lo[c * LTM + rt] = make_uint4(0, 0, 0, 0);
                }
                publish(CIN / 32);
            }
        };
        if (vblock < n_tiles) { gather_load(vblock); gather_store(); }
        for (int tile = vblock; tile < n_tiles; tile += vgrid) {
            const long long r = (long long)tile * LTM + rt;
            const long long bm = r / KNBR;
            const float nx = gx, ny = gy, nz = gz;                            // this tile's neighbour (the attention phase needs it)
            // ---- first layers of both stacks, then the second ones as their results arrive ------------------------------
            wait_acc();                                                       // L0: C1d | C1x
            LW_STAMP(1);
            drain_plain(Cfg::lacc(0), C);                                     // -> d2
            drain_plain(Cfg::lacc(0) + C, C, C / 32);                         // -> x2 (descriptor half of the buffer)
            LW_STAMP(2);
            wait_acc();                                                       // d2
            LW_STAMP(3);
            drain_plain(Cfg::lacc(Cfg::L_D2), C);                             // -> d3
            wait_acc();                                                       // x2
            drain_plain(Cfg::lacc(Cfg::L_X2), C, C / 32);                     // -> x3, which runs under the attention phase
            LW_STAMP(4);
            wait_acc();                                                       // d3: E
            LW_STAMP(5);
            // attention: a = softmax_k(max_c E), keypoint = sum_k a * nn (layers.py:151-155)
            float x1 = 0.f;                                                   // post-ReLU values are >= 0
            for (int b = h; b < CO / 32; b += NH) {
                uint32_t v[32];
                tmem_ld32(tmem + lane_base + Cfg::lacc(Cfg::L_D3) + 32 * b, v);
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    x1 = fmaxf(x1, fmaxf(fmaxf(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1])),
                                         fmaxf(__uint_as_float(v[4 * e + 2]), __uint_as_float(v[4 * e + 3]))));
            }
            sX[h * LTM + rt] = x1;
            ebar();
#pragma unroll
            for (int o = 1; o < NH; ++o) x1 = fmaxf(x1, sX[((h + o) % NH) * LTM + rt]);
            const float gmax = lw_seg_max<KSEG>(x1);
            const float ex = expf(x1 - gmax);
            const float s0 = lw_seg_sum<KSEG>(ex);
            const float a = ex / s0;
            if (h == 0) {
                const float s1 = lw_seg_sum<KSEG>(ex * nx), s2 = lw_seg_sum<KSEG>(ex * ny), s3 = lw_seg_sum<KSEG>(ex * nz);
                if (pos < 3) out_xyz[bm * 3 + pos] = (pos == 0 ? s1 : (pos == 1 ? s2 : s3)) / s0;
            }
            LW_STAMP(6);                                                      // attention
            // attentive feature map E*a = first K-segment of mlp1; its column sums = the attentive feature (layers.py:157-159)
            bool x3_done = false;
            for (int b = h; b < CO / 32; b += NH) {
                uint32_t v[32];
                float f[32];
                if (b >= C / 32 && !x3_done) { wait_acc(); x3_done = true; }      // x3 has read the descriptor half of the buffer
                tmem_ld32(tmem + lane_base + Cfg::lacc(Cfg::L_D3) + 32 * b, v);
                const f32x2_t a2 = f2_pack(a, a);
#pragma unroll
                for (int e = 0; e < 32; e += 2) {                 // relu(x) * a: the column sums below need the clamped values
                    f2_unpack(f2_mul(f2_pack(fmaxf(__uint_as_float(v[e]), 0.f), fmaxf(__uint_as_float(v[e + 1]), 0.f)), a2), f[e], f[e + 1]);
                }
                store_block(f, b);
                publish(b);
                lw_seg_reduce<KSEG, false>(f, lane);
#pragma unroll
                for (int i = 0; i < PER; ++i) out_af[bm * CO + 32 * b + pos * PER + i] = f[i];
            }
            LW_STAMP(7);                                                      // E*a drain + attentive feature
            // ---- descriptor head ------------------------------------------------------------------------------------
            wait_acc();                                                       // m1c: the E*a segment has been consumed
            LW_STAMP(8);
            LW_STAMP(12);
            for (int b = h; b < CO / 32; b += NH) {
                uint32_t v[32];
                float f[32];
                tmem_ld32(tmem + lane_base + Cfg::lacc(Cfg::L_X3) + 32 * b, v);
#pragma unroll
                for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(v[e]);
                store_block_relu(f, b);
                publish(b);
                lw_seg_reduce<KSEG, true>(f, lane);                           // max over the group (layers.py:202), relu(max) = max(relu)
#pragma unroll
                for (int i = 0; i < PER; ++i) sCol[kp_local * CO + 32 * b + pos * PER + i] = fmaxf(f[i], 0.f);
            }
            LW_STAMP(13);                                                     // X1 drain + group max
            ebar();                                                           // sCol complete
            LW_STAMP(14);
            // per-keypoint bias  Wa . max_k(X1)  on the CUDA cores (fp32), while the tensor core runs the X1 segment
            {
                constexpr int TPK = EW * 32 / C;                    // thread sets along the keypoints
                constexpr int OPT = (KPT + TPK - 1) / TPK;                    // keypoints per thread
                const int j = et % C, k0 = (et / C) * OPT;
                // two partial sums per keypoint (even / odd input channels of every group of four), packed fp32x2 FMAs:
                // half the arithmetic instructions of the scalar chain
                f32x2_t acc[OPT];
#pragma unroll
                for (int o = 0; o < OPT; ++o) acc[o] = f2_pack(0.f, 0.f);
                if (k0 < KPT) {
                    const float bj = __ldg(biases + Cfg::B_M1 + j);               // mlp1's bias rides in the per-keypoint row
                    const float4* w4 = reinterpret_cast<const float4*>(WaT) + j;      // Wa4 [CO/4][C][4]: 16 B per lane, coalesced
#pragma unroll (EW == 16 ? 4 : 8)
                    for (int c = 0; c < CO; c += 8) {
                        const float4 wa = __ldg(w4 + (size_t)(c / 4) * C), wb = __ldg(w4 + (size_t)(c / 4 + 1) * C);
                        const f32x2_t wa0 = f2_pack(wa.x, wa.y), wa1 = f2_pack(wa.z, wa.w), wb0 = f2_pack(wb.x, wb.y), wb1 = f2_pack(wb.z, wb.w);
#pragma unroll
                        for (int o = 0; o < OPT; ++o) {
                            const float4 ma = *reinterpret_cast<const float4*>(sCol + (k0 + o) * CO + c);
                            const float4 mb = *reinterpret_cast<const float4*>(sCol + (k0 + o) * CO + c + 4);
                            acc[o] = f2_fma(wa1, f2_pack(ma.z, ma.w), f2_fma(wa0, f2_pack(ma.x, ma.y), acc[o]));
                            acc[o] = f2_fma(wb1, f2_pack(mb.z, mb.w), f2_fma(wb0, f2_pack(mb.x, mb.y), acc[o]));
                        }
                    }
#pragma unroll
                    for (int o = 0; o < OPT; ++o) {
                        float lo_, hi_;
                        f2_unpack(acc[o], lo_, hi_);
                        sKpb[(k0 + o) * C + j] = (lo_ + hi_) + bj;
                    }
                }
            }
            LW_STAMP(15);                                                     // mat-vec
            ebar();                                                           // sKpb complete
            wait_acc();                                                       // m1b: M1 = Wc.(E*a) + Wb.X1
            LW_STAMP(16);
            for (int b = h; b < C / 32; b += NH) {
                uint32_t v[32];
                float f[32];
                tmem_ld32(tmem + lane_base + Cfg::lacc(Cfg::L_M1B) + 32 * b, v);
#pragma unroll
                for (int e = 0; e < 32; e += 4) {
                    const float4 k4 = *reinterpret_cast<const float4*>(sKpb + kp_local * C + 32 * b + e);
                    f2_unpack(f2_add(f2_pack(__uint_as_float(v[e]), __uint_as_float(v[e + 1])), f2_pack(k4.x, k4.y)), f[e], f[e + 1]);
                    f2_unpack(f2_add(f2_pack(__uint_as_float(v[e + 2]), __uint_as_float(v[e + 3])), f2_pack(k4.z, k4.w)), f[e + 2], f[e + 3]);
                }
                store_block_relu(f, b);
                publish(b);
            }
            LW_STAMP(17);                                                     // M1 drain
            const bool more = tile + vgrid < n_tiles;
            if (more) gather_load(tile + vgrid);                              // in flight under the wait
            wait_acc();                                                       // mlp2: descriptor = max_k (layers.py:207-208)
            LW_STAMP(18);
            if (more) gather_store();                                         // every MMA of this tile is complete: the buffer is free
            LW_STAMP(0);                                                      // gather (of the next tile)
            for (int b = h; b < CO / 32; b += NH) {
                uint32_t v[32];
                float f[32];
                tmem_ld32(tmem + lane_base + Cfg::lacc(Cfg::L_M2) + 32 * b, v);
#pragma unroll
                for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(v[e]);
                lw_seg_reduce<KSEG, true>(f, lane);
#pragma unroll
                for (int i = 0; i < PER; ++i) out_desc[bm * CO + 32 * b + pos * PER + i] = fmaxf(f[i], 0.f);
            }
            LW_STAMP(19);                                                     // descriptor epilogue
            // The next tile's L0 (already running) leaves [2C,4C) alone; its d2 writes there and is issued once the first
            // C1d block is published -- by warps of ONE column half.  All eight warps must have read the descriptor by
            // then: a named barrier closes the tile.  sX / sCol / sKpb are rewritten only after the next tile's barriers.
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            ebar();
        }
    } else if (warp == EW) {
        // ================= MMA issue ==================================================================================
        if (lane == 0) {
            constexpr uint64_t DESC_FIXED = ((uint64_t)(128 >> 4) << 32) | (1ull << 46);          // SBO = 128 B
            const uint64_t a_desc0 = DESC_FIXED | ((uint64_t)((LTM * 16) >> 4) << 16);             // LBO = 2048 B
            constexpr uint32_t LO16 = Cfg::OP_PLANE >> 4, P16 = (2 * LTM * 16) >> 4, SLOT16 = Cfg::SLOT >> 4;
            uint32_t ws = 0, wpar = 0, opph = 0;
            const uint32_t ones16 = smem_u32(sOnes) >> 4;
#ifdef LW_PROF
            const bool prof_on = blockIdx.x == 0 && grp == 0;
            long long prof_t = clock64();
#endif
            for (int tile = vblock; tile < n_tiles; tile += vgrid) {
                LW_STAMP(32);
#pragma unroll
                for (int l = 0; l < LW_NL; ++l) {
                    const int K = Cfg::lk(l), Nn = Cfg::ln(l);
                    const uint32_t d = tmem + Cfg::lacc(l), idesc = lw_idesc(Nn);
                    const uint64_t w_desc0 = DESC_FIXED | ((uint64_t)Nn << 16);                     // LBO = N * 16 B
                    const uint32_t wlo16 = 2 * Nn;
                    const int B0 = Cfg::lblock0(l), P0 = Cfg::lpiece0(l);
                    const int PRE = Cfg::lprewait(l);
                    for (int b = 0; b < PRE; ++b) { mbar_wait(opb0 + 8 * (B0 + b), (opph >> (B0 + b)) & 1); opph ^= 1u << (B0 + b); }
                    // D = 1 * b_hi + 1 * b_lo in front of the layer's K pieces.  d3 / x3 / m2 issue it at once (it needs no
                    // operand block: it runs while the layer waits for its first one; their target columns are free by
                    // then, see the header).  d2 / x2 must see their first block: the next tile's L0 is issued under this
                    // tile's descriptor epilogue, d2 / x2 overwrite the columns that epilogue reads, and their first block
                    // is published behind the barrier that closes it.
                    auto bias_mma = [&]() {
                        mbar_wait(wfull0 + 8 * ws, wpar);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        umma_bf16(d, a_desc0 | ones16, w_desc0 | ((ring_a >> 4) + ws * SLOT16), idesc, 0u);
                        umma_commit(wempty0 + 8 * ws);
                        if (++ws == (uint32_t)RING) { ws = 0; wpar ^= 1; }
                    };
                    const bool late_bias = (l == Cfg::L_D2 || l == Cfg::L_X2);         // the layer loop is fully unrolled: a constant
                    if (Cfg::lbias(l) && !late_bias) bias_mma();
                    for (int p = 0; p < K / 16; ++p) {
                        if ((p >> 1) >= PRE && (p & 1) == 0) {
                            const int bi = B0 + (p >> 1);
                            mbar_wait(opb0 + 8 * bi, (opph >> bi) & 1); opph ^= 1u << bi;
                        }
                        LW_STAMP(33);                                         // waiting for operand blocks
                        if (p == 0 && late_bias) bias_mma();
                        mbar_wait(wfull0 + 8 * ws, wpar);
                        LW_STAMP(34);                                         // waiting for weights
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t a16 = (op_a >> 4) + (P0 + p) * P16, w16 = (ring_a >> 4) + ws * SLOT16;
                        const uint64_t ah = a_desc0 | a16, al = a_desc0 | (a16 + LO16);
                        const uint64_t wh = w_desc0 | w16, wl = w_desc0 | (w16 + wlo16);
                        umma_bf16(d, ah, wh, idesc, (p > 0 || l == Cfg::L_M1B || Cfg::lbias(l)) ? 1u : 0u);
                        umma_bf16(d, al, wh, idesc, 1u);
                        umma_bf16(d, ah, wl, idesc, 1u);
                        umma_commit(wempty0 + 8 * ws);
                        if (++ws == (uint32_t)RING) { ws = 0; wpar ^= 1; }
                        LW_STAMP(35);                                         // issuing
                    }
                    umma_commit(accf + 8 * (l & 1));
                }
            }
        }
    } else {
        // ================= weight stream ==============================================================================
        if (lane == 0) {
            uint32_t ws = 0, wpar = 0;
            for (int tile = vblock; tile < n_tiles; tile += vgrid) {
                const uint8_t* src = Wpack;
#pragma unroll
                for (int l = 0; l < LW_NL; ++l) {
                    // a layer's bias piece (hi plane only) goes first, then its K=16 pieces
                    for (int p = Cfg::lbias(l) ? -1 : 0; p < Cfg::lk(l) / 16; ++p) {
                        const uint32_t bytes = (uint32_t)Cfg::ln(l) * (p < 0 ? 32u : 64u);
                        mbar_wait(wempty0 + 8 * ws, wpar ^ 1);    // no sleep: the refill latency of a slot bounds the stream
                        mbar_expect_tx(wfull0 + 8 * ws, bytes);
                        bulk_g2s(ring_a + ws * Cfg::SLOT, src, bytes, wfull0 + 8 * ws);
                        src += bytes;
                        if (++ws == (uint32_t)RING) { ws = 0; wpar ^= 1; }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp_all == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_all), "n"(Cfg::T_COLS) : "memory");
}

using CfgW3 = LwCfg<128, 16, 128, 1, 5, LW3_EW>;       // detector_3 / desc_extractor_3 (models.py:16,24)
using CfgW2 = LwCfg<64, 32, 64, 2, 5, 8>;         // detector_2 / desc_extractor_2 (models.py:15,23)

template <class Cfg>
int launch_level_ws(const float* q, const float* xyz, const float* feat, const int32_t* idx, const void* Wpack,
                    const float* WaT, const float* biases, float* out_xyz, float* out_af, float* out_desc, int B, int M,
                    int N, cudaStream_t st) {
    const int n_tiles = (int)((long long)B * M * Cfg::KNBR / LTM);
    auto kern = level_ws_kernel<Cfg>;
    static hrn_once_per_device attr;
    if (attr.need()) HRN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int want = (n_tiles + Cfg::NG - 1) / Cfg::NG;
    const int grid = want < sms ? want : sms;
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(q, xyz, feat, idx, (const uint8_t*)Wpack, WaT, biases, out_xyz, out_af,
                                                out_desc, M, N, n_tiles);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}

}  // namespace

// Fused detector + descriptor of hierarchy level `level` (2 or 3), warp-specialised.  q [B*M,3] sampled keypoint
// coordinates, xyz [B,N,3], feat [B,N,CIN] channels-last previous-level attentive features, idx [B*M*k] int32 neighbour
// indices; Wpack = K=16 weight pieces of the 8 MMA layers in execution order, WaT [2C/4][C][4] fp32 = the max_k(X1) block
// of mlp1, input channels in groups of four, biases as in LwCfg (engine_tc.pack_level_ws).  Outputs per keypoint: out_xyz [B*M,3],
// out_af [B*M,2C], out_desc [B*M,2C].
HRN_API int hrn_level_ws(int level, const float* q, const float* xyz, const float* feat, const int32_t* idx,
                         const void* Wpack, const float* WaT, const float* biases, float* out_xyz, float* out_af,
                         float* out_desc, int B, int M, int N, int k, void* stream) {
    if (!q || !xyz || !feat || !idx || !Wpack || !WaT || !biases || !out_xyz || !out_af || !out_desc || B < 0 || M <= 0 || N <= 0)
        return HRN_ERR_BAD_ARG;
    if (((long long)B * M * k) % LTM != 0) return HRN_ERR_UNSUPPORTED;
    if (B == 0) return HRN_OK;
    if (level == 3 && k == CfgW3::KNBR)
        return launch_level_ws<CfgW3>(q, xyz, feat, idx, Wpack, WaT, biases, out_xyz, out_af, out_desc, B, M, N, (cudaStream_t)stream);
    if (level == 2 && k == CfgW2::KNBR)
        return launch_level_ws<CfgW2>(q, xyz, feat, idx, Wpack, WaT, biases, out_xyz, out_af, out_desc, B, M, N, (cudaStream_t)stream);
    return HRN_ERR_UNSUPPORTED;
}

#ifdef LW_PROF
HRN_API int hrn_level_ws_prof(unsigned long long* host64, int reset) {
    if (reset) { unsigned long long z[64] = {0}; return (int)cudaMemcpyToSymbol(g_lw_prof, z, sizeof(z)); }
    return (int)cudaMemcpyFromSymbol(host64, g_lw_prof, 64 * sizeof(unsigned long long));
}
#endif

HRN_API int hrn_level_ws_pack_bytes(int level) { return level == 3 ? CfgW3::W_BYTES : level == 2 ? CfgW2::W_BYTES : -1; }
HRN_API int hrn_level_ws_bias_count(int level) { return level == 3 ? CfgW3::B_COUNT : level == 2 ? CfgW2::B_COUNT : -1; }
