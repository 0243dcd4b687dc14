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
// Per-layer kernels write and re-read every [rows, C] intermediate (3 launches, 6 HBM passes); here a CTA owns 128
// rows, the first operand is gathered from the virtual rows into shared memory, each layer's result goes
// TMEM -> registers -> bias+ReLU -> bf16 hi/lo split -> back IN PLACE as the next layer's A operand, and weights
// stream from L2 in K=16 pieces through a 4-slot cp.async.bulk ring (3 pieces of look-ahead).  bf16x3 products,
// fp32 accumulation, same numerics as mlp_tc.cu.
#include "common.cuh"
#include "tc_common.cuh"
#include <math_constants.h>

namespace {

constexpr int CTM = 128;                 // rows per CTA
constexpr int OPC_MAX = 32;              // operand buffer: 32 chunks of 8 channels = 256 wide
constexpr int RING_MAX = 4;
constexpr int CH_SMEM_MAX = 2 * OPC_MAX * CTM * 16 + RING_MAX * 256 * 64 + 3 * 256 * 4 + 64;

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
    int opc;                 // chunks of the operand buffer = max(first pass, hidden widths / 8)
    int slot_bytes;          // weight ring slot = widest layer x 64 B (one K=16 piece)
    int tmem_cols;           // power of two >= widest layer
    int ring;                // weight ring slots (2..4)
    int chunks0;             // 8-wide K chunks of the virtual input (segments padded to 8, total padded to even)
    int mode;
    int kseg;                // rows per group (8, 16 or 32)
};

__device__ __forceinline__ uint32_t ch_idesc(int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(CTM >> 4) << 24);
}

struct RowSrcC {
    const float* p[4];
    float sc[4];
    int c0[5];
    int ch[4];
};

__device__ __forceinline__ void ch_load_chunk(const RowSrcC& rs, int cg, float4& v0, float4& v1, float& sc) {
    int sgi = 0;
#pragma unroll
    for (int q = 1; q < 4; ++q) if (cg >= rs.c0[q]) sgi = q;
    const float* p = rs.p[0]; int cs = rs.c0[0], chn = rs.ch[0]; sc = rs.sc[0];
#pragma unroll
    for (int q = 1; q < 4; ++q) if (sgi == q) { p = rs.p[q]; cs = rs.c0[q]; chn = rs.ch[q]; sc = rs.sc[q]; }
    const int ch0 = (cg - cs) << 3;
    const int nvalid = (p != nullptr && cg < rs.c0[4]) ? chn - ch0 : 0;
    v0 = make_float4(0.f, 0.f, 0.f, 0.f); v1 = v0;
    if (nvalid >= 4) v0 = __ldg(reinterpret_cast<const float4*>(p + ch0));
    if (nvalid >= 8) v1 = __ldg(reinterpret_cast<const float4*>(p + ch0) + 1);
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

template <int KSEG>
__global__ void __launch_bounds__(CTM) chain3_kernel(const ChainArgs A) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_bar[2 * RING_MAX + 1];    // [0..3] piece landed, [4..7] slot drained, [8] accumulator
    __shared__ uint32_t s_tmem;

    const int RING = A.ring;
    const uint32_t OP_PLANE = (uint32_t)A.opc * CTM * 16;
    const uint32_t SLOT_BYTES = (uint32_t)A.slot_bytes;
    uint8_t* sOp = smem;
    uint8_t* sRing = smem + 2 * OP_PLANE;
    float* sBias = reinterpret_cast<float*>(sRing + RING * SLOT_BYTES);
    uint4* op_hi = reinterpret_cast<uint4*>(sOp);
    uint4* op_lo = reinterpret_cast<uint4*>(sOp + OP_PLANE);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_acc = smem_u32(&s_bar[2 * RING_MAX]);
    const int nl = A.nl;
    const int n1 = A.n[0];
    const int n3 = A.n[nl - 1];                       // issued width of the last layer
    const int cout = A.cout;

    if (tid == 0) {
        for (int i = 0; i < 2 * RING_MAX + 1; ++i) mbar_init(smem_u32(&s_bar[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(A.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    const int nbias = A.n[0] + (nl == 3 ? A.n[1] : 0) + cout;
    for (int i = tid; i < nbias; i += CTM) sBias[i] = __ldg(A.bias + i);

    // ---- row bookkeeping -----------------------------------------------------------------------------------
    const long long r = (long long)blockIdx.x * CTM + tid;
    const bool rvalid = r < A.rows;
    RowSrcC rs;
    int run = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        rs.p[s] = nullptr; rs.sc[s] = 1.f; rs.ch[s] = 0; rs.c0[s] = 0x7fffffff;
        if (s < A.in.n_seg) {
            const hrn_seg_t sg = A.in.seg[s];
            rs.c0[s] = run;
            run += (sg.channels + 7) >> 3;
            rs.ch[s] = sg.channels;
            if (rvalid) {
                rs.p[s] = sg.ptr + hrn_src_row(A.in, sg.mode, r) * sg.ld + sg.col0;
                if (sg.row_scale) rs.sc[s] = __ldg(sg.row_scale + r);
            }
        }
    }
    rs.c0[4] = run;

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t lane_base = ((uint32_t)(warp * 32) << 16);
    const uint32_t op_hi_a = smem_u32(sOp), op_lo_a = op_hi_a + OP_PLANE;
    const uint32_t ring_a = smem_u32(sRing);

    // weight piece stream (thread 0 only): global piece counters
    uint32_t g_next = 0;        // next piece to request
    uint32_t g_use = 0;         // next piece to consume
    size_t w_off = 0;           // byte offset of piece g_next in A.W
    const uint32_t pieces[3] = {(uint32_t)(A.chunks0 / 2), (uint32_t)(A.n[0] / 16), nl == 3 ? (uint32_t)(A.n[1] / 16) : 0u};
    const uint32_t g_total = pieces[0] + pieces[1] + pieces[2];
    auto piece_bytes = [&](uint32_t g) -> uint32_t {
        const int N = g < pieces[0] ? A.n[0] : (g < pieces[0] + pieces[1] ? A.n[1] : A.n[2]);
        return (uint32_t)N * 64u;
    };
    auto prefetch = [&]() {     // keep up to RING-1 pieces in flight beyond the one being consumed
        while (g_next < g_total && g_next < g_use + RING) {
            const uint32_t slot = g_next % RING;
            if (g_next >= (uint32_t)RING) mbar_wait(smem_u32(&s_bar[RING_MAX + slot]), ((g_next / RING) - 1) & 1);
            const uint32_t bytes = piece_bytes(g_next);
            mbar_expect_tx(smem_u32(&s_bar[slot]), bytes);
            bulk_g2s(ring_a + slot * SLOT_BYTES, A.W + w_off, bytes, smem_u32(&s_bar[slot]));
            w_off += bytes;
            ++g_next;
        }
    };
    if (tid == 0) prefetch();

    uint32_t acc_phase = 0;
    // issue `np` K=16 pieces of a layer of width N, A operand chunks starting at chunk a0 of the operand buffer
    auto mma_pieces = [&](int np, int a0, int N, bool accumulate) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t idesc = ch_idesc(N);
            const uint32_t a_lbo = CTM * 16, w_lbo = (uint32_t)N * 16;
            for (int p = 0; p < np; ++p) {
                prefetch();
                const uint32_t slot = g_use % RING;
                mbar_wait(smem_u32(&s_bar[slot]), (g_use / RING) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t w_hi = ring_a + slot * SLOT_BYTES, w_lo = w_hi + 2 * w_lbo;
                const uint64_t ah = umma_desc(op_hi_a + (a0 + 2 * p) * a_lbo, a_lbo, 128);
                const uint64_t al = umma_desc(op_lo_a + (a0 + 2 * p) * a_lbo, a_lbo, 128);
                const uint64_t wh = umma_desc(w_hi, w_lbo, 128);
                const uint64_t wl = umma_desc(w_lo, w_lbo, 128);
                umma_bf16(tmem, ah, wh, idesc, (accumulate || p > 0) ? 1u : 0u);
                umma_bf16(tmem, al, wh, idesc, 1u);
                umma_bf16(tmem, ah, wl, idesc, 1u);
                umma_commit(smem_u32(&s_bar[RING_MAX + slot]));
                ++g_use;
            }
            umma_commit(bar_acc);
            prefetch();
        }
        mbar_wait(bar_acc, acc_phase);
        acc_phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    };

    // ---- layer 1: virtual rows -> operand buffer, in passes of <= 32 chunks --------------------------------
    for (int c_base = 0; c_base < A.chunks0; c_base += OPC_MAX) {
        const int nc = min(OPC_MAX, A.chunks0 - c_base);
        for (int c4 = 0; c4 < nc; c4 += 4) {
            float4 v[8];
            float sc[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) ch_load_chunk(rs, c_base + c4 + c, v[2 * c], v[2 * c + 1], sc[c]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (c4 + c < nc) {
                    const float s = sc[c];
                    const float x[8] = {v[2 * c].x * s, v[2 * c].y * s, v[2 * c].z * s, v[2 * c].w * s,
                                        v[2 * c + 1].x * s, v[2 * c + 1].y * s, v[2 * c + 1].z * s, v[2 * c + 1].w * s};
                    split_store8(x, op_hi + (c4 + c) * CTM + tid, op_lo + (c4 + c) * CTM + tid);
                }
            }
        }
        mma_pieces(nc / 2, 0, n1, c_base > 0);
    }
    // accumulator -> relu(x + b) -> operand (in place)
    auto epi_to_operand = [&](int N, const float* bb) {
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem + lane_base + c0, v);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                if (c0 + ch * 8 < N) {
                    float x[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) x[e] = fmaxf(__uint_as_float(v[ch * 8 + e]) + bb[c0 + ch * 8 + e], 0.f);
                    split_store8(x, op_hi + (c0 / 8 + ch) * CTM + tid, op_lo + (c0 / 8 + ch) * CTM + tid);
                }
            }
        }
    };
    epi_to_operand(n1, sBias);
    mma_pieces(n1 / 16, 0, A.n[1], false);
    if (nl == 3) {
        epi_to_operand(A.n[1], sBias + n1);
        mma_pieces(A.n[1] / 16, 0, n3, false);
    }

    // ---- final epilogue -----------------------------------------------------------------------------------------
    const float* b3 = sBias + n1 + (nl == 3 ? A.n[1] : 0);
    const int act = A.act;
    const int pos = lane % KSEG;                       // position inside the group
    const long long grp = r / KSEG;
    constexpr int PER = 32 / KSEG;                     // reduced columns per lane and 32-column chunk
    float a_w = 1.f;
    if (A.mode == EPI_ATTN) {
        float x1 = -CUDART_INF_F;
        for (int c0 = 0; c0 < cout; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem + lane_base + c0, v);
#pragma unroll
            for (int e = 0; e < 32; ++e)
                if (c0 + e < cout) x1 = fmaxf(x1, fmaxf(__uint_as_float(v[e]) + b3[c0 + e], 0.f));   // EPI_ATTN: ReLU only
        }
        float gm = x1;
#pragma unroll
        for (int o = KSEG / 2; o > 0; o >>= 1) gm = fmaxf(gm, __shfl_xor_sync(0xffffffffu, gm, o));
        const float ex = expf(x1 - gm);
        float sm = ex;
#pragma unroll
        for (int o = KSEG / 2; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
        a_w = ex / sm;
        if (rvalid && A.a) A.a[r] = a_w;
    }
    const bool vec_ok = A.Y && ((cout & 3) == 0);
    for (int c0 = 0; c0 < cout; c0 += 32) {
        uint32_t v[32];
        float f[32];
        tmem_ld32(tmem + lane_base + c0, v);
        if (act == HRN_ACT_RELU) {
#pragma unroll
            for (int e = 0; e < 32; ++e) f[e] = (c0 + e < cout) ? fmaxf(__uint_as_float(v[e]) + b3[c0 + e], 0.f) * a_w : 0.f;
        } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) f[e] = (c0 + e < cout) ? act_fn(__uint_as_float(v[e]) + b3[c0 + e], act) : 0.f;
        }
        if (rvalid && A.Y) {
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
            if (rvalid) {
#pragma unroll
                for (int i = 0; i < PER; ++i) {
                    const int c = c0 + pos * PER + i;
                    if (c < cout) A.G[grp * cout + c] = f[i];
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(A.tmem_cols) : "memory");
}

}  // namespace

// Two or three fused layers on a virtual rows matrix.  W: packed K=16 pieces of the layers in execution order
// (engine_tc.pack_chain), bias = b1|b2|b3 (last: `cout` entries), issued widths n1,n2,(n3) multiples of 16 and <= 256
// (the last one = cout padded to 16), hidden activations ReLU, last activation `act`.  mode / outputs as in the file
// header; `kseg` = rows per group (8, 16 or 32; ignored for mode 0); rows must be a multiple of 128.
HRN_API int hrn_chain_tc(const hrn_rows_t* in, const void* W, const float* bias, int nl, int n1, int n2, int n3, int cout,
                         int act, int chunks0, int mode, int kseg, float* Y, int ldy, float* G, float* a, long long rows,
                         void* stream) {
    if (!in || !W || !bias || rows < 0 || in->n_seg < 1 || in->n_seg > 4 || (nl != 2 && nl != 3)) return HRN_ERR_BAD_ARG;
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
    if (rows == 0) return HRN_OK;
    ChainArgs A;
    A.in = *in; A.W = (const uint8_t*)W; A.bias = bias; A.Y = Y; A.G = G; A.a = a; A.rows = rows; A.ldy = ldy;
    A.n[0] = n1; A.n[1] = n2; A.n[2] = nl == 3 ? n3 : 16; A.nl = nl; A.cout = cout; A.act = act;
    A.chunks0 = chunks0; A.mode = mode; A.kseg = kseg;
    int maxn = n1 > n2 ? n1 : n2;
    if (nl == 3 && n3 > maxn) maxn = n3;
    int opc = chunks0 < OPC_MAX ? chunks0 : OPC_MAX;
    if (n1 / 8 > opc) opc = n1 / 8;
    if (nl == 3 && n2 / 8 > opc) opc = n2 / 8;
    A.opc = opc;
    A.slot_bytes = maxn * 64;
    A.tmem_cols = 32;
    while (A.tmem_cols < maxn) A.tmem_cols <<= 1;
    // 4 weight slots normally; 2 when that lets two CTAs share an SM (<= 113 KB each) -- overlapping two tiles'
    // load / MMA / epilogue phases is worth more than look-ahead depth
    const int fixed = 2 * opc * CTM * 16 + 3 * 256 * 4 + 64;
    int ring = RING_MAX;
    if (fixed + RING_MAX * A.slot_bytes > 113 * 1024 && fixed + 2 * A.slot_bytes <= 113 * 1024 && A.tmem_cols <= 256) ring = 2;
    A.ring = ring;
    const int CH_SMEM = fixed + ring * A.slot_bytes;
    static bool attr_set = false;
    if (!attr_set) {
        HRN_CUDA(cudaFuncSetAttribute(chain3_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_MAX));
        HRN_CUDA(cudaFuncSetAttribute(chain3_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_MAX));
        HRN_CUDA(cudaFuncSetAttribute(chain3_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_MAX));
        attr_set = true;
    }
    const int grid = (int)(rows / CTM);
    cudaStream_t st = (cudaStream_t)stream;
    if (kseg == 8) chain3_kernel<8><<<grid, CTM, CH_SMEM, st>>>(A);
    else if (kseg == 16) chain3_kernel<16><<<grid, CTM, CH_SMEM, st>>>(A);
    else chain3_kernel<32><<<grid, CTM, CH_SMEM, st>>>(A);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}
