// Brute-force exact k-nearest-neighbours for sm_100a.
//
// Replaces pytorch3d.ops.knn_points (pytorch3d 0.7.8, third-party, not vendored in the reference); call sites
// reference models/HRegNet/layers.py:20 (xyz, K=64/32/16), :278 (256-d descriptors, K=8), :316,:322 (self-kNN,
// K=8), :434 (fine levels, K=8).
//
// Contract (stated in oracle/native_ops.c): dist = sum_d (q_d - r_d)^2 in fp32 accumulated
// sequentially d = 0..D-1 with fma;  K smallest in (dist asc, index asc) order;  squared distances returned.
//
// Kernel shape: one WARP per query, the 32 lanes sweep the reference points of the cloud in chunks of 32 from a
// shared-memory tile that the whole CTA (8 queries of the same cloud) streams through once; the running
// top-K is an unsorted set distributed over the lanes' registers with a REDUX-tracked maximum (knn_select.cuh), so a
// candidate test is one compare against the broadcast K-th entry + a ballot, and an insertion replaces the maximum --
// no shared or local memory per query, nothing spilled to global (pytorch3d keeps the K=64 running set of its
// one-thread-per-query kernel in global memory); the set is sorted once at the end.
#include "common.cuh"
#include "knn_select.cuh"
#include <math_constants.h>

namespace {

using namespace knn_sel;

constexpr int KNN_WARPS = 8;
constexpr int KNN3_TILE = 2048;  // reference points per shared-memory tile (24 KB, AoS: stride-3 is conflict-free)

// D == 3.  p2 [B,N,3]; queries either p1 [B,M,3] or p2 rows selected by q_idx [B,M] (fused FPS gather).
template <int KPL>
__global__ void __launch_bounds__(KNN_WARPS * 32)
knn3_kernel(const float* __restrict__ p1, const int32_t* __restrict__ q_idx, const float* __restrict__ p2,
            float* __restrict__ out_d, int64_t* __restrict__ out_i64, int32_t* __restrict__ out_i32,
            float* __restrict__ out_nn, float* __restrict__ out_q, int M, int N, int K) {
    __shared__ float s_ref[KNN3_TILE * 3];
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = blockIdx.x * KNN_WARPS + warp;
    const bool qvalid = m < M;
    p2 += (size_t)b * N * 3;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (qvalid) {
        const float* q = q_idx ? p2 + (size_t)q_idx[(size_t)b * M + m] * 3 : p1 + ((size_t)b * M + m) * 3;
        qx = q[0]; qy = q[1]; qz = q[2];
        if (out_q && lane < 3) out_q[((size_t)b * M + m) * 3 + lane] = q[lane];
    }
    WarpSet<KPL> top;
    top.init_empty(K, lane);
    for (int t0 = 0; t0 < N; t0 += KNN3_TILE) {
        const int tn = min(KNN3_TILE, N - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < tn * 3; i += blockDim.x) s_ref[i] = __ldg(p2 + (size_t)t0 * 3 + i);
        __syncthreads();
        if (qvalid) {
            for (int c = 0; c < tn; c += 32) {
                const int n = c + lane;
                float dist = CUDART_INF_F;
                if (n < tn) {
                    const float dx = qx - s_ref[n * 3 + 0], dy = qy - s_ref[n * 3 + 1], dz = qz - s_ref[n * 3 + 2];
                    dist = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
                }
                top.offer(dist, n < tn ? t0 + n : 0x7fffffff);
            }
        }
    }
    if (!qvalid) return;
    top.sort_set(lane);
    top.sanitize(N);
    const size_t base = ((size_t)b * M + m) * K;
    top.for_each_sorted(K, lane, [&](int pos, float d, int i) {
        if (out_d) out_d[base + pos] = d;
        if (out_i64) out_i64[base + pos] = (int64_t)i;
        if (out_i32) out_i32[base + pos] = i;
        if (out_nn) {
            const float* r = p2 + (size_t)i * 3;
            float* o = out_nn + (base + pos) * 3;
            o[0] = r[0]; o[1] = r[1]; o[2] = r[2];
        }
    });
}

// Generic D (descriptor space, D = 256 in CoarseReg).  Sequential fma chain over d per (query, ref) pair, so the
// distances are bit-identical to the oracle; register-tiled: a CTA owns 32 queries, a warp 4 of them, a lane 4
// references of the current 128-reference tile -> 16 independent accumulators per thread.  8 warps per CTA: with 4 warps
// x 8 queries the coarse-level search (32 pairs x 256 x 256 x 256) put fewer than two warps on a scheduler and ran at the
// dependent-issue latency of a single warp (175 us, ncu: 33 % issue).  References stream through shared memory in
// [128 refs] x [64 dims] slices, double-buffered with 16-byte cp.async (4-byte copies when D or the base address is not
// 16-byte aligned).  The accumulators are fp32x2 pairs over two QUERIES of the warp: the queries are stored
// pair-interleaved ([pair][d][2]), so one broadcast 16-byte load hands every lane two ready-made fp32x2 operands, a
// lane's reference values come four dims at a time (rows of 68 words: conflict-free for 16-byte loads) and enter
// FADD2 as the scalar-broadcast operand: a chunk of 4 dims costs 4 + 4 LDS.128 for 64 packed instructions (the scalar
// version: 32 LDS, each right in front of its first use, and 4-byte cp.async only).  (r - q)^2 == (q - r)^2 exactly.
// Dims beyond D are zero on both sides (padded query rows, zero-filled slices): fma(0, 0, acc) = acc, so whole chunks
// of 4 stay bit-exact.
constexpr int KD_Q = 32;         // queries per CTA
constexpr int KD_R = 128;        // references per tile
constexpr int KD_D = 64;         // dims per slice
constexpr int KD_WARPS = 8;      // warps per CTA
constexpr int KD_QW = KD_Q / KD_WARPS;   // queries per warp
constexpr int KD_LD = KD_D + 4;          // words per reference row of a slice
constexpr int KD_RBUF = KD_R * KD_LD;    // words per slice buffer
static_assert(KD_QW == 4, "two query pairs per warp");
template <int KPL>
__global__ void __launch_bounds__(KD_WARPS * 32)
knnd_kernel(const float* __restrict__ p1, const float* __restrict__ p2, float* __restrict__ out_d,
            int64_t* __restrict__ out_i64, int32_t* __restrict__ out_i32, float* __restrict__ out_nn, int M, int N,
            int D, int K, int vec16) {
    extern __shared__ __align__(16) float smem[];
    const int Dp = (D + 3) & ~3;                // query rows padded to whole chunks
    float* s_q = smem;                          // [KD_Q / 2 pairs][Dp][2]
    float* s_r = smem + KD_Q * Dp;              // 2 x [KD_R][KD_LD]
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * KD_Q;
    p1 += (size_t)b * M * D;
    p2 += (size_t)b * N * D;
    WarpSet<KPL> top[KD_QW];
#pragma unroll
    for (int q = 0; q < KD_QW; ++q) top[q].init_empty(K, lane);
    // Small searches (K <= 16 of N <= 512 references, the CoarseReg shape: 8 of 256): the running top-K of a query is
    // kept SORTED (rank r in lane 32 - K + r) and merged with a tile's 128 distances by K rounds of "extract the
    // minimum" (two REDUX + a rescan of the winner lane's five candidates per round) instead of one replace-the-maximum
    // per candidate below the threshold: K (1 + ln(N / K)) ~ 34 replacements of ~60 instructions per query were 60 % of
    // the kernel's instructions at that shape (ncu).  Same total order (distance bits, index), same "empty" members
    // (+inf, 0x7fffffff - pos), NaN distances never selected -> the same set in the same order as the WarpSet path.
    constexpr int XS_TAKEN = 0x7fffffff;
    const bool xsel = (KPL == 1) && K <= 16 && N <= 512;
    int car_k[KD_QW], car_i[KD_QW];
#pragma unroll
    for (int q = 0; q < KD_QW; ++q) {
        car_k[q] = lane >= 32 - K ? 0x7f800000 : XS_TAKEN;
        car_i[q] = lane >= 32 - K ? 0x7fffffff - (31 - lane) : 0x7fffffff;
    }
    // reference slices [128 refs] x [64 dims]: slice s+1 is in flight while slice s is consumed; missing references /
    // dims are zero-filled
    const int n_dslice = (D + KD_D - 1) / KD_D;
    const int n_slice = ((N + KD_R - 1) / KD_R) * n_dslice;
    auto fill = [&](int sl, int buf) {
        const int t0s = (sl / n_dslice) * KD_R, d0s = (sl % n_dslice) * KD_D;
        float* dst = s_r + buf * KD_RBUF;
        if (vec16) {
            for (int i = threadIdx.x; i < KD_R * (KD_D / 4); i += blockDim.x) {
                const int rr = i / (KD_D / 4), dd = (i - rr * (KD_D / 4)) * 4;
                const int nb = (t0s + rr < N) ? max(0, min(4, D - d0s - dd)) * 4 : 0;
                const float* src = nb ? p2 + (size_t)(t0s + rr) * D + d0s + dd : p2;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + rr * KD_LD + dd)), "l"(src), "r"(nb) : "memory");
            }
        } else {
            for (int i = threadIdx.x; i < KD_R * KD_D; i += blockDim.x) {
                const int rr = i / KD_D, dd = i - rr * KD_D;
                const bool ok = t0s + rr < N && d0s + dd < D;
                const float* src = ok ? p2 + (size_t)(t0s + rr) * D + d0s + dd : p2;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + rr * KD_LD + dd)), "l"(src), "r"(ok ? 4 : 0) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    fill(0, 0);
    // queries of the CTA, pair-interleaved; 16-byte loads, all of a thread's loads in flight before its stores
    if (vec16 && ((uintptr_t)p1 & 15) == 0) {
        const int dq = Dp >> 2, nq4 = KD_Q * dq;
        for (int base = 0; base < nq4; base += 8 * KD_WARPS * 32) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = base + u * KD_WARPS * 32 + threadIdx.x, q = i / dq;
                v[u] = (i < nq4 && m0 + q < M) ? __ldg(reinterpret_cast<const float4*>(p1 + (size_t)(m0 + q) * D) + (i - q * dq))
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = base + u * KD_WARPS * 32 + threadIdx.x, q = i / dq, d = (i - q * dq) * 4;
                if (i < nq4) {
                    float* o = s_q + ((q >> 1) * Dp + d) * 2 + (q & 1);
                    o[0] = v[u].x; o[2] = v[u].y; o[4] = v[u].z; o[6] = v[u].w;
                }
            }
        }
    } else {
        for (int i = threadIdx.x; i < KD_Q * Dp; i += blockDim.x) {
            const int q = i / Dp, d = i - q * Dp;
            s_q[((q >> 1) * Dp + d) * 2 + (q & 1)] = (m0 + q < M && d < D) ? p1[(size_t)(m0 + q) * D + d] : 0.f;
        }
    }
    int sl = 0;
    for (int t0 = 0; t0 < N; t0 += KD_R) {
        // acc[qp][j]: queries (2 qp, 2 qp + 1) of the warp x reference lane + 32 j, packed as fp32x2: FADD2 / FFMA2
        // round each half like the scalar instructions, so the sequential chain over d stays bit-identical to the oracle
        f32x2_t acc[2][4];
#pragma unroll
        for (int qp = 0; qp < 2; ++qp)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[qp][j] = f2_pack(0.f, 0.f);
        for (int d0 = 0; d0 < D; d0 += KD_D, ++sl) {
            const int n4 = (min(KD_D, D - d0) + 3) >> 2;       // chunks of 4 dims in this slice
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();                                   // slice sl visible; everybody is done with slice sl-1
            if (sl + 1 < n_slice) fill(sl + 1, (sl + 1) & 1);
            const float* sr = s_r + (sl & 1) * KD_RBUF + lane * KD_LD;
            const float* qb = s_q + ((warp * 2) * Dp + d0) * 2;
            float4 R[4], Qv[2][2];
#pragma unroll
            for (int j = 0; j < 4; ++j) R[j] = *reinterpret_cast<const float4*>(sr + 32 * j * KD_LD);
#pragma unroll
            for (int qp = 0; qp < 2; ++qp)
#pragma unroll
                for (int h = 0; h < 2; ++h) Qv[qp][h] = *reinterpret_cast<const float4*>(qb + qp * Dp * 2 + 4 * h);
#pragma unroll 2
            for (int c = 0; c < n4; ++c) {
                const int cn = min(c + 1, n4 - 1);             // next chunk's operands in flight under this chunk's math
                float4 nR[4], nQ[2][2];
#pragma unroll
                for (int j = 0; j < 4; ++j) nR[j] = *reinterpret_cast<const float4*>(sr + 32 * j * KD_LD + 4 * cn);
#pragma unroll
                for (int qp = 0; qp < 2; ++qp)
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        nQ[qp][h] = *reinterpret_cast<const float4*>(qb + qp * Dp * 2 + 8 * cn + 4 * h);
#pragma unroll
                for (int dd = 0; dd < 4; ++dd) {               // dims 4c .. 4c+3, in order
#pragma unroll
                    for (int qp = 0; qp < 2; ++qp) {
                        const float4 qq = Qv[qp][dd >> 1];
                        const f32x2_t q2 = (dd & 1) ? f2_pack(qq.z, qq.w) : f2_pack(qq.x, qq.y);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float rv = dd == 0 ? R[j].x : dd == 1 ? R[j].y : dd == 2 ? R[j].z : R[j].w;
                            const f32x2_t e = f2_sub(f2_pack(rv, rv), q2);
                            acc[qp][j] = f2_fma(e, e, acc[qp][j]);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) R[j] = nR[j];
#pragma unroll
                for (int qp = 0; qp < 2; ++qp)
#pragma unroll
                    for (int h = 0; h < 2; ++h) Qv[qp][h] = nQ[qp][h];
            }
        }
        float a[KD_QW][4];
#pragma unroll
        for (int qp = 0; qp < 2; ++qp)
#pragma unroll
            for (int j = 0; j < 4; ++j) f2_unpack(acc[qp][j], a[2 * qp][j], a[2 * qp + 1][j]);
        if (xsel) {
            int ck[KD_QW][4], bk[KD_QW], bi[KD_QW], nk[KD_QW], ni[KD_QW];
#pragma unroll
            for (int q = 0; q < KD_QW; ++q) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    ck[q][j] = (t0 + lane + 32 * j < N && a[q][j] == a[q][j]) ? __float_as_int(a[q][j]) : XS_TAKEN;
                nk[q] = XS_TAKEN; ni[q] = 0x7fffffff;
            }
            auto local_best = [&](int q) {
                int k = car_k[q], i = car_i[q];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = t0 + lane + 32 * j;
                    if (ck[q][j] < k || (ck[q][j] == k && n < i)) { k = ck[q][j]; i = n; }
                }
                bk[q] = k; bi[q] = i;
            };
#pragma unroll
            for (int q = 0; q < KD_QW; ++q) local_best(q);
            for (int r = 0; r < K; ++r) {
#pragma unroll
                for (int q = 0; q < KD_QW; ++q) {
                    const int mk = __reduce_min_sync(0xffffffffu, bk[q]);
                    const int wi = __reduce_min_sync(0xffffffffu, bk[q] == mk ? bi[q] : 0x7fffffff);
                    if (lane == 32 - K + r) { nk[q] = mk; ni[q] = wi; }
                    if (bk[q] == mk && bi[q] == wi) {          // the winner's lane: remove it, rescan
                        if (car_k[q] == mk && car_i[q] == wi) car_k[q] = XS_TAKEN;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (ck[q][j] == mk && t0 + lane + 32 * j == wi) ck[q][j] = XS_TAKEN;
                        local_best(q);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < KD_QW; ++q) { car_k[q] = nk[q]; car_i[q] = ni[q]; }
        } else {
#pragma unroll
            for (int q = 0; q < KD_QW; ++q) {
                if (m0 + warp * KD_QW + q < M) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int n = t0 + lane + 32 * j;
                        top[q].offer(n < N ? a[q][j] : CUDART_INF_F, n < N ? n : 0x7fffffff);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < KD_QW; ++q) {
        const int m = m0 + warp * KD_QW + q;
        if (m >= M) continue;
        const size_t base = ((size_t)b * M + m) * K;
        if (xsel) {                              // already sorted: rank r in lane 32 - K + r, dummies in front
            top[q].d[0] = lane >= 32 - K ? __int_as_float(car_k[q]) : -1.f;
            top[q].i[0] = lane >= 32 - K ? car_i[q] : -1;
        } else {
            top[q].sort_set(lane);
        }
        top[q].sanitize(N);
        top[q].for_each_sorted(K, lane, [&](int pos, float d, int i) {
            if (out_d) out_d[base + pos] = d;
            if (out_i64) out_i64[base + pos] = (int64_t)i;
            if (out_i32) out_i32[base + pos] = i;
        });
        if (out_nn) {
            __syncwarp();
            for (int pos = 0; pos < K; ++pos) {
                const int sp = pos + 32 * KPL - K;              // sorted position (dummies first)
                const int s = sp >> 5, l = sp & 31;
                int iv = top[q].i[0];
#pragma unroll
                for (int t = 1; t < KPL; ++t) if (s == t) iv = top[q].i[t];
                const int src = __shfl_sync(0xffffffffu, iv, l);
                for (int d = lane; d < D; d += 32) out_nn[(base + pos) * D + d] = __ldg(p2 + (size_t)src * D + d);
            }
        }
    }
}

}  // namespace

// p1 [B,M,D] (ignored when q_idx != NULL: queries are p2[b, q_idx[b,m], :], D must be 3), p2 [B,N,D].
// Outputs (each nullable): dists [B,M,K] f32 squared ascending; idx64 / idx32 [B,M,K]; nn [B,M,K,D];
// q_out [B,M,3] (the gathered queries, only with q_idx).
HRN_API int hrn_knn(const float* p1, const int32_t* q_idx, const float* p2, int B, int M, int N, int D, int K,
                    float* dists, int64_t* idx64, int32_t* idx32, float* nn, float* q_out, void* stream) {
    if (!p2 || (!p1 && !q_idx) || B < 0 || M < 0 || N <= 0 || D <= 0 || K <= 0) return HRN_ERR_BAD_ARG;
    if (K > N || K > 64) return HRN_ERR_UNSUPPORTED;
    if (q_idx && D != 3) return HRN_ERR_BAD_ARG;
    if (B == 0 || M == 0) return HRN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(hrn_divup(M, KNN_WARPS), B);
    if (D == 3) {
        if (K <= 32) knn3_kernel<1><<<grid, KNN_WARPS * 32, 0, st>>>(p1, q_idx, p2, dists, idx64, idx32, nn, q_out, M, N, K);
        else knn3_kernel<2><<<grid, KNN_WARPS * 32, 0, st>>>(p1, q_idx, p2, dists, idx64, idx32, nn, q_out, M, N, K);
    } else {
        const size_t smem = ((size_t)KD_Q * ((D + 3) & ~3) + 2 * (size_t)KD_RBUF) * sizeof(float);
        const int vec16 = (D % 4 == 0) && ((uintptr_t)p2 % 16 == 0);
        if (smem > 200 * 1024) return HRN_ERR_UNSUPPORTED;
        dim3 gridd(hrn_divup(M, KD_Q), B);
        if (K <= 32) {
            if (smem > 48 * 1024) HRN_CUDA(cudaFuncSetAttribute(knnd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            knnd_kernel<1><<<gridd, KD_WARPS * 32, smem, st>>>(p1, p2, dists, idx64, idx32, nn, M, N, D, K, vec16);
        } else {
            if (smem > 48 * 1024) HRN_CUDA(cudaFuncSetAttribute(knnd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            knnd_kernel<2><<<gridd, KD_WARPS * 32, smem, st>>>(p1, p2, dists, idx64, idx32, nn, M, N, D, K, vec16);
        }
    }
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}
