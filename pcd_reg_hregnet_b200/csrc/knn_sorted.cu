// Exact kNN in 3-D with spatial culling, for clouds of 128 .. 32768 points (level 1: 1024 queries x 16384 points, K = 64).
//
// Same contract as knn3_kernel (knn.cu): dist = fma(dz,dz, fma(dy,dy, dx*dx)), K smallest by (dist, index) -- the
// result is bit-identical to the brute-force kernel, only the work changes:
//   1. knn_sort_kernel   one CTA per cloud: Morton key of (x, y) on a fixed 0.25 m lattice, bitonic sort of
//                        (key, index) in 128 KB of shared memory, points re-written in that order as float4
//                        {x, y, z, original index}, and one axis-aligned box per 32 consecutive points.
//   2. knn3_sorted_kernel one warp per query: distance to every box of the cloud (16 boxes per lane, a true lower
//                        bound of the members' distances in fp32 because subtraction and fma round monotonically),
//                        seed the running top-K (an unsorted set with a tracked maximum) from the closest boxes,
//                        then visit only boxes whose bound does not exceed the current K-th distance (typically
//                        ~20 of 512 boxes are opened), and sort the K survivors once at the end.
#include "common.cuh"
#include "knn_select.cuh"
#include <math_constants.h>
#include <type_traits>
#include <stdlib.h>

namespace knn_sorted {

using namespace knn_sel;

__device__ __forceinline__ unsigned part1by1(unsigned v) {
    v &= 0x0000ffffu;
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}

constexpr int SORT_THREADS = 1024;

// pts_out [B, N2] float4 (x,y,z,idx bits; padding = +inf / 0x7fffffff), boxes [B, N2/32, 6] (min xyz, max xyz).
// WIDE (N2 <= 16384): 8-byte keys (20-bit Morton code of the 0.25 m lattice << 32 | index), 128 KB of shared memory.
// !WIDE (N2 = 32768): 4-byte keys (17-bit Morton code: 0.5 m cells in x, 1 m in y, << 15 | 15-bit index) -- the same
// 128 KB hold twice the points; coarser cells only make the boxes a little larger, the search stays exact.
template <bool WIDE>
__global__ void __launch_bounds__(SORT_THREADS)
knn_sort_kernel(const float* __restrict__ xyz, float4* __restrict__ pts_out, float* __restrict__ boxes, int N, int N2) {
    using Key = typename std::conditional<WIDE, unsigned long long, unsigned>::type;
    extern __shared__ __align__(8) unsigned char s_raw[];
    Key* s_key = reinterpret_cast<Key*>(s_raw);     // N2 entries
    const Key PAD = ~(Key)0;                        // padding sorts last
    const int b = blockIdx.x, tid = threadIdx.x;
    xyz += (size_t)b * N * 3;
    for (int i = tid; i < N2; i += SORT_THREADS) {
        Key k = PAD;
        if (i < N) {
            const float x = xyz[i * 3 + 0], y = xyz[i * 3 + 1];
            if (WIDE) {
                const int ix = min(1023, max(0, (int)floorf((x + 128.f) * 4.f)));
                const int iy = min(1023, max(0, (int)floorf((y + 128.f) * 4.f)));
                const unsigned m = part1by1((unsigned)ix) | (part1by1((unsigned)iy) << 1);
                k = (Key)(((unsigned long long)m << 32) | (unsigned)i);
            } else {
                const int ix = min(511, max(0, (int)floorf((x + 128.f) * 2.f)));
                const int iy = min(255, max(0, (int)floorf(y + 128.f)));
                const unsigned m = part1by1((unsigned)ix) | (part1by1((unsigned)iy) << 1);     // 17 bits
                k = (Key)((m << 15) | (unsigned)i);
            }
        }
        s_key[i] = k;
    }
    __syncthreads();
    for (int size = 2; size <= N2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < (N2 >> 1); t += SORT_THREADS) {
                const int lo = ((t / stride) * stride * 2) + (t % stride);
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const Key a = s_key[lo], c = s_key[hi];
                if ((a > c) == up) { s_key[lo] = c; s_key[hi] = a; }
            }
            __syncthreads();
        }
    }
    float4* po = pts_out + (size_t)b * N2;
    float* bo = boxes + (size_t)b * (N2 / 32) * 6;
    const int lane = tid & 31;
    for (int i = tid; i < N2; i += SORT_THREADS) {          // i / 32 is warp-uniform: one chunk per warp step
        const Key k = s_key[i];
        const bool valid = i < N;                           // the N2 - N padding keys are the largest: they sort last
        const int src = WIDE ? (int)((unsigned long long)k & 0xffffffffu) : (int)((unsigned)k & 0x7fffu);
        float x = CUDART_INF_F, y = CUDART_INF_F, z = CUDART_INF_F;
        if (valid) { x = xyz[src * 3 + 0]; y = xyz[src * 3 + 1]; z = xyz[src * 3 + 2]; }
        po[i] = make_float4(x, y, z, __int_as_float(valid ? src : 0x7fffffff));
        float mnx = x, mny = y, mnz = z;
        float mxx = valid ? x : -CUDART_INF_F, mxy = valid ? y : -CUDART_INF_F, mxz = valid ? z : -CUDART_INF_F;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
            mnz = fminf(mnz, __shfl_xor_sync(0xffffffffu, mnz, o)); mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
            mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o)); mxz = fmaxf(mxz, __shfl_xor_sync(0xffffffffu, mxz, o));
        }
        if (lane == 0) {
            float* o6 = bo + (i >> 5) * 6;
            o6[0] = mnx; o6[1] = mny; o6[2] = mnz; o6[3] = mxx; o6[4] = mxy; o6[5] = mxz;
        }
    }
}

// Radix variant of knn_sort_kernel for N2 <= 16384: 4-byte keys (18-bit Morton code of the 0.5 m lattice << 14 | index),
// two STABLE counting-sort passes of 9 bits each in shared memory.  A pass: every warp owns a contiguous range of
// positions and walks it 32 at a time; lanes with the same digit find each other with MATCH.ANY (rank inside the step =
// popc of the lower lanes, the first lane of a group carries the count); per-warp digit counts [32][512] are scanned in
// (digit, warp) order, which makes the scatter stable.  ~20 us per 16384-point cloud instead of ~200 us for the bitonic
// network (105 block-wide passes), which matters since the culled FPS no longer hides a long sort on the side stream.
// Same outputs as knn_sort_kernel; the order inside a Morton cell (ties) is by point index, as there.
constexpr int RDX_BITS = 9, RDX_DIGITS = 1 << RDX_BITS;
__global__ void __launch_bounds__(SORT_THREADS)
knn_sort_radix_kernel(const float* __restrict__ xyz, float4* __restrict__ pts_out, float* __restrict__ boxes, int N, int N2) {
    extern __shared__ __align__(8) unsigned char s_raw[];
    unsigned* s_a = reinterpret_cast<unsigned*>(s_raw);          // [N2]
    unsigned* s_b = s_a + N2;                                    // [N2]
    unsigned* s_h = s_b + N2;                                    // [32 warps][512 digits]
    __shared__ unsigned s_wsum[32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    xyz += (size_t)b * N * 3;
    for (int i = tid; i < N2; i += SORT_THREADS) {
        unsigned k = 0xffffffffu;                                // padding sorts last (stable: behind every real point)
        if (i < N) {
            const float x = xyz[i * 3 + 0], y = xyz[i * 3 + 1];
            const int ix = min(511, max(0, (int)floorf((x + 128.f) * 2.f)));
            const int iy = min(511, max(0, (int)floorf((y + 128.f) * 2.f)));
            k = ((part1by1((unsigned)ix) | (part1by1((unsigned)iy) << 1)) << 14) | (unsigned)i;
        }
        s_a[i] = k;
    }
    const int range = N2 / 32;                                   // positions per warp (a multiple of 32)
    unsigned* src = s_a;
    unsigned* dst = s_b;
    unsigned* hw = s_h + warp * RDX_DIGITS;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const int shift = 14 + pass * RDX_BITS;
        for (int d = lane; d < RDX_DIGITS; d += 32) hw[d] = 0u;
        __syncthreads();                                         // keys (first pass) / the previous scatter are complete
        for (int p0 = warp * range; p0 < (warp + 1) * range; p0 += 32) {
            const unsigned d = (src[p0 + lane] >> shift) & (RDX_DIGITS - 1);
            const unsigned m = __match_any_sync(0xffffffffu, d);
            if ((int)(__ffs(m) - 1) == lane) hw[d] += __popc(m);
            __syncwarp();
        }
        __syncthreads();
        {   // exclusive scan of the counts in (digit, warp) order: thread t = (digit t / 2, warps 16 (t & 1) .. + 16)
            const int d = tid >> 1, w0 = (tid & 1) * 16;
            unsigned v[16], run = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) { v[i] = s_h[(w0 + i) * RDX_DIGITS + d]; run += v[i]; }
            unsigned inc = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            if (lane == 31) s_wsum[warp] = inc;
            __syncthreads();
            unsigned base = 0;
            for (int w = 0; w < warp; ++w) base += s_wsum[w];
            unsigned off = base + inc - run;
#pragma unroll
            for (int i = 0; i < 16; ++i) { s_h[(w0 + i) * RDX_DIGITS + d] = off; off += v[i]; }
        }
        __syncthreads();
        for (int p0 = warp * range; p0 < (warp + 1) * range; p0 += 32) {
            const unsigned k = src[p0 + lane];
            const unsigned d = (k >> shift) & (RDX_DIGITS - 1);
            const unsigned m = __match_any_sync(0xffffffffu, d);
            const unsigned base = hw[d];
            dst[base + __popc(m & ((1u << lane) - 1u))] = k;
            __syncwarp();
            if ((int)(__ffs(m) - 1) == lane) hw[d] = base + __popc(m);
            __syncwarp();
        }
        unsigned* t = src; src = dst; dst = t;
    }
    __syncthreads();
    float4* po = pts_out + (size_t)b * N2;
    float* bo = boxes + (size_t)b * (N2 / 32) * 6;
    for (int i = tid; i < N2; i += SORT_THREADS) {          // i / 32 is warp-uniform: one chunk per warp step
        const unsigned k = src[i];
        const bool valid = i < N;                           // the N2 - N padding keys sort last
        const int sidx = (int)(k & 0x3fffu);
        float x = CUDART_INF_F, y = CUDART_INF_F, z = CUDART_INF_F;
        if (valid) { x = xyz[sidx * 3 + 0]; y = xyz[sidx * 3 + 1]; z = xyz[sidx * 3 + 2]; }
        po[i] = make_float4(x, y, z, __int_as_float(valid ? sidx : 0x7fffffff));
        float mnx = x, mny = y, mnz = z;
        float mxx = valid ? x : -CUDART_INF_F, mxy = valid ? y : -CUDART_INF_F, mxz = valid ? z : -CUDART_INF_F;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
            mnz = fminf(mnz, __shfl_xor_sync(0xffffffffu, mnz, o)); mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
            mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o)); mxz = fmaxf(mxz, __shfl_xor_sync(0xffffffffu, mxz, o));
        }
        if (lane == 0) {
            float* o6 = bo + (i >> 5) * 6;
            o6[0] = mnx; o6[1] = mny; o6[2] = mnz; o6[3] = mxx; o6[4] = mxy; o6[5] = mxz;
        }
    }
}

constexpr int QWARPS = 8;

// KPL: list registers per lane (K <= 32*KPL); BPL: boxes per lane (cloud <= 1024*BPL points)
template <int KPL, int BPL>
__global__ void __launch_bounds__(QWARPS * 32)
knn3_sorted_kernel(const float* __restrict__ p1, const int32_t* __restrict__ q_idx, const float* __restrict__ p2,
                   const float4* __restrict__ pts, const float* __restrict__ boxes, float* __restrict__ out_d,
                   int64_t* __restrict__ out_i64, int32_t* __restrict__ out_i32, float* __restrict__ out_nn,
                   float* __restrict__ out_q, int M, int N, int N2, int K) {
    extern __shared__ float s_box[];                        // [nbox][6], then (BPL >= 4) [QWARPS][BPL * 32] box bounds
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nbox = N2 >> 5;
    // Large clouds keep the per-query box bounds in shared memory, entry g * 32 + lane of the warp's row (a lane only ever
    // touches its own column): with them in BPL registers every loop over them is unrolled BPL times, the sweep with its
    // inlined insertion code included -- 3.5 K instructions at 16 boxes per lane, and 20 % of the level-1 search's stall
    // samples were instruction fetches (profiles/r02zz_ncu_knn3_search_l1.txt).  As run-time loops the body exists once,
    // and 16 registers per thread go back to the occupancy.
    constexpr bool SBD = BPL >= 4;
    float* sbd = s_box + nbox * 6 + warp * (BPL * 32) + lane;
    const float* bsrc = boxes + (size_t)b * nbox * 6;
    for (int i = threadIdx.x; i < nbox * 6; i += blockDim.x) s_box[i] = bsrc[i];
    __syncthreads();
    const int m = blockIdx.x * QWARPS + warp;
    if (m >= M) return;
    p2 += (size_t)b * N * 3;
    pts += (size_t)b * N2;
    const float* q = q_idx ? p2 + (size_t)q_idx[(size_t)b * M + m] * 3 : p1 + ((size_t)b * M + m) * 3;
    const float qx = q[0], qy = q[1], qz = q[2];
    if (out_q && lane < 3) out_q[((size_t)b * M + m) * 3 + lane] = q[lane];

    // lower bound of the squared distance to every box (same operation order as the point distance)
    auto box_bound = [&](int bx) -> float {
        float v = CUDART_NAN_F;                             // NaN = "never open": out of range or already visited
        if (bx < nbox) {
            const float* o = s_box + bx * 6;
            const float dx = fmaxf(fmaxf(o[0] - qx, qx - o[3]), 0.f);
            const float dy = fmaxf(fmaxf(o[1] - qy, qy - o[4]), 0.f);
            const float dz = fmaxf(fmaxf(o[2] - qz, qz - o[5]), 0.f);
            v = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
        }
        return v;
    };
    float bd[SBD ? 1 : BPL];
    if (SBD) {
#pragma unroll 1
        for (int g = 0; g < BPL; ++g) sbd[g * 32] = box_bound(g * 32 + lane);
    } else {
#pragma unroll
        for (int g = 0; g < BPL; ++g) bd[g] = box_bound(g * 32 + lane);
    }
    WarpSet<KPL> top;
    // pick the box with the smallest bound among the not yet visited ones (warp-uniform result, -1 = none left)
    auto closest_box = [&]() -> int {
        float best = CUDART_INF_F; int bg = 0;
        if (SBD) {
#pragma unroll 4
            for (int g = 0; g < BPL; ++g) { const float v = sbd[g * 32]; if (v < best) { best = v; bg = g; } }
        } else {
#pragma unroll
            for (int g = 0; g < BPL; ++g) if (bd[g] < best) { best = bd[g]; bg = g; }
        }
        const unsigned ob = hrn_ford(best);
        const unsigned wmin = __reduce_min_sync(0xffffffffu, ob);
        if (wmin == hrn_ford(CUDART_INF_F)) return -1;
        const int src = __ffs(__ballot_sync(0xffffffffu, ob == wmin)) - 1;
        const int g_sel = __shfl_sync(0xffffffffu, bg, src);
        if (lane == src) {
            if (SBD) sbd[g_sel * 32] = CUDART_NAN_F;        // visited
            else {
#pragma unroll
                for (int g = 0; g < BPL; ++g) if (g == g_sel) bd[g] = CUDART_NAN_F;
            }
        }
        return g_sel * 32 + src;
    };
    // candidate of this lane in box bx; padding (+inf) gets a distinct index above every real one
    auto box_dist = [&](int bx, float& dist, int& pid, int uniq) {
        const float4 p = __ldg(pts + bx * 32 + lane);
        const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
        dist = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
        pid = __float_as_int(p.w);
        if (pid == 0x7fffffff) pid -= uniq;
    };
    auto open_box = [&](int bx) {
        float dist; int pid;
        box_dist(bx, dist, pid, 0);                         // padding: inf is never below the threshold -> never offered
        top.offer(dist, pid);
    };
    // seed: the KPL closest boxes fill the set as they are (no sorting), then two more boxes tighten the threshold
    // before the sweep
    {
#pragma unroll
        for (int s = 0; s < KPL; ++s) {
            const int bx = closest_box();
            top.d[s] = CUDART_INF_F; top.i[s] = 0x7fffffff - (s * 32 + lane);           // "empty", distinct
            if (bx >= 0) box_dist(bx, top.d[s], top.i[s], s * 32 + lane);
            // the seed enters the set without a comparison: a NaN distance (NaN query or NaN point) would break the
            // ordering of the set and of the final sort; +inf keeps the member last and comparable
            if (!(top.d[s] == top.d[s])) top.d[s] = CUDART_INF_F;
        }
        if (K < 32 * KPL) {      // keep the K best of the seed, the other 32*KPL - K positions become dummies
            top.sort_set(lane);
#pragma unroll
            for (int s = 0; s < KPL; ++s)
                if (s * 32 + lane >= K) { top.d[s] = -1.f; top.i[s] = -1; }
        }
        top.refresh();
        for (int extra = 0; extra < 2; ++extra) {
            const int bx = closest_box();
            if (bx < 0) break;
            open_box(bx);
        }
    }
    // sweep: every box whose bound can still beat the K-th candidate
    auto sweep_chunk = [&](int g, float v) {
        unsigned mask = __ballot_sync(0xffffffffu, v <= top.thr_d);
        while (mask) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const float bnd = __shfl_sync(0xffffffffu, v, src);
            if (bnd <= top.thr_d) open_box(g * 32 + src);
        }
    };
    if (SBD) {
#pragma unroll 1
        for (int g = 0; g < BPL; ++g) {
            if (g * 32 >= nbox) break;
            sweep_chunk(g, sbd[g * 32]);
        }
    } else {
#pragma unroll
        for (int g = 0; g < BPL; ++g) {
            if (g * 32 >= nbox) break;
            sweep_chunk(g, bd[g]);
        }
    }
    top.sort_set(lane);                                     // dummies (-1, -1) sort first
    top.sanitize(N);
    const int ndummy = 32 * KPL - K;
    const size_t base = ((size_t)b * M + m) * K;
#pragma unroll
    for (int s = 0; s < KPL; ++s) {
        const int pos = s * 32 + lane - ndummy;
        if (pos >= 0) {
            const float d = top.d[s];
            const int i = top.i[s];
            if (out_d) out_d[base + pos] = d;
            if (out_i64) out_i64[base + pos] = (int64_t)i;
            if (out_i32) out_i32[base + pos] = i;
            if (out_nn) {
                const float* rr = p2 + (size_t)i * 3;
                float* o = out_nn + (base + pos) * 3;
                o[0] = rr[0]; o[1] = rr[1]; o[2] = rr[2];
            }
        }
    }
}

}  // namespace knn_sorted

// Scratch the caller provides for the culled search: pts [B*N2] float4 and boxes [B*(N2/32)*6] floats, N2 = next
// power of two >= N.  Returns HRN_ERR_UNSUPPORTED outside 128 <= N <= 32768 (callers then use hrn_knn).
// The two halves are also exported on their own: the sort depends on the reference cloud only, so a caller can run it
// on a second stream while the queries are still being chosen (FPS).
static int knn3_pow2(int N) {
    int N2 = 1024;
    while (N2 < N) N2 <<= 1;
    return N2;
}

HRN_API int hrn_knn3_sort(const float* p2, int B, int N, void* scratch_pts, float* scratch_boxes, void* stream) {
    using namespace knn_sorted;
    if (!p2 || !scratch_pts || !scratch_boxes || B < 0 || N <= 0) return HRN_ERR_BAD_ARG;
    if (N < 128 || N > 32768) return HRN_ERR_UNSUPPORTED;
    if (B == 0) return HRN_OK;
    const int N2 = knn3_pow2(N);
    static hrn_once_per_device attr_set;
    if (attr_set.need()) {
        HRN_CUDA(cudaFuncSetAttribute(knn_sort_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
        HRN_CUDA(cudaFuncSetAttribute(knn_sort_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 4));
    }
    static const bool bitonic = [] { const char* e = getenv("HRN_KNN_SORT"); return e && e[0] == 'b'; }();    // A/B switch
    if (N2 <= 16384 && !bitonic) {
        static hrn_once_per_device attr_rdx;
        if (attr_rdx.need())
            HRN_CUDA(cudaFuncSetAttribute(knn_sort_radix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8 + 32 * RDX_DIGITS * 4));
        knn_sort_radix_kernel<<<B, SORT_THREADS, (size_t)N2 * 8 + 32 * RDX_DIGITS * 4, (cudaStream_t)stream>>>(p2, (float4*)scratch_pts, scratch_boxes, N, N2);
    } else if (N2 <= 16384)
        knn_sort_kernel<true><<<B, SORT_THREADS, (size_t)N2 * 8, (cudaStream_t)stream>>>(p2, (float4*)scratch_pts, scratch_boxes, N, N2);
    else
        knn_sort_kernel<false><<<B, SORT_THREADS, (size_t)N2 * 4, (cudaStream_t)stream>>>(p2, (float4*)scratch_pts, scratch_boxes, N, N2);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}

HRN_API int hrn_knn3_search(const float* p1, const int32_t* q_idx, const float* p2, int B, int M, int N, int K,
                            const void* sorted_pts, const float* sorted_boxes, float* dists, int64_t* idx64,
                            int32_t* idx32, float* nn, float* q_out, void* stream) {
    using namespace knn_sorted;
    if (!p2 || (!p1 && !q_idx) || !sorted_pts || !sorted_boxes || B < 0 || M < 0 || N <= 0 || K <= 0) return HRN_ERR_BAD_ARG;
    if (K > N || K > 64 || N < 128 || N > 32768) return HRN_ERR_UNSUPPORTED;
    if (B == 0 || M == 0) return HRN_OK;
    const int N2 = knn3_pow2(N);
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(hrn_divup(M, QWARPS), B);
    const size_t bsm0 = (size_t)(N2 / 32) * 6 * sizeof(float);
    static hrn_once_per_device attr32;
    if (N2 > 16384 && attr32.need()) {                       // 24 KB of boxes + 32 KB of bounds
        HRN_CUDA(cudaFuncSetAttribute(knn3_sorted_kernel<1, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        HRN_CUDA(cudaFuncSetAttribute(knn3_sorted_kernel<2, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    }
#define HRN_KNN3_LAUNCH(KPL, BPL)                                                                                     \
    do {                                                                                                              \
        const size_t bsm = bsm0 + ((BPL) >= 4 ? (size_t)QWARPS * (BPL) * 32 * sizeof(float) : 0);                      \
        knn3_sorted_kernel<KPL, BPL><<<grid, QWARPS * 32, bsm, st>>>(p1, q_idx, p2, (const float4*)sorted_pts,          \
                                                                     sorted_boxes, dists, idx64, idx32, nn, q_out, M, N, N2, K); \
    } while (0)
    if (K <= 32) {
        if (N2 <= 1024) HRN_KNN3_LAUNCH(1, 1);
        else if (N2 <= 4096) HRN_KNN3_LAUNCH(1, 4);
        else if (N2 <= 16384) HRN_KNN3_LAUNCH(1, 16);
        else HRN_KNN3_LAUNCH(1, 32);
    } else {
        if (N2 <= 1024) HRN_KNN3_LAUNCH(2, 1);
        else if (N2 <= 4096) HRN_KNN3_LAUNCH(2, 4);
        else if (N2 <= 16384) HRN_KNN3_LAUNCH(2, 16);
        else HRN_KNN3_LAUNCH(2, 32);
    }
#undef HRN_KNN3_LAUNCH
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}

HRN_API int hrn_knn3_sorted(const float* p1, const int32_t* q_idx, const float* p2, int B, int M, int N, int K,
                            void* scratch_pts, float* scratch_boxes, float* dists, int64_t* idx64, int32_t* idx32,
                            float* nn, float* q_out, void* stream) {
    if (!p2 || (!p1 && !q_idx) || !scratch_pts || !scratch_boxes || B < 0 || M < 0 || N <= 0 || K <= 0) return HRN_ERR_BAD_ARG;
    if (K > N || K > 64 || N < 128 || N > 32768) return HRN_ERR_UNSUPPORTED;
    if (B == 0 || M == 0) return HRN_OK;
    const int rc = hrn_knn3_sort(p2, B, N, scratch_pts, scratch_boxes, stream);
    if (rc != HRN_OK) return rc;
    return hrn_knn3_search(p1, q_idx, p2, B, M, N, K, scratch_pts, scratch_boxes, dists, idx64, idx32, nn, q_out, stream);
}
