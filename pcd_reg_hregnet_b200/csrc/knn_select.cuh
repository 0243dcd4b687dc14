// Warp-level exact top-K selection shared by the kNN kernels (knn.cu, knn_sorted.cu).
//
// The K best candidates so far live as an UNSORTED set spread over the warp (slot s of lane l), together with its
// largest member (thr_d, thr_i) in (dist, index) order.  A better candidate replaces the largest member and the
// maximum is found again with two REDUX -- ~16 instructions per insertion instead of ~40 for a list kept sorted across
// lanes with shuffles; the set is sorted once, at the end, with a bitonic network.  The kNN kernels are bound by the
// instruction count of these insertions.
// Members are unique: real points by their index, +inf padding / "empty" entries get distinct indices above every real
// one, and the positions beyond K (K < 32*KPL) hold dummies (-1, -1) that can never be the maximum (distances are
// >= 0; the REDUX runs on the float bits as SIGNED integers, which orders {-1} < [0, +inf]).
#pragma once
#include <math_constants.h>

namespace knn_sel {

__device__ __forceinline__ bool cand_less(float da, int ia, float db, int ib) {
    return da < db || (da == db && ia < ib);
}

// 32-lane bitonic networks on (dist, index) pairs
__device__ __forceinline__ void cmpx(float& d, int& i, int j, bool keep_min) {
    const float od = __shfl_xor_sync(0xffffffffu, d, j);
    const int oi = __shfl_xor_sync(0xffffffffu, i, j);
    const bool self_less = cand_less(d, i, od, oi);
    const bool take_self = (self_less == keep_min);
    d = take_self ? d : od;
    i = take_self ? i : oi;
}
__device__ __forceinline__ void bitonic_sort32(float& d, int& i, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) cmpx(d, i, j, ((lane & j) == 0) == ((lane & k) == 0));
}
__device__ __forceinline__ void bitonic_merge32(float& d, int& i, int lane) {   // bitonic in -> ascending out
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) cmpx(d, i, j, (lane & j) == 0);
}

template <int KPL>
struct WarpSet {
    float d[KPL];
    int i[KPL];
    float thr_d;
    int thr_i;
    __device__ __forceinline__ void refresh() {
        float ld = d[0]; int li = i[0];
#pragma unroll
        for (int s = 1; s < KPL; ++s)
            if (d[s] > ld || (d[s] == ld && i[s] > li)) { ld = d[s]; li = i[s]; }
        const int md = __reduce_max_sync(0xffffffffu, __float_as_int(ld));
        thr_i = __reduce_max_sync(0xffffffffu, (__float_as_int(ld) == md) ? li : (int)0x80000000);
        thr_d = __int_as_float(md);
    }
    // K "empty" members (+inf with distinct indices above every real one), dummies beyond K
    __device__ __forceinline__ void init_empty(int K, int lane) {
#pragma unroll
        for (int s = 0; s < KPL; ++s) {
            const int pos = s * 32 + lane;
            d[s] = pos < K ? CUDART_INF_F : -1.f;
            i[s] = pos < K ? 0x7fffffff - pos : -1;
        }
        refresh();
    }
    __device__ __forceinline__ void replace_max(float xd, int xi) {
#pragma unroll
        for (int s = 0; s < KPL; ++s)
            if (d[s] == thr_d && i[s] == thr_i) { d[s] = xd; i[s] = xi; }
        refresh();
    }
    // per-lane candidate (cd, ci); +inf with an index >= 0x7fffffff - 63 is never below the threshold of a set that
    // still holds "empty" members only if its index is larger -- callers offer padding as (+inf, 0x7fffffff)
    __device__ __forceinline__ void offer(float cd, int ci) {
        unsigned mask = __ballot_sync(0xffffffffu, cand_less(cd, ci, thr_d, thr_i));
        while (mask) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const float xd = __shfl_sync(0xffffffffu, cd, src);
            const int xi = __shfl_sync(0xffffffffu, ci, src);
            if (cand_less(xd, xi, thr_d, thr_i)) replace_max(xd, xi);
        }
    }
    // ascending by (dist, index): rank r ends up in slot r / 32 of lane r % 32 (dummies first)
    __device__ __forceinline__ void sort_set(int lane) {
#pragma unroll
        for (int s = 0; s < KPL; ++s) bitonic_sort32(d[s], i[s], lane);
        static_assert(KPL == 1 || KPL == 2, "merge written for one or two registers per lane");
        if (KPL == 2) {          // merge the two sorted halves into one ascending list of 64
            const float rd = __shfl_sync(0xffffffffu, d[KPL - 1], 31 - lane);
            const int ri = __shfl_sync(0xffffffffu, i[KPL - 1], 31 - lane);
            const bool lo_self = cand_less(d[0], i[0], rd, ri);
            const float hd = lo_self ? rd : d[0];
            const int hi = lo_self ? ri : i[0];
            if (!lo_self) { d[0] = rd; i[0] = ri; }
            d[KPL - 1] = hd; i[KPL - 1] = hi;
            bitonic_merge32(d[0], i[0], lane);
            bitonic_merge32(d[KPL - 1], i[KPL - 1], lane);
        }
    }
    // "Empty" members that survived the search (fewer than K candidates with a comparable distance: NaN coordinates in
    // the query or the cloud -- a NaN distance never passes cand_less) carry indices above every real point.  The callers
    // of knn_points dereference the indices unchecked, so they are mapped into [0, N) before anything is written:
    // member 0x7fffffff - p becomes point p (distinct, in range; its distance stays +inf).  pytorch3d likewise returns
    // in-range indices and lets the NaN propagate through the values.
    __device__ __forceinline__ void sanitize(int N) {
#pragma unroll
        for (int s = 0; s < KPL; ++s)
            if (i[s] != -1 && (unsigned)i[s] >= (unsigned)N) {
                const int p = 0x7fffffff - i[s];
                i[s] = (p >= 0 && p < N) ? p : 0;
            }
    }
    // after sort_set: f(pos, dist, index) for the K members, pos = 0..K-1 ascending
    template <typename F>
    __device__ __forceinline__ void for_each_sorted(int K, int lane, F f) const {
        const int ndummy = 32 * KPL - K;
#pragma unroll
        for (int s = 0; s < KPL; ++s) {
            const int pos = s * 32 + lane - ndummy;
            if (pos >= 0) f(pos, d[s], i[s]);
        }
    }
};

}  // namespace knn_sel
