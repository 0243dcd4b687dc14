/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the product path
 * (package pcd_reg_hregnet_b200); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker / the CPU baseline.
 *
 * Plain-C, CPU restatement of the NATIVE ops on the HRegNet registration forward path:
 *
 *   oracle_fps / weighted       <- /root/reference/models/PointUtils/src/furthest_point_sampling_gpu.cu
 *                                  :84-206 (FPS kernel), :254-375 (weighted), :75-80 (__update),
 *                                  cuda_utils.h:22-27 (opt_n_threads), models/utils.py:14-58 (buffers)
 *   oracle_gather_points        <- furthest_point_sampling_gpu.cu:7-21
 *   oracle_gather_points_grad   <- furthest_point_sampling_gpu.cu:41-55
 *   oracle_knn                  <- pytorch3d.ops.knn_points (pytorch3d==0.7.8, Dockerfile:45-46; NOT
 *                                  vendored in the reference).  Call sites: models/HRegNet/layers.py:20,
 *                                  278,316,322,434.  PARITY UNPINNED for this function: the reference has
 *                                  no test/golden vector for it and the dependency is not installable
 *                                  here, so this file DEFINES the contract:
 *                                  squared L2 in fp32, accumulated d = 0..D-1 as dist = fmaf(diff,diff,dist);
 *                                  K smallest ordered by (dist asc, index asc); int64 indices.
 *
 * Parity status of FPS / gather: the rounding sequence and the tie-break below are checked on the GPU
 * box against the UNMODIFIED reference kernels recompiled for sm_100a (oracle/build_ref.py ->
 * oracle/_ref/point_utils_cuda.so, tests/test_gpu_ref_ext.py).
 *
 * The FPS restatement is a literal, sequential emulation of the CUDA block: per-"thread" strided scan
 * with strict '>' (first index wins inside a thread), followed by the same halving tree
 * (slot tid merges slot tid+s, keeps tid on ties) -- so the reference's tie-break is reproduced by
 * construction rather than by a closed-form rule.
 *
 * Compile: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC  (see oracle/Makefile)
 * -ffp-contract=off matters: every fused multiply-add below is an explicit fmaf().
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* cuda_utils.h:22-27 -- same double-precision log ratio, truncation and clamp */
int oracle_opt_n_threads(int work_size) {
    const int pow_2 = (int)(log((double)work_size) / log(2.0));
    int t = 1 << pow_2;
    if (t > 1024) t = 1024;
    if (t < 1) t = 1;
    return t;
}

/* The distance of furthest_point_sampling_gpu.cu:127 as nvcc compiles it (-O2, default -fmad=true,
 * verified in the SASS of the sm_100a rebuild, see DESIGN.md "rounding contract"):
 *     FMUL t = dy*dy ; FFMA t = dx*dx + t ; FFMA d = dz*dz + t                                        */
static inline float fps_dist(float x1, float y1, float z1, float x2, float y2, float z2) {
    const float dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* xyz [B,N,3]; weights [B,N] or NULL; temp [B,N] (in: caller-initialised, 1e10 by models/utils.py:25;
 * out: clobbered with the final min-distances); idx [B,M] int32 out. */
void oracle_fps(const float *xyz, const float *weights, float *temp, int32_t *idx, int B, int N, int M) {
    if (M <= 0) return;                                   /* .cu:92 / :261 */
    const int T = oracle_opt_n_threads(N);
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        const float *p = xyz + (size_t)b * N * 3;
        const float *w = weights ? weights + (size_t)b * N : NULL;
        float *tmp = temp + (size_t)b * N;
        int32_t *out = idx + (size_t)b * M;
        float *dists = (float *)malloc(sizeof(float) * T);
        int *dists_i = (int *)malloc(sizeof(int) * T);
        int old = 0;
        out[0] = old;                                     /* .cu:107-109 */
        for (int j = 1; j < M; ++j) {
            const float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
            /* Every CUDA thread tid scans k = tid, tid+T, ... in ascending k with a strict '>' (.cu:117-134).
             * The per-thread states are independent, so visiting k = 0..N-1 in order and updating slot
             * k mod T performs exactly the same comparisons in the same per-thread order. */
            for (int tid = 0; tid < T; ++tid) { dists[tid] = -1.0f; dists_i[tid] = 0; }
            for (int k = 0; k < N; ++k) {
                const int tid = k & (T - 1);              /* T is a power of two */
                float d = fps_dist(x1, y1, z1, p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2]);
                if (w) d = w[k] * d;                      /* .cu:299: one more rounding, candidate's weight */
                const float d2 = fminf(d, tmp[k]);
                tmp[k] = d2;
                if (d2 > dists[tid]) { dists[tid] = d2; dists_i[tid] = k; }
            }
            for (int s = T >> 1; s >= 1; s >>= 1) {       /* .cu:140-199, __update at :75-80 */
                for (int tid = 0; tid < s; ++tid) {
                    const float v1 = dists[tid], v2 = dists[tid + s];
                    const int i1 = dists_i[tid], i2 = dists_i[tid + s];
                    dists[tid] = fmaxf(v1, v2);
                    dists_i[tid] = v2 > v1 ? i2 : i1;
                }
            }
            old = dists_i[0];
            out[j] = old;
        }
        free(dists);
        free(dists_i);
    }
}

/* points [B,C,N], idx [B,M] int32 -> out [B,C,M]   (.cu:7-21) */
void oracle_gather_points(const float *points, const int32_t *idx, float *out, int B, int C, int N, int M) {
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int m = 0; m < M; ++m)
                out[((size_t)b * C + c) * M + m] = points[((size_t)b * C + c) * N + idx[(size_t)b * M + m]];
}

/* grad_out [B,C,M], idx [B,M] -> grad_points [B,C,N] += (caller pre-zeroes; .cu:41-55, utils.py:84) */
void oracle_gather_points_grad(const float *grad_out, const int32_t *idx, float *grad_points,
                               int B, int C, int N, int M) {
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int m = 0; m < M; ++m)
                grad_points[((size_t)b * C + c) * N + idx[(size_t)b * M + m]] +=
                    grad_out[((size_t)b * C + c) * M + m];
}

/* ---- kNN ----------------------------------------------------------------------------------------- */
typedef struct { float d; int32_t i; } cand_t;

static inline int cand_less(cand_t a, cand_t b) {          /* (dist asc, idx asc) */
    return a.d < b.d || (a.d == b.d && a.i < b.i);
}

static void sift_down(cand_t *h, int n, int i) {          /* max-heap on cand_less */
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && cand_less(h[m], h[l])) m = l;
        if (r < n && cand_less(h[m], h[r])) m = r;
        if (m == i) return;
        cand_t t = h[i]; h[i] = h[m]; h[m] = t;
        i = m;
    }
}

/* p1 [B,M,D] queries, p2 [B,N,D] references -> dists [B,M,K] (squared, ascending), idx [B,M,K] int64,
 * nn [B,M,K,D] (may be NULL).  Requires K <= N. */
void oracle_knn(const float *p1, const float *p2, float *dists, int64_t *idx, float *nn,
                int B, int M, int N, int D, int K) {
#pragma omp parallel
    {
        cand_t *heap = (cand_t *)malloc(sizeof(cand_t) * (K > 0 ? K : 1));
#pragma omp for schedule(static) collapse(2)
        for (int b = 0; b < B; ++b) {
            for (int m = 0; m < M; ++m) {
                const float *q = p1 + ((size_t)b * M + m) * D;
                int n_heap = 0;
                for (int n = 0; n < N; ++n) {
                    const float *r = p2 + ((size_t)b * N + n) * D;
                    float dist = 0.0f;
                    for (int d = 0; d < D; ++d) {
                        const float diff = q[d] - r[d];
                        dist = fmaf(diff, diff, dist);
                    }
                    cand_t c = { dist, n };
                    if (n_heap < K) {
                        heap[n_heap++] = c;
                        if (n_heap == K)
                            for (int i = K / 2 - 1; i >= 0; --i) sift_down(heap, K, i);
                    } else if (cand_less(c, heap[0])) {
                        heap[0] = c;
                        sift_down(heap, K, 0);
                    }
                }
                /* heap-sort ascending */
                if (n_heap < K)   /* K > N: not supported, leave partially filled (caller checks) */
                    for (int i = n_heap / 2 - 1; i >= 0; --i) sift_down(heap, n_heap, i);
                for (int e = n_heap - 1; e > 0; --e) {
                    cand_t t = heap[0]; heap[0] = heap[e]; heap[e] = t;
                    sift_down(heap, e, 0);
                }
                for (int j = 0; j < n_heap; ++j) {
                    const size_t o = ((size_t)b * M + m) * K + j;
                    dists[o] = heap[j].d;
                    idx[o] = heap[j].i;
                    if (nn) memcpy(nn + o * D, p2 + ((size_t)b * N + heap[j].i) * D, sizeof(float) * D);
                }
            }
        }
        free(heap);
    }
}
