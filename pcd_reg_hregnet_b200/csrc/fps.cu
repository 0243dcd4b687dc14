// Furthest-point sampling (plain and weighted) for sm_100a.
//
// Replaces: reference models/PointUtils/src/furthest_point_sampling_gpu.cu:84-252 (FPS) and :254-419
// (weighted FPS), reached through furthest_point_sampling.cpp:33-56 / models/utils.py:14-58.
//
// Design (one persistent CTA per cloud, everything on-chip):
//   * thread `tid` owns the points k = tid, tid+T, tid+2T, ... (T = the reference's opt_n_threads(N),
//     cuda_utils.h:22-27) -- xyz (and weights) stay in REGISTERS for the whole kernel, the running
//     min-distance array lives in SHARED memory (P*T floats, conflict-free column access), so an
//     iteration touches neither L2 nor HBM.  The reference re-reads xyz and reads+writes `temp` from global
//     memory every iteration and pays a 10-step __syncthreads tree.
//   * block arg-max = two REDUX (warp-wide integer max on an order-preserving float->uint map) around
//     ONE barrier for the value, then only the warp(s) that hold the maximum resolve the winner and
//     publish its coordinates through a shared slot (second barrier).  2 barriers / iteration instead of 11.
//   * bit-exactness contract with the reference:
//       - distance rounding  d = fma(dz,dz, fma(dx,dx, rn(dy*dy)))   [weighted: rn(w_k * d)]
//       - winner among equal maxima = smallest (bitrev_log2T(k mod T), k div T): the reference's strided
//         scan keeps the first k per thread (strict '>'), its halving tree keeps slot `tid` over `tid+s`,
//         i.e. prefers threads in bit-reversed order.  We reduce on exactly that key.
//   * N > 16*T (e.g. 32768..131072 points): streaming variant, xyz through L1/L2 and `temp` in global
//     memory (the caller's scratch buffer), same reduction.
#include "common.cuh"
#include <math_constants.h>

namespace {

struct FpsSlot {
    unsigned key;
    float x, y, z;
};

__device__ __forceinline__ float fps_dist(float x, float y, float z, float x1, float y1, float z1) {
    const float dx = x - x1, dy = y - y1, dz = z - z1;
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// P > 0: thread tid holds the P points k = tid + p*THREADS in registers, temp in shared memory.
// P == 0: streaming (xyz from global, temp in the caller's global scratch).
// The thread<->point mapping is free: the reference's tie-break is carried as an explicit key computed
// from the point index k:  key(k) = (bitrev_log2T(k mod Tref) << 12) | (k div Tref).
template <int THREADS, int P, bool WEIGHTED>
__global__ void __launch_bounds__(THREADS, 1)
fps_kernel(const float* __restrict__ xyz, const float* __restrict__ weights, float* __restrict__ temp_io,
           int32_t* __restrict__ idx_out, int N, int M, int log2T) {
    constexpr bool REG = (P > 0);
    constexpr int PP = REG ? P : 1;
    constexpr int NWARP = THREADS / 32;
    extern __shared__ float s_temp[];  // [P][THREADS]
    __shared__ unsigned s_wmax[32];
    __shared__ FpsSlot s_slot[2][32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    xyz += (size_t)b * N * 3;
    if (WEIGHTED) weights += (size_t)b * N;
    if (temp_io) temp_io += (size_t)b * N;
    idx_out += (size_t)b * M;
    const unsigned tmask = (1u << log2T) - 1u;
    auto key_of = [&](int k) -> unsigned {
        const unsigned r = (log2T > 0) ? (__brev((unsigned)k & tmask) >> (32 - log2T)) : 0u;
        return (r << 12) | ((unsigned)k >> log2T);
    };

    float px[PP], py[PP], pz[PP], pw[PP];
    if (REG) {
#pragma unroll
        for (int p = 0; p < PP; ++p) {
            const int k = tid + p * THREADS;
            const bool valid = k < N;
            px[p] = valid ? xyz[k * 3 + 0] : 0.f;
            py[p] = valid ? xyz[k * 3 + 1] : 0.f;
            pz[p] = valid ? xyz[k * 3 + 2] : 0.f;
            pw[p] = (WEIGHTED && valid) ? weights[k] : 0.f;
            s_temp[p * THREADS + tid] = valid ? (temp_io ? temp_io[k] : 1e10f) : -CUDART_INF_F;
        }
    }
    if (tid < 32) {
        s_slot[0][tid].key = 0xffffffffu;
        s_slot[1][tid].key = 0xffffffffu;
    }
    float x1 = xyz[0], y1 = xyz[1], z1 = xyz[2];
    if (tid == 0) idx_out[0] = 0;
    __syncthreads();

    for (int j = 1; j < M; ++j) {
        // ---- update min-distances against the last pick, track the thread's maximum ------------------
        float m = -CUDART_INF_F;
        if (REG) {
#pragma unroll
            for (int p = 0; p < PP; ++p) {
                float d = fps_dist(px[p], py[p], pz[p], x1, y1, z1);
                if (WEIGHTED) d = __fmul_rn(pw[p], d);
                const float t = s_temp[p * THREADS + tid];
                const float t2 = fminf(d, t);
                if (t2 < t) s_temp[p * THREADS + tid] = t2;
                m = fmaxf(m, t2);
            }
        } else {
            for (int k = tid; k < N; k += THREADS) {
                float d = fps_dist(xyz[k * 3 + 0], xyz[k * 3 + 1], xyz[k * 3 + 2], x1, y1, z1);
                if (WEIGHTED) d = __fmul_rn(weights[k], d);
                const float t = temp_io[k];
                const float t2 = fminf(d, t);
                if (t2 < t) temp_io[k] = t2;
                m = fmaxf(m, t2);
            }
        }
        // ---- block-wide maximum value: REDUX, one barrier, REDUX -------------------------------------
        const unsigned om = hrn_ford(m);
        const unsigned wm = __reduce_max_sync(0xffffffffu, om);
        if (lane == 0) s_wmax[warp] = wm;
        __syncthreads();
        const unsigned g = __reduce_max_sync(0xffffffffu, lane < NWARP ? s_wmax[lane] : 0u);
        const int buf = j & 1;
        // ---- only holders of the maximum resolve the reference's tie-break ---------------------------
        const bool cand = (om == g);
        const unsigned cmask = __ballot_sync(0xffffffffu, cand);
        if (cand) {
            unsigned key = 0xffffffffu;
            float cx = 0.f, cy = 0.f, cz = 0.f;
            if (REG) {
#pragma unroll
                for (int p = 0; p < PP; ++p) {
                    const int k = tid + p * THREADS;
                    const unsigned kk = key_of(k);
                    if (k < N && s_temp[p * THREADS + tid] == m && kk < key) {
                        key = kk; cx = px[p]; cy = py[p]; cz = pz[p];
                    }
                }
            } else {
                int kf = 0;
                for (int k = tid; k < N; k += THREADS) {
                    const unsigned kk = key_of(k);
                    if (temp_io[k] == m && kk < key) { key = kk; kf = k; }
                }
                cx = xyz[kf * 3 + 0]; cy = xyz[kf * 3 + 1]; cz = xyz[kf * 3 + 2];
            }
            const unsigned wkey = __reduce_min_sync(cmask, key);
            if (key == wkey) {
                FpsSlot s; s.key = key; s.x = cx; s.y = cy; s.z = cz;
                s_slot[buf][warp] = s;
            }
        }
        __syncthreads();
        // ---- every warp decodes the winner ------------------------------------------------------------
        const unsigned k2 = lane < NWARP ? s_slot[buf][lane].key : 0xffffffffu;
        const unsigned kmin = __reduce_min_sync(0xffffffffu, k2);
        const int src = __ffs(__ballot_sync(0xffffffffu, k2 == kmin)) - 1;
        x1 = s_slot[buf][src].x; y1 = s_slot[buf][src].y; z1 = s_slot[buf][src].z;
        if (lane == 0) s_slot[buf ^ 1][warp].key = 0xffffffffu;   // re-arm the other buffer
        if (tid == 0) {
            const unsigned r = kmin >> 12;
            const int lo = (log2T > 0) ? (int)(__brev(r) >> (32 - log2T)) : 0;
            idx_out[j] = lo + (int)((kmin & 0xfffu) << log2T);
        }
    }
    // drop-in: `temp` is scratch the reference leaves holding the final min-distances
    if (REG && temp_io) {
#pragma unroll
        for (int p = 0; p < PP; ++p) {
            const int k = tid + p * THREADS;
            if (k < N) temp_io[k] = s_temp[p * THREADS + tid];
        }
    }
}

template <int THREADS, int P, bool W>
int launch_fps(const float* xyz, const float* w, float* temp, int32_t* idx, int B, int N, int M, int log2T,
               cudaStream_t st) {
    const size_t smem = (size_t)(P > 0 ? P : 0) * THREADS * sizeof(float);
    auto kern = fps_kernel<THREADS, P, W>;
    if (smem > 48 * 1024) HRN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<B, THREADS, smem, st>>>(xyz, w, temp, idx, N, M, log2T);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}

template <bool W>
int dispatch_fps(const float* xyz, const float* w, float* temp, int32_t* idx, int B, int N, int M, cudaStream_t st) {
    // opt_n_threads (cuda_utils.h:22-27): 2^floor(log2 N) clamped to [1,1024]  (defines the tie-break only)
    int log2T = 0;
    while ((2 << log2T) <= N && log2T < 10) ++log2T;
    if (N <= 256) return launch_fps<256, 1, W>(xyz, w, temp, idx, B, N, M, log2T, st);
    if (N <= 512) return launch_fps<256, 2, W>(xyz, w, temp, idx, B, N, M, log2T, st);
    if (N <= 1024) return launch_fps<256, 4, W>(xyz, w, temp, idx, B, N, M, log2T, st);
    if (N <= 2048) return launch_fps<512, 4, W>(xyz, w, temp, idx, B, N, M, log2T, st);
    if (N <= 4096) return launch_fps<512, 8, W>(xyz, w, temp, idx, B, N, M, log2T, st);
    if (N <= 8192) return launch_fps<512, 16, W>(xyz, w, temp, idx, B, N, M, log2T, st);
    if (N <= 16384) return launch_fps<512, 32, W>(xyz, w, temp, idx, B, N, M, log2T, st);
    if (N > (4096 << 10) || temp == nullptr) return HRN_ERR_BAD_ARG;   // streaming path needs the scratch buffer
    return launch_fps<1024, 0, W>(xyz, w, temp, idx, B, N, M, log2T, st);
}

}  // namespace

HRN_API int hrn_fps(const float* xyz, const float* weights, float* temp, int32_t* idx, int B, int N, int M,
                    void* stream) {
    if (B < 0 || N <= 0 || !xyz || !idx) return HRN_ERR_BAD_ARG;
    if (M <= 0 || B == 0) return HRN_OK;  // reference kernels return immediately for m <= 0 (.cu:92,261)
    cudaStream_t st = (cudaStream_t)stream;
    return weights ? dispatch_fps<true>(xyz, weights, temp, idx, B, N, M, st)
                   : dispatch_fps<false>(xyz, nullptr, temp, idx, B, N, M, st);
}
