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

// fps_dist of two points against the same centre
__device__ __forceinline__ f32x2_t fps_dist2(f32x2_t x, f32x2_t y, f32x2_t z, f32x2_t cx, f32x2_t cy, f32x2_t cz) {
    const f32x2_t dx = f2_sub(x, cx), dy = f2_sub(y, cy), dz = f2_sub(z, cz);
    return f2_fma(dz, dz, f2_fma(dx, dx, f2_mul(dy, dy)));
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


// ---------------------------------------------------------------------------------------------------------------
// Cluster variant: CS CTAs (thread-block cluster) cooperate on ONE cloud; every point lives in registers
// (x, y, z, running min-distance [, weight]) -- no shared-memory traffic in the update loop at all.
// Per iteration (3 CTA barriers; warps that are not on the critical path are parked at a hardware barrier, they
// never spin):
//   A. every thread updates its P points (8 FP instructions each), warp maximum by REDUX, barrier;
//   B. only the warps that hold the CTA maximum resolve the reference's tie-break key and publish {key, xyz}, barrier;
//   C. warp 0 picks the CTA winner and sends it as a 32-byte packet to the mailbox of every CTA of the cluster with
//      st.async (distributed shared memory; the store itself performs complete_tx on the receiver's mbarrier), waits
//      on its own mbarrier for the CS packets of this iteration, takes the best, barrier.
// Mailboxes / mbarriers are double-buffered by iteration parity: a CTA can run at most one iteration ahead of its
// peers because iteration j+1 needs every peer's packet j.
struct __align__(16) FpsMail {
    unsigned ordval, key;
    float x, y;
    float z;
    unsigned pad0, pad1, pad2;
};

__device__ __forceinline__ uint32_t fps_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int THREADS, int P, bool WEIGHTED>
__global__ void __launch_bounds__(THREADS, 1)
fps_cluster_kernel(const float* __restrict__ xyz, const float* __restrict__ weights, float* __restrict__ temp_io,
                   int32_t* __restrict__ idx_out, int N, int M, int log2T, int CS) {
    constexpr int NWARP = THREADS / 32;
    __shared__ unsigned s_wmax[32];
    __shared__ FpsMail s_slot[32];
    __shared__ FpsMail s_mail[2][8];
    __shared__ FpsMail s_best;
    __shared__ __align__(8) uint64_t s_mbar[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int b = blockIdx.x / CS;
    xyz += (size_t)b * N * 3;
    if (WEIGHTED) weights += (size_t)b * N;
    if (temp_io) temp_io += (size_t)b * N;
    idx_out += (size_t)b * M;
    const unsigned tmask = (1u << log2T) - 1u;
    auto key_of = [&](int k) -> unsigned {
        const unsigned r = (log2T > 0) ? (__brev((unsigned)k & tmask) >> (32 - log2T)) : 0u;
        return (r << 12) | ((unsigned)k >> log2T);
    };
    const int stride = CS * THREADS;
    const int k0 = (int)rank * THREADS + tid;

    float px[P], py[P], pz[P], pt[P], pw[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int k = k0 + p * stride;
        const bool valid = k < N;
        px[p] = valid ? xyz[k * 3 + 0] : 0.f;
        py[p] = valid ? xyz[k * 3 + 1] : 0.f;
        pz[p] = valid ? xyz[k * 3 + 2] : 0.f;
        pw[p] = (WEIGHTED && valid) ? weights[k] : 0.f;
        pt[p] = valid ? (temp_io ? temp_io[k] : 1e10f) : -CUDART_INF_F;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fps_smem_u32(&s_mbar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fps_smem_u32(&s_mbar[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    float x1 = xyz[0], y1 = xyz[1], z1 = xyz[2];
    if (tid == 0 && rank == 0) idx_out[0] = 0;
    __syncthreads();
    if (CS > 1) {   // every CTA of the cluster has initialised its mbarriers before anyone signals them remotely
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }

    for (int j = 1; j < M; ++j) {
        // ---- A: update, warp maximum --------------------------------------------------------------------------
        float m = -CUDART_INF_F;
        if (P % 2 == 0) {
            const f32x2_t cx2 = f2_pack(x1, x1), cy2 = f2_pack(y1, y1), cz2 = f2_pack(z1, z1);
#pragma unroll
            for (int p = 0; p < P; p += 2) {
                f32x2_t d2 = fps_dist2(f2_pack(px[p], px[p + 1]), f2_pack(py[p], py[p + 1]), f2_pack(pz[p], pz[p + 1]), cx2, cy2, cz2);
                if (WEIGHTED) d2 = f2_mul(f2_pack(pw[p], pw[p + 1]), d2);
                float d0, d1;
                f2_unpack(d2, d0, d1);
                pt[p] = fminf(d0, pt[p]);
                pt[p + 1] = fminf(d1, pt[p + 1]);
                m = fmaxf(m, fmaxf(pt[p], pt[p + 1]));
            }
        } else {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                float d = fps_dist(px[p], py[p], pz[p], x1, y1, z1);
                if (WEIGHTED) d = __fmul_rn(pw[p], d);
                pt[p] = fminf(d, pt[p]);
                m = fmaxf(m, pt[p]);
            }
        }
        const unsigned om = hrn_ford(m);
        const unsigned wm = __reduce_max_sync(0xffffffffu, om);
        if (lane == 0) s_wmax[warp] = wm;
        __syncthreads();
        // ---- B: holders of the CTA maximum resolve the tie-break key ---------------------------------------------
        const unsigned g = __reduce_max_sync(0xffffffffu, lane < NWARP ? s_wmax[lane] : 0u);
        if (wm == g) {                                   // warp-uniform
            unsigned key = 0xffffffffu;
            float cx = 0.f, cy = 0.f, cz = 0.f;
            if (om == g) {
                if ((stride & (int)tmask) == 0) {
                    // stride is a multiple of Tref: (k mod Tref) is the same for all of this thread's points, so the
                    // reference key grows with p -> the winner is simply the first p that holds the maximum
                    int pf = 0;
#pragma unroll
                    for (int p = P - 1; p >= 0; --p)
                        if (pt[p] == m && k0 + p * stride < N) pf = p;
                    cx = px[0]; cy = py[0]; cz = pz[0];
#pragma unroll
                    for (int p = 1; p < P; ++p)
                        if (pf == p) { cx = px[p]; cy = py[p]; cz = pz[p]; }
                    key = key_of(k0 + pf * stride);
                } else {
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const int k = k0 + p * stride;
                        const unsigned kk = key_of(k);
                        if (k < N && pt[p] == m && kk < key) { key = kk; cx = px[p]; cy = py[p]; cz = pz[p]; }
                    }
                }
            }
            const unsigned wkey = __reduce_min_sync(0xffffffffu, key);
            if (key == wkey && om == g) {
                s_slot[warp].key = key; s_slot[warp].x = cx; s_slot[warp].y = cy; s_slot[warp].z = cz;
            }
        } else if (lane == 0) {
            s_slot[warp].key = 0xffffffffu;
        }
        __syncthreads();
        // ---- C: warp 0 = CTA winner; cluster exchange through the mailboxes ------------------------------------
        const int par = j & 1;
        const uint32_t mb = fps_smem_u32(&s_mbar[par]);
        if (warp == 0) {
            const unsigned k2 = lane < NWARP ? s_slot[lane].key : 0xffffffffu;
            const unsigned kmin = __reduce_min_sync(0xffffffffu, k2);
            const int src = __ffs(__ballot_sync(0xffffffffu, k2 == kmin)) - 1;
            const float sx = s_slot[src].x, sy = s_slot[src].y, sz = s_slot[src].z;
            if (CS > 1) {
                if (lane == 0)
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(CS * 32) : "memory");
                if (lane < CS) {
                    uint32_t rbox, rbar;
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbox) : "r"(fps_smem_u32(&s_mail[par][rank])), "r"((unsigned)lane));
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(mb), "r"((unsigned)lane));
                    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                                 ::"r"(rbox), "r"(g), "r"(kmin), "r"(__float_as_uint(sx)), "r"(__float_as_uint(sy)), "r"(rbar) : "memory");
                    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                                 ::"r"(rbox + 16), "r"(__float_as_uint(sz)), "r"(0u), "r"(0u), "r"(0u), "r"(rbar) : "memory");
                }
            } else if (lane == 0) {
                s_best.key = kmin; s_best.x = sx; s_best.y = sy; s_best.z = sz;
            }
        }
        unsigned bkey;
        if (CS > 1) {
            // every thread sleeps on the CTA's own mailbox barrier until the CS packets of iteration j have landed
            // (hardware wait, no polling traffic), then takes the best of them -- no third CTA barrier
            unsigned done;
            const unsigned parity = (unsigned)((j - 1) >> 1) & 1u;   // u-th use of mbarrier j&1, u = (j-1)/2
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(mb), "r"(parity) : "memory");
            } while (!done);
            unsigned bval = 0u;
            bkey = 0xffffffffu;
#pragma unroll 1
            for (int c = 0; c < CS; ++c) {
                const FpsMail mm = s_mail[par][c];
                if (mm.ordval > bval || (mm.ordval == bval && mm.key < bkey)) {
                    bval = mm.ordval; bkey = mm.key; x1 = mm.x; y1 = mm.y; z1 = mm.z;
                }
            }
        } else {
            __syncthreads();
            bkey = s_best.key; x1 = s_best.x; y1 = s_best.y; z1 = s_best.z;
        }
        if (tid == 0 && rank == 0) {
            const unsigned r = bkey >> 12;
            const int lo = (log2T > 0) ? (int)(__brev(r) >> (32 - log2T)) : 0;
            idx_out[j] = lo + (int)((bkey & 0xfffu) << log2T);
        }
    }
    if (temp_io) {
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int k = k0 + p * stride;
            if (k < N) temp_io[k] = pt[p];
        }
    }
    if (CS > 1) {   // nobody leaves while a peer may still write into its mailbox
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
}

template <int THREADS, int P, bool W>
int launch_fps_cluster(const float* xyz, const float* w, float* temp, int32_t* idx, int B, int N, int M, int log2T,
                       int CS, cudaStream_t st) {
    auto kern = fps_cluster_kernel<THREADS, P, W>;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B * CS);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    HRN_CUDA(cudaLaunchKernelEx(&cfg, kern, xyz, w, temp, idx, N, M, log2T, CS));
    return HRN_OK;
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
    // all-in-registers cluster kernel: cluster size = largest of {8,4,2,1} that keeps one wave (B*CS <= 148 SMs)
    // while P = ceil(N / (CS*THREADS)) <= 8 points per thread
    if (N <= 1024) {
        if (N <= 256) return launch_fps_cluster<256, 1, W>(xyz, w, temp, idx, B, N, M, log2T, 1, st);
        if (N <= 512) return launch_fps_cluster<256, 2, W>(xyz, w, temp, idx, B, N, M, log2T, 1, st);
        return launch_fps_cluster<256, 4, W>(xyz, w, temp, idx, B, N, M, log2T, 1, st);
    }
    if (N <= 65536) {
        // cluster size: as many CTAs per cloud as keep one wave (B*CS <= 148) but at least ~4096 points per CTA --
        // below that the DSMEM exchange costs more than the shorter update loop saves
        int CS = 8;
        while (CS > 1 && (B * CS > 148 || N / CS < 4096)) CS >>= 1;
        while (CS < 8 && (N + CS - 1) / CS > (W ? 4096 : 8192)) CS <<= 1;
        const int n_cta = (N + CS - 1) / CS;          // points per CTA
        if (n_cta <= 1024) return launch_fps_cluster<1024, 1, W>(xyz, w, temp, idx, B, N, M, log2T, CS, st);
        if (n_cta <= 2048) return launch_fps_cluster<1024, 2, W>(xyz, w, temp, idx, B, N, M, log2T, CS, st);
        if (n_cta <= 4096) return launch_fps_cluster<1024, 4, W>(xyz, w, temp, idx, B, N, M, log2T, CS, st);
        if (n_cta <= 8192 && !W) return launch_fps_cluster<512, 16, false>(xyz, w, temp, idx, B, N, M, log2T, CS, st);
    }
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
