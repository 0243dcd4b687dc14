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
#include <stdlib.h>

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
// (x, y, z, running min-distance [, weight]) -- no shared-memory traffic in the update loop at all -- and the
// iteration has NO barrier of any kind (no CTA barrier, no mbarrier, no leader warp):
//   A. every thread updates its P points (packed fp32x2), warp maximum by REDUX;
//   B. every warp resolves the reference's tie-break key among its own holders of the warp maximum (REDUX min) and
//      writes ONE self-validating 8-byte packet {ordered max | key + iteration tag} into the mailbox slot (rank, warp)
//      of EVERY CTA of the cluster, its own included (one relaxed store per destination CTA);
//   C. every warp polls the CS*NWARP packets of this iteration (one or a few per lane) until all tags match, reduces
//      them with two REDUX, decodes the winner's point index from the key and reads its coordinates from a copy of
//      the whole cloud in shared memory (FULL: N <= 18 K points) or from global memory / L2 (larger clouds).
// What the ncu source view and tools/fps_probe.cu showed on B200: with 16 warps per SM the iteration is bound by the
// ISSUE slots of the per-warp bookkeeping (300 instructions per warp and iteration in the first barrier-free version,
// only 75 of them the distance update) and by branch-resolve stalls, not by the DSMEM round trip or the barriers -- three
// different synchronisation schemes (CTA barriers + st.async/mbarrier leader exchange, st.async packets per warp,
// polled packets per warp) all took 2130 cycles per iteration.  Hence the 8-byte packet (no coordinates, no select
// chains, no REDUX.OR broadcasts) and the branch-free send.
// 8-byte aligned stores/loads are single-copy atomic and every packet validates itself.  Mailboxes are double-buffered
// by iteration parity: iteration j+1 of any warp needs packet j of every warp of the cluster, so nobody can be more
// than one iteration ahead of anybody else, and buffer j&1 is only rewritten (iteration j+2) after every warp has
// consumed it.  Packet: low word = ordered maximum, high word = key (22 bits) | (j & 1023) << 22; the mailboxes start
// as all-ones, which matches no tag in use before it has been overwritten.
#ifdef FPS_PROBE
__device__ long long g_fps_probe[64 * 8];
__device__ unsigned g_fps_sink;
#define FPS_STAMP(i) do { if (blockIdx.x == 0 && tid == FPS_PROBE_TID && j >= 500 && j < 564) g_fps_probe[(j - 500) * 8 + (i)] = clock64(); } while (0)
#else
#define FPS_STAMP(i) do { } while (0)
#endif

__device__ __forceinline__ uint32_t fps_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int THREADS, int Q, int R, bool WEIGHTED, bool MONO, bool FULL>
__global__ void __launch_bounds__(THREADS, 1)
fps_cluster_kernel(const float* __restrict__ xyz, const float* __restrict__ weights, float* __restrict__ temp_io,
                   int32_t* __restrict__ idx_out, int N, int M, int log2T, int CS) {
    constexpr int NWARP = THREADS / 32;
    constexpr int P = Q * R;                               // points per thread: R residues (mod stride) x Q points each
    extern __shared__ __align__(16) unsigned char s_dyn[];
    const int S = CS * NWARP;                              // packets per iteration and mailbox
    unsigned long long* s_mail = reinterpret_cast<unsigned long long*>(s_dyn);      // [2][S]
    float* s_cloud = reinterpret_cast<float*>(s_mail + 2 * S);                       // FULL: [N][3]
    __shared__ __align__(8) uint64_t s_mbar[2];                 // MBAR only

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
    // thread <-> points: k(p) = k0 + r * (CS*THREADS) + q * stride with p = r*Q + q, stride = R * CS * THREADS: few,
    // fat warps (the per-warp bookkeeping, not the distance update, bounds an iteration) without giving up MONO
    const int span = CS * THREADS, stride = R * span;
    const int k0 = (int)rank * THREADS + tid;
    auto k_of = [&](int p) -> int { return k0 + (p / Q) * span + (p % Q) * stride; };
    // MONO (chosen by the launcher): stride is a multiple of Tref, so (k mod Tref) is the same for the Q points of a
    // residue and key(k0 + r*span + q*stride) = key(k0 + r*span) + q * (stride / Tref): the winner inside a residue is
    // simply the first q that holds the thread's maximum
    const unsigned key_step = (unsigned)stride >> log2T;
    unsigned key_base[R];
#pragma unroll
    for (int r = 0; r < R; ++r) key_base[r] = key_of(k0 + r * span);
    unsigned pkey[MONO ? 1 : P];                           // general case: the key of every point of the thread
    if (!MONO) {
#pragma unroll
        for (int p = 0; p < P; ++p) pkey[p] = (k_of(p) < N) ? key_of(k_of(p)) : 0xffffffffu;
    }

    float px[P], py[P], pz[P], pt[P], pw[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int k = k_of(p);
        const bool valid = k < N;
        px[p] = valid ? xyz[k * 3 + 0] : 0.f;
        py[p] = valid ? xyz[k * 3 + 1] : 0.f;
        pz[p] = valid ? xyz[k * 3 + 2] : 0.f;
        pw[p] = (WEIGHTED && valid) ? weights[k] : 0.f;
        pt[p] = valid ? (temp_io ? temp_io[k] : 1e10f) : -CUDART_INF_F;
    }
    if (FULL)
        for (int i = tid; i < 3 * N; i += THREADS) s_cloud[i] = xyz[i];
    for (int i = tid; i < 2 * S; i += THREADS) s_mail[i] = ~0ull;
    // lane c < CS writes this warp's packet into CTA c: generic address of the slot (parity 0; parity 1 = + S*8 bytes)
    const uint32_t mail_u32 = fps_smem_u32(s_mail);
    const uint32_t mail_par_bytes = (uint32_t)S * 8u;
    // How a warp waits for the packets (measured, tools/fps_probe.cu): with 8 warps per CTA plain polling of the tagged
    // packets is fastest (342 vs 414 ns per iteration at N = 1024); with 16 warps the polling loads of the early warps
    // fight the late warps for issue slots and the shared-memory port (a tight poll loop: 1710 ns, a lazier one: 815 ns
    // at N = 16384), so those kernels sleep on an mbarrier instead and the packets arrive by st.async (807 ns).
    constexpr bool MBAR = THREADS > 256;
    // st.async delivers the packet and counts its 8 bytes on the receiver's mbarrier; the waiters sleep in hardware
    const uint32_t mbar_u32 = fps_smem_u32(&s_mbar[0]);
    const unsigned tx_bytes = (unsigned)S * 8u;
    uint32_t rslot32 = 0u, rbar32 = 0u;
    unsigned long long rslot = 0ull;
    if (MBAR && lane < CS) {
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rslot32)
                     : "r"(mail_u32 + (uint32_t)((int)rank * NWARP + warp) * 8u), "r"((unsigned)lane));
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar32) : "r"(mbar_u32), "r"((unsigned)lane));
    }
    if (MBAR && tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_u32) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_u32 + 8u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // iterations 1 (parity 1) and 2 (parity 0) are armed here, iteration j+2 right after iteration j completed
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_u32), "r"(tx_bytes) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_u32 + 8u), "r"(tx_bytes) : "memory");
    }
    if (!MBAR && lane < CS) {
        uint32_t r32;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r32)
                     : "r"(mail_u32 + (uint32_t)((int)rank * NWARP + warp) * 8u), "r"((unsigned)lane));
        asm volatile("cvta.shared::cluster.u64 %0, %1;" : "=l"(rslot) : "l"((unsigned long long)r32));
    }
    float x1 = xyz[0], y1 = xyz[1], z1 = xyz[2];
    if (tid == 0 && rank == 0) idx_out[0] = 0;
    __syncthreads();
    // every CTA of the cluster has initialised its mailboxes before anyone writes into them
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");

#pragma unroll 1
    for (int j = 1; j < M; ++j) {
        // ---- A: update, warp maximum --------------------------------------------------------------------------
        FPS_STAMP(0);
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
        // unweighted distances are >= +0 (or -inf for padding): flipping the sign bit is order preserving there
        const unsigned om = WEIGHTED ? hrn_ford(m) : (__float_as_uint(m) ^ 0x80000000u);
        const unsigned wm = __reduce_max_sync(0xffffffffu, om);
        FPS_STAMP(1);
        // ---- B: the warp's winner (reference tie-break key among the holders of the warp maximum), packet out ------
        // (a thread without any valid point holds -inf: its warp can only tie with it if the whole warp is padding, and
        // such a packet never wins: some warp of the cluster always holds a valid point with a distance >= 0)
        unsigned key;
        if (MONO) {
            key = 0xffffffffu;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                unsigned kk = (R == 1) ? key_base[r] + (unsigned)(Q - 1) * key_step : 0xffffffffu;
#pragma unroll
                for (int q = (R == 1) ? Q - 2 : Q - 1; q >= 0; --q)
                    if (pt[r * Q + q] == m) kk = key_base[r] + (unsigned)q * key_step;   // per-thread constants
                key = min(key, kk);
            }
        } else {
            key = 0xffffffffu;
#pragma unroll
            for (int p = 0; p < P; ++p)
                if (pt[p] == m && pkey[p] < key) key = pkey[p];
        }
        if (om != wm) key = 0xffffffffu;
        const unsigned wkey = __reduce_min_sync(0xffffffffu, key);
        const int par = j & 1;
        const unsigned tag = ((unsigned)j & 0x3ffu) << 22;
        if (lane < CS) {
            const unsigned long long v = ((unsigned long long)((wkey & 0x3fffffu) | tag) << 32) | wm;
            if (MBAR)
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
                         ::"r"(rslot32 + (par ? mail_par_bytes : 0u)), "l"(v), "r"(rbar32 + (par ? 8u : 0u)) : "memory");
            else if (CS == 1)   // single-CTA launch: plain shared-memory store, no trip through the cluster window
            asm volatile("st.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(mail_u32 + (uint32_t)warp * 8u + (par ? mail_par_bytes : 0u)), "l"(v) : "memory");
            else
            asm volatile("st.relaxed.cluster.b64 [%0], %1;" ::"l"(rslot + (par ? mail_par_bytes : 0u)), "l"(v) : "memory");
        }
        FPS_STAMP(2);
        if (MBAR) {
            const uint32_t mb = mbar_u32 + (par ? 8u : 0u);
            const unsigned parity = (unsigned)((j - 1) >> 1) & 1u;   // u-th use of mbarrier j&1, u = (j-1)/2
            unsigned done;
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(mb), "r"(parity) : "memory");
            } while (!done);
            if (tid == 0 && j + 2 < M)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(tx_bytes) : "memory");
        }
        // ---- C: poll this iteration's packets (one or a few per lane), pick the winner ---------------------------------
        unsigned bo = 0u, bk = 0x3fffffu;
        {
            const uint32_t mail = mail_u32 + (par ? mail_par_bytes : 0u) + (uint32_t)lane * 8u;
            if (lane < S) {
                unsigned h;
                for (;;) {
                    asm volatile("ld.relaxed.cluster.shared::cta.v2.u32 {%0, %1}, [%2];" : "=r"(bo), "=r"(h) : "r"(mail) : "memory");
                    if ((h & 0xffc00000u) == tag) break;
                }
                bk = h & 0x3fffffu;
            }
            for (int sl = 32; sl < S; sl += 32) {                  // S is a multiple of 32 here (NWARP >= 8, CS >= 4)
                unsigned o, h;
                do {
                    asm volatile("ld.relaxed.cluster.shared::cta.v2.u32 {%0, %1}, [%2];" : "=r"(o), "=r"(h) : "r"(mail + (uint32_t)sl * 8u) : "memory");
                } while ((h & 0xffc00000u) != tag);
                h &= 0x3fffffu;
                if (o > bo || (o == bo && h < bk)) { bo = o; bk = h; }
            }
        }
        FPS_STAMP(3);
        const unsigned g = __reduce_max_sync(0xffffffffu, bo);
        const unsigned kmin = __reduce_min_sync(0xffffffffu, (bo == g) ? bk : 0xffffffffu);
        const int kw = ((log2T > 0) ? (int)(__brev(kmin >> 12) >> (32 - log2T)) : 0) + (int)((kmin & 0xfffu) << log2T);
        if (FULL) {
            x1 = s_cloud[kw * 3 + 0]; y1 = s_cloud[kw * 3 + 1]; z1 = s_cloud[kw * 3 + 2];
        } else {
            x1 = __ldg(xyz + kw * 3 + 0); y1 = __ldg(xyz + kw * 3 + 1); z1 = __ldg(xyz + kw * 3 + 2);
        }
#ifdef FPS_PROBE
        if (__float_as_uint(x1) + __float_as_uint(y1) + __float_as_uint(z1) == 0x12345u) g_fps_sink = 3;   // consume
#endif
        FPS_STAMP(4);
        if (tid == 0 && rank == 0) idx_out[j] = kw;
    }
    if (temp_io) {
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int k = k_of(p);
            if (k < N) temp_io[k] = pt[p];
        }
    }
    // nobody leaves while a peer may still write into its mailbox
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// Culled cluster variant (unweighted, whole cloud in shared memory): the same iteration as fps_cluster_kernel, but
// the points are first BINNED by a 10-bit Morton cell (one counting sort in shared memory), so that the 512 points of
// a warp are neighbours in space, and every warp keeps the bounding box of its points.  A new centre can only lower
// the min-distance of a point that is closer to it than that min-distance; per warp:
//     bound = fma(bz,bz, fma(bx,bx, rn(by*by))),   b? = max(lo? - c?, c? - hi?, 0)      (the kernel's own rounding)
// is a lower bound of every distance the update would compute (rn, x*x and fma are monotone), so if
// bound >= max_p pt[p] the update cannot change anything: the warp skips it and resends its previous packet.  Late in
// the sampling a centre touches 1-3 of the 32 warps of a cloud; the iteration is then the packet exchange alone.
// The result is bit-identical: same arithmetic per point, the reference's tie-break key travels with every point
// (the thread <-> point mapping is free, see fps_kernel).
template <int THREADS, int Q, bool MBAR>
__global__ void __launch_bounds__(THREADS, 1)
fps_cull_kernel(const float* __restrict__ xyz, float* __restrict__ temp_io, int32_t* __restrict__ idx_out, int N, int M,
                int log2T) {
    constexpr int NWARP = THREADS / 32;
    constexpr int CS = 2;
    constexpr int PER_CTA = THREADS * Q;                   // sorted positions per CTA
    constexpr int NBIN = 1024;
    constexpr int S = CS * NWARP;                          // packets per iteration
    static_assert(S <= 32, "at most one packet per lane");
    static_assert(Q % 2 == 0, "packed update");
    extern __shared__ __align__(16) unsigned char s_dyn[];
    unsigned long long* s_mail = reinterpret_cast<unsigned long long*>(s_dyn);      // [2][S]
    float* s_cloud = reinterpret_cast<float*>(s_mail + 2 * S);                       // [N][3]
    unsigned* s_hist = reinterpret_cast<unsigned*>(s_cloud + 3 * (size_t)N);          // [NBIN]
    unsigned short* s_perm = reinterpret_cast<unsigned short*>(s_hist + NBIN);        // [PER_CTA]
    __shared__ __align__(8) uint64_t s_mbar[2];
    __shared__ float s_box[NWARP][6];
    __shared__ unsigned s_wsum[NWARP];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int b = blockIdx.x / CS;
    xyz += (size_t)b * N * 3;
    if (temp_io) temp_io += (size_t)b * N;
    idx_out += (size_t)b * M;
    const unsigned tmask = (1u << log2T) - 1u;
    auto key_of = [&](int k) -> unsigned {
        const unsigned r = (log2T > 0) ? (__brev((unsigned)k & tmask) >> (32 - log2T)) : 0u;
        return (r << 12) | ((unsigned)k >> log2T);
    };
    auto k_of_key = [&](unsigned key) -> int {
        return ((log2T > 0) ? (int)(__brev(key >> 12) >> (32 - log2T)) : 0) + (int)((key & 0xfffu) << log2T);
    };

    // ---- the cloud, its bounding box ----------------------------------------------------------------------------
    for (int i = tid; i < 3 * N; i += THREADS) s_cloud[i] = xyz[i];
    for (int i = tid; i < 2 * S; i += THREADS) s_mail[i] = ~0ull;
    for (int i = tid; i < NBIN; i += THREADS) s_hist[i] = 0u;
    __syncthreads();
    float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    for (int k = tid; k < N; k += THREADS) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { const float v = s_cloud[k * 3 + a]; lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v); }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { s_box[warp][a] = lo[a]; s_box[warp][3 + a] = hi[a]; }
    }
    __syncthreads();
    float sc[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float l = CUDART_INF_F, h = -CUDART_INF_F;
        for (int w = 0; w < NWARP; ++w) { l = fminf(l, s_box[w][a]); h = fmaxf(h, s_box[w][3 + a]); }
        lo[a] = l;
        const float ext = h - l;
        const float cells = a == 0 ? 16.f : 8.f;                          // 4 + 3 + 3 bits
        sc[a] = (ext > 0.f && ext < CUDART_INF_F) ? cells / ext : 0.f;
    }
    // cell of a point: Morton interleave of the quantised coordinates (x: 4 bits, y, z: 3 bits); any total function of
    // the coordinates would do -- the cells only decide which thread holds which point
    auto cell_of = [&](int k) -> unsigned {
        const unsigned qx = (unsigned)fminf(fmaxf((s_cloud[k * 3 + 0] - lo[0]) * sc[0], 0.f), 15.f);
        const unsigned qy = (unsigned)fminf(fmaxf((s_cloud[k * 3 + 1] - lo[1]) * sc[1], 0.f), 7.f);
        const unsigned qz = (unsigned)fminf(fmaxf((s_cloud[k * 3 + 2] - lo[2]) * sc[2], 0.f), 7.f);
        unsigned c = (qx >> 3) << 9;
#pragma unroll
        for (int i = 0; i < 3; ++i)
            c |= (((qx >> i) & 1u) << (3 * i + 2)) | (((qy >> i) & 1u) << (3 * i + 1)) | (((qz >> i) & 1u) << (3 * i));
        return c;
    };
    // ---- counting sort by cell: histogram, exclusive scan, scatter.  The order inside a cell is whatever the atomics
    // give, so ONE CTA (rank 0) sorts and writes the peer's half of the order into the peer's shared memory: two
    // independent sorts would disagree about the cell that straddles the halves (points held twice or by nobody).
    if (rank == 0) {
        for (int k = tid; k < N; k += THREADS) atomicAdd(&s_hist[cell_of(k)], 1u);
        __syncthreads();
        constexpr int BPT = NBIN / THREADS;                              // bins per thread (2)
        unsigned v[BPT], run = 0;
#pragma unroll
        for (int i = 0; i < BPT; ++i) { v[i] = s_hist[tid * BPT + i]; run += v[i]; }
        unsigned inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_wsum[warp] = inc;
        __syncthreads();
        unsigned base = 0;
        for (int w = 0; w < warp; ++w) base += s_wsum[w];
        unsigned off = base + inc - run;
#pragma unroll
        for (int i = 0; i < BPT; ++i) { s_hist[tid * BPT + i] = off; off += v[i]; }
        __syncthreads();
        uint32_t perm_peer;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(perm_peer) : "r"(fps_smem_u32(s_perm)), "r"(1u));
        for (int k = tid; k < N; k += THREADS) {
            const unsigned pos = atomicAdd(&s_hist[cell_of(k)], 1u);
            if (pos < (unsigned)PER_CTA) s_perm[pos] = (unsigned short)k;
            else asm volatile("st.shared::cluster.u16 [%0], %1;" ::"r"(perm_peer + 2u * (pos - (unsigned)PER_CTA)), "h"((unsigned short)k) : "memory");
        }
    }
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    // ---- this thread's points: sorted positions warp*32*Q + q*32 + lane of this CTA's half ---------------------------
    float px[Q], py[Q], pz[Q], pt[Q];
    unsigned pkey[Q];
    float wlo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, whi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int pl = warp * 32 * Q + q * 32 + lane;
        const bool valid = (int)rank * PER_CTA + pl < N;
        const int k = valid ? (int)s_perm[pl] : 0;
        px[q] = valid ? s_cloud[k * 3 + 0] : 0.f;
        py[q] = valid ? s_cloud[k * 3 + 1] : 0.f;
        pz[q] = valid ? s_cloud[k * 3 + 2] : 0.f;
        pkey[q] = valid ? key_of(k) : 0xffffffffu;
        pt[q] = valid ? (temp_io ? temp_io[k] : 1e10f) : -CUDART_INF_F;
        if (valid) {
            wlo[0] = fminf(wlo[0], px[q]); whi[0] = fmaxf(whi[0], px[q]);
            wlo[1] = fminf(wlo[1], py[q]); whi[1] = fmaxf(whi[1], py[q]);
            wlo[2] = fminf(wlo[2], pz[q]); whi[2] = fmaxf(whi[2], pz[q]);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            wlo[a] = fminf(wlo[a], __shfl_xor_sync(0xffffffffu, wlo[a], o));
            whi[a] = fmaxf(whi[a], __shfl_xor_sync(0xffffffffu, whi[a], o));
        }
    }
    // ---- packet exchange set-up (as fps_cluster_kernel, mbarrier variant) -----------------------------------------
    const uint32_t mail_u32 = fps_smem_u32(s_mail);
    constexpr uint32_t mail_par_bytes = (uint32_t)S * 8u;
    const uint32_t mbar_u32 = fps_smem_u32(&s_mbar[0]);
    constexpr unsigned tx_bytes = (unsigned)S * 8u;
    uint32_t rslot32 = 0u, rbar32 = 0u;
    unsigned long long rslot = 0ull;
    if (lane < CS) {
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rslot32)
                     : "r"(mail_u32 + (uint32_t)((int)rank * NWARP + warp) * 8u), "r"((unsigned)lane));
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar32) : "r"(mbar_u32), "r"((unsigned)lane));
        asm volatile("cvta.shared::cluster.u64 %0, %1;" : "=l"(rslot) : "l"((unsigned long long)rslot32));
    }
    if (MBAR && tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_u32) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_u32 + 8u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_u32), "r"(tx_bytes) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_u32 + 8u), "r"(tx_bytes) : "memory");
    }
    float x1 = s_cloud[0], y1 = s_cloud[1], z1 = s_cloud[2];
    if (tid == 0 && rank == 0) idx_out[0] = 0;
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");

    unsigned wm = 0u, wkey = 0xffffffffu;                  // this warp's packet: ordered maximum, key of its holder
    float wmf = CUDART_INF_F;                              // the warp maximum as a float
#pragma unroll 1
    for (int j = 1; j < M; ++j) {
        FPS_STAMP(0);
        // ---- A: can the new centre lower any min-distance of this warp?  (warp-uniform) -----------------------------
        const float bx = fmaxf(fmaxf(__fsub_rn(wlo[0], x1), __fsub_rn(x1, whi[0])), 0.f);
        const float by = fmaxf(fmaxf(__fsub_rn(wlo[1], y1), __fsub_rn(y1, whi[1])), 0.f);
        const float bz = fmaxf(fmaxf(__fsub_rn(wlo[2], z1), __fsub_rn(z1, whi[2])), 0.f);
        const float bound = __fmaf_rn(bz, bz, __fmaf_rn(bx, bx, __fmul_rn(by, by)));
        if (j == 1 || !(bound >= wmf)) {                   // (a NaN bound runs the update, which then changes nothing)
            float m = -CUDART_INF_F;
            const f32x2_t cx2 = f2_pack(x1, x1), cy2 = f2_pack(y1, y1), cz2 = f2_pack(z1, z1);
#pragma unroll
            for (int p = 0; p < Q; p += 2) {
                const f32x2_t d2 = fps_dist2(f2_pack(px[p], px[p + 1]), f2_pack(py[p], py[p + 1]), f2_pack(pz[p], pz[p + 1]), cx2, cy2, cz2);
                float d0, d1;
                f2_unpack(d2, d0, d1);
                pt[p] = fminf(d0, pt[p]);
                pt[p + 1] = fminf(d1, pt[p + 1]);
                m = fmaxf(m, fmaxf(pt[p], pt[p + 1]));
            }
            // distances are >= +0 (or -inf for padding): flipping the sign bit is order preserving there
            const unsigned om = __float_as_uint(m) ^ 0x80000000u;
            wm = __reduce_max_sync(0xffffffffu, om);
            unsigned key = 0xffffffffu;
#pragma unroll
            for (int p = 0; p < Q; ++p)
                if (pt[p] == m && pkey[p] < key) key = pkey[p];
            if (om != wm) key = 0xffffffffu;
            wkey = __reduce_min_sync(0xffffffffu, key);
            wmf = __uint_as_float(wm ^ 0x80000000u);
        }
        FPS_STAMP(1);
        // ---- B: packet out (the previous one again if nothing changed) ---------------------------------------------
        const int par = j & 1;
        const unsigned tag = ((unsigned)j & 0x3ffu) << 22;
        if (lane < CS) {
            const unsigned long long v = ((unsigned long long)((wkey & 0x3fffffu) | tag) << 32) | wm;
            if (MBAR)
                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
                             ::"r"(rslot32 + (par ? mail_par_bytes : 0u)), "l"(v), "r"(rbar32 + (par ? 8u : 0u)) : "memory");
            else
                asm volatile("st.relaxed.cluster.b64 [%0], %1;" ::"l"(rslot + (par ? mail_par_bytes : 0u)), "l"(v) : "memory");
        }
        FPS_STAMP(2);
        if (MBAR) {
            const uint32_t mb = mbar_u32 + (par ? 8u : 0u);
            const unsigned parity = (unsigned)((j - 1) >> 1) & 1u;
            unsigned done;
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(mb), "r"(parity) : "memory");
            } while (!done);
            if (tid == 0 && j + 2 < M)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(tx_bytes) : "memory");
        }
        FPS_STAMP(3);
        // ---- C: this iteration's packets (one per lane), the winner ---------------------------------------------------
        unsigned bo = 0u, h = 0x3fffffu;
        if (S == 32 || lane < S) {
            const uint32_t mail = mail_u32 + (par ? mail_par_bytes : 0u) + (uint32_t)lane * 8u;
            for (;;) {
                asm volatile("ld.relaxed.cluster.shared::cta.v2.u32 {%0, %1}, [%2];" : "=r"(bo), "=r"(h) : "r"(mail) : "memory");
                if ((h & 0xffc00000u) == tag) break;
            }
        }
        const unsigned bk = h & 0x3fffffu;
        const unsigned g = __reduce_max_sync(0xffffffffu, bo);
        const unsigned kmin = __reduce_min_sync(0xffffffffu, (bo == g) ? bk : 0xffffffffu);
        const int kw = k_of_key(kmin);
        x1 = s_cloud[kw * 3 + 0]; y1 = s_cloud[kw * 3 + 1]; z1 = s_cloud[kw * 3 + 2];
#ifdef FPS_PROBE
        if (__float_as_uint(x1) + __float_as_uint(y1) + __float_as_uint(z1) == 0x12345u) g_fps_sink = 3;   // consume
        if (blockIdx.x == 0 && tid == FPS_PROBE_TID && j >= 500 && j < 564) g_fps_probe[(j - 500) * 8 + 5] = (long long)(!(bound >= wmf));
#endif
        FPS_STAMP(4);
        if (tid == 0 && rank == 0) idx_out[j] = kw;
    }
    if (temp_io) {
#pragma unroll
        for (int q = 0; q < Q; ++q)
            if (pkey[q] != 0xffffffffu) temp_io[k_of_key(pkey[q])] = pt[q];
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int THREADS, int Q, bool MBAR>
int launch_fps_cull_t(const float* xyz, float* temp, int32_t* idx, int B, int N, int M, int log2T, cudaStream_t st);

template <int THREADS, int Q>
int launch_fps_cull(const float* xyz, float* temp, int32_t* idx, int B, int N, int M, int log2T, cudaStream_t st) {
    static const bool poll = [] { const char* e = getenv("HRN_FPS_POLL"); return e && e[0] == '1'; }();
    static const bool fat = [] { const char* e = getenv("HRN_FPS_FAT"); return e && e[0] == '1'; }();   // 8 warps x 32 points
    if (fat && THREADS * Q == 8192) {
        if (poll) return launch_fps_cull_t<256, 32, false>(xyz, temp, idx, B, N, M, log2T, st);
        return launch_fps_cull_t<256, 32, true>(xyz, temp, idx, B, N, M, log2T, st);
    }
    if (poll) return launch_fps_cull_t<THREADS, Q, false>(xyz, temp, idx, B, N, M, log2T, st);
    return launch_fps_cull_t<THREADS, Q, true>(xyz, temp, idx, B, N, M, log2T, st);
}

template <int THREADS, int Q, bool MBAR>
int launch_fps_cull_t(const float* xyz, float* temp, int32_t* idx, int B, int N, int M, int log2T, cudaStream_t st) {
    auto kern = fps_cull_kernel<THREADS, Q, MBAR>;
    const size_t smem = (size_t)2 * 2 * (THREADS / 32) * 8 + (size_t)N * 12 + 1024 * 4 + (size_t)THREADS * Q * 2;
    static hrn_once_per_device attr;
    if (attr.need()) HRN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B * 2);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    HRN_CUDA(cudaLaunchKernelEx(&cfg, kern, xyz, temp, idx, N, M, log2T));
    return HRN_OK;
}

template <int THREADS, int Q, int R, bool W, bool MONO, bool FULL>
int launch_fps_cluster_t(const float* xyz, const float* w, float* temp, int32_t* idx, int B, int N, int M, int log2T,
                         int CS, cudaStream_t st) {
    auto kern = fps_cluster_kernel<THREADS, Q, R, W, MONO, FULL>;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B * CS);
    cfg.blockDim = dim3(THREADS);
    const size_t smem = (size_t)2 * CS * (THREADS / 32) * 8 + (FULL ? (size_t)N * 12 : 0);
    if (smem > 48 * 1024) HRN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (CS > 8) HRN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));   // 16 CTAs: opt-in
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    HRN_CUDA(cudaLaunchKernelEx(&cfg, kern, xyz, w, temp, idx, N, M, log2T, CS));
    return HRN_OK;
}

template <int THREADS, int Q, bool W, int R = 1>
int launch_fps_cluster(const float* xyz, const float* w, float* temp, int32_t* idx, int B, int N, int M, int log2T,
                       int CS, cudaStream_t st) {
    constexpr int P = Q * R;
    const bool mono = ((R * CS * THREADS) % (1 << log2T)) == 0;   // see MONO in the kernel
    const bool full = (size_t)N * 12 + (size_t)2 * CS * (THREADS / 32) * 8 <= 200 * 1024;   // whole cloud fits in shared memory
    // instantiated: the general key path only for P <= 8 (keys in registers), the L2 coordinate lookup only for the
    // 8192-points-per-CTA kernel (the only one used for clouds that do not fit)
    if constexpr (P <= 8) {
        if (!mono && full) return launch_fps_cluster_t<THREADS, Q, R, W, false, true>(xyz, w, temp, idx, B, N, M, log2T, CS, st);
    }
    if constexpr (P >= 8) {
        if (mono && !full) return launch_fps_cluster_t<THREADS, Q, R, W, true, false>(xyz, w, temp, idx, B, N, M, log2T, CS, st);
    }
    if (mono && full) return launch_fps_cluster_t<THREADS, Q, R, W, true, true>(xyz, w, temp, idx, B, N, M, log2T, CS, st);
    return HRN_ERR_UNSUPPORTED;
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

// ---------------------------------------------------------------------------------------------------------------
// A few warps per cloud, for the small power-of-two clouds of the lower hierarchy levels (N = 256 / 512 / 1024: the
// weighted samplings 1024 -> 512 and 512 -> 256 of models.py:31-39).  An iteration of these is a chain of dependent
// steps (update -> REDUX -> key -> REDUX -> exchange -> winner lookup); with 8 warps and 4 points per thread the tagged
// packet exchange alone cost ~650 cycles, with ONE warp (32 points per lane) the update runs at a single warp's issue
// rate (588 cycles).  NW warps with P = N / (32 NW) points per lane and the cheapest possible exchange -- one 8-byte
// packet per warp in shared memory, ONE named barrier per iteration (buffers alternate by iteration parity), every
// thread compares the NW packets itself -- sit between the two.
// Slot s of lane l of warp w holds point k = l + 32 * (w + NW * bitrev_logP(s)): the reference's tie-break key of that
// point is bitrev_logN(k) = bitrev_5(l) << (logNW + logP) | bitrev_logNW(w) << logP | s  (N = T here), i.e. a lane's
// smallest key is its FIRST slot with the maximum.
template <int P, int NW, bool WEIGHTED>
__global__ void __launch_bounds__(32 * NW)
fps_warp_kernel(const float* __restrict__ xyz, const float* __restrict__ weights, float* __restrict__ temp_io,
                int32_t* __restrict__ idx_out, int M) {
    constexpr int N = 32 * P * NW;
    constexpr int LOG2P = P == 32 ? 5 : P == 16 ? 4 : P == 8 ? 3 : P == 4 ? 2 : P == 2 ? 1 : -1;
    constexpr int LOG2W = NW == 1 ? 0 : NW == 2 ? 1 : NW == 4 ? 2 : NW == 8 ? 3 : -1;
    static_assert(LOG2P > 0 && LOG2W >= 0, "P = 2..32, NW = 1, 2, 4 or 8");
    __shared__ float s_cloud[3 * N];
    __shared__ __align__(16) unsigned long long s_pkt[2][NW > 1 ? NW : 2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, b = blockIdx.x;
    xyz += (size_t)b * N * 3;
    if (WEIGHTED) weights += (size_t)b * N;
    if (temp_io) temp_io += (size_t)b * N;
    idx_out += (size_t)b * M;
    for (int i = threadIdx.x; i < 3 * N; i += 32 * NW) s_cloud[i] = xyz[i];
    __syncthreads();
    auto k_of_slot = [&](int l, int w, int s) -> int {
        return l + 32 * (w + NW * (int)(__brev((unsigned)s) >> (32 - LOG2P)));
    };
    float px[P], py[P], pz[P], pt[P], pw[WEIGHTED ? P : 1];
#pragma unroll
    for (int s = 0; s < P; ++s) {
        const int k = k_of_slot(lane, warp, s);
        px[s] = s_cloud[k * 3 + 0]; py[s] = s_cloud[k * 3 + 1]; pz[s] = s_cloud[k * 3 + 2];
        if (WEIGHTED) pw[s] = weights[k];
        pt[s] = temp_io ? temp_io[k] : 1e10f;
    }
    const unsigned bw = LOG2W ? (__brev((unsigned)warp) >> (32 - (LOG2W ? LOG2W : 1))) : 0u;
    const unsigned lanekey = ((__brev((unsigned)lane) >> 27) << (LOG2W + LOG2P)) | (bw << LOG2P);
    float x1 = s_cloud[0], y1 = s_cloud[1], z1 = s_cloud[2];
    if (threadIdx.x == 0) idx_out[0] = 0;
#pragma unroll 1
    for (int j = 1; j < M; ++j) {
        const f32x2_t cx2 = f2_pack(x1, x1), cy2 = f2_pack(y1, y1), cz2 = f2_pack(z1, z1);
        float mm[P / 2];
#pragma unroll
        for (int s = 0; s < P; s += 2) {
            f32x2_t d2 = fps_dist2(f2_pack(px[s], px[s + 1]), f2_pack(py[s], py[s + 1]), f2_pack(pz[s], pz[s + 1]), cx2, cy2, cz2);
            if (WEIGHTED) d2 = f2_mul(f2_pack(pw[s], pw[s + 1]), d2);
            float d0, d1;
            f2_unpack(d2, d0, d1);
            pt[s] = fminf(d0, pt[s]);
            pt[s + 1] = fminf(d1, pt[s + 1]);
            mm[s / 2] = fmaxf(pt[s], pt[s + 1]);
        }
#pragma unroll
        for (int w = P / 4; w >= 1; w >>= 1)
#pragma unroll
            for (int i = 0; i < w; ++i) mm[i] = fmaxf(mm[i], mm[i + w]);
        const float m = mm[0];
        const unsigned om = WEIGHTED ? hrn_ford(m) : (__float_as_uint(m) ^ 0x80000000u);
        unsigned wm = __reduce_max_sync(0xffffffffu, om);
        unsigned sb = P - 1;
#pragma unroll
        for (int s = P - 2; s >= 0; --s)
            if (pt[s] == m) sb = (unsigned)s;
        unsigned kmin = __reduce_min_sync(0xffffffffu, (om == wm) ? (lanekey | sb) : 0xffffffffu);
        if (NW > 1) {
            // one packet per warp, one barrier, everybody picks the best of the NW packets
            const int par = j & 1;
            if (lane == 0) s_pkt[par][warp] = ((unsigned long long)wm << 32) | kmin;
            asm volatile("bar.sync 1, %0;" ::"n"(32 * NW) : "memory");
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const unsigned long long pk = s_pkt[par][w];
                const unsigned o = (unsigned)(pk >> 32), kk = (unsigned)pk;
                if (w == 0 || o > wm || (o == wm && kk < kmin)) { wm = o; kmin = kk; }
            }
        }
        const int wl = (int)(__brev(kmin >> (LOG2W + LOG2P)) >> 27);
        const int ww = LOG2W ? (int)(__brev((kmin >> LOG2P) & (unsigned)(NW - 1)) >> (32 - (LOG2W ? LOG2W : 1))) : 0;
        const int kw = k_of_slot(wl, ww, (int)(kmin & (unsigned)(P - 1)));
        x1 = s_cloud[kw * 3 + 0]; y1 = s_cloud[kw * 3 + 1]; z1 = s_cloud[kw * 3 + 2];
        if (threadIdx.x == 0) idx_out[j] = kw;
    }
    if (temp_io) {
#pragma unroll
        for (int s = 0; s < P; ++s) temp_io[k_of_slot(lane, warp, s)] = pt[s];
    }
}

template <int P, int NW, bool W>
int launch_fps_warp(const float* xyz, const float* w, float* temp, int32_t* idx, int B, int M, cudaStream_t st) {
    fps_warp_kernel<P, NW, W><<<B, 32 * NW, 0, st>>>(xyz, w, temp, idx, M);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}

// HRN_FPS_CULL=0 in the environment selects the un-culled cluster kernel (A/B measurements)
inline bool fps_cull_enabled() {
    static const bool on = [] { const char* e = getenv("HRN_FPS_CULL"); return !(e && e[0] == '0'); }();
    return on;
}

inline bool fps_warp_enabled() {       // HRN_FPS_WARP=0: the multi-warp cluster kernels also for the small clouds (A/B)
    static const bool on = [] { const char* e = getenv("HRN_FPS_WARP"); return !(e && e[0] == '0'); }();
    return on;
}

inline bool fps_warp_forced() { return getenv("HRN_FPS_WARPS") != nullptr; }
inline int fps_warp_count() {          // HRN_FPS_WARPS = 1, 2, 4 (default) or 8 warps per small cloud (A/B)
    static const int n = [] { const char* e = getenv("HRN_FPS_WARPS"); const int v = e ? atoi(e) : 4; return (v == 1 || v == 2 || v == 8) ? v : 4; }();
    return n;
}

template <bool W>
int dispatch_fps(const float* xyz, const float* w, float* temp, int32_t* idx, int B, int N, int M, cudaStream_t st) {
    // opt_n_threads (cuda_utils.h:22-27): 2^floor(log2 N) clamped to [1,1024]  (defines the tie-break only)
    int log2T = 0;
    while ((2 << log2T) <= N && log2T < 10) ++log2T;
    // all-in-registers cluster kernels.  Few, fat warps: the per-warp bookkeeping of an iteration costs more issue
    // slots than the distance update of 16 points per thread, so 512 threads x 16 points (8192 points per CTA) and the
    // smallest cluster that holds the cloud; clusters of >= 2 CTAs keep the stride a multiple of Tref (MONO).
    // B * CS may exceed the SM count: clusters are independent, the surplus simply runs as a second wave.
    // small power-of-two clouds (N = T: the key is the bit-reversed index): one warp per cloud, no exchange at all
    if (fps_warp_enabled()) {
        const int nw = fps_warp_count();
        if (N == 1024) {                    // measured (64 clouds, 1024 -> 512): 1 / 2 / 4 / 8 warps = 152.9 / 162.8 / 126.1 / 177.3 us
            if (nw == 1) return launch_fps_warp<32, 1, W>(xyz, w, temp, idx, B, M, st);
            if (nw == 2) return launch_fps_warp<16, 2, W>(xyz, w, temp, idx, B, M, st);
            if (nw == 8) return launch_fps_warp<4, 8, W>(xyz, w, temp, idx, B, M, st);
            return launch_fps_warp<8, 4, W>(xyz, w, temp, idx, B, M, st);
        }
        if (N == 512) {                     // measured (64 clouds, 512 -> 256): 1 / 2 / 4 / 8 warps = 64.5 / 60.3 / 64.4 / 74.6 us
            if (nw == 1) return launch_fps_warp<16, 1, W>(xyz, w, temp, idx, B, M, st);
            if (nw == 8) return launch_fps_warp<2, 8, W>(xyz, w, temp, idx, B, M, st);
            if (nw == 4 && fps_warp_forced()) return launch_fps_warp<4, 4, W>(xyz, w, temp, idx, B, M, st);
            return launch_fps_warp<8, 2, W>(xyz, w, temp, idx, B, M, st);
        }
        if (N == 256) return launch_fps_warp<8, 1, W>(xyz, w, temp, idx, B, M, st);
    }
    if (N <= 1024) {
        if (N <= 256) return launch_fps_cluster<256, 1, W>(xyz, w, temp, idx, B, N, M, log2T, 1, st);
        if (N <= 512) return launch_fps_cluster<256, 2, W>(xyz, w, temp, idx, B, N, M, log2T, 1, st);
        return launch_fps_cluster<256, 4, W>(xyz, w, temp, idx, B, N, M, log2T, 1, st);
    }
    if (N <= 2048) return launch_fps_cluster<512, 4, W>(xyz, w, temp, idx, B, N, M, log2T, 1, st);
    if (N <= 4096) return launch_fps_cluster<512, 8, W>(xyz, w, temp, idx, B, N, M, log2T, 1, st);
    if (N <= 8192) return launch_fps_cluster<512, 8, W>(xyz, w, temp, idx, B, N, M, log2T, 2, st);
    if (W) {                                            // 5 registers per point: 4096 points per CTA
        if (N <= 16384) return launch_fps_cluster<512, 8, true>(xyz, w, temp, idx, B, N, M, log2T, 4, st);
        if (N <= 32768) return launch_fps_cluster<512, 8, true>(xyz, w, temp, idx, B, N, M, log2T, 8, st);
    } else {
        // (8 warps x 32 points -- two residues per thread, <256, 16, false, 2> -- is 4 % faster alone, 794 vs 827 us at 64
        // clouds, but not inside the step, where the Morton sort runs beside it: 6.30 vs 6.32 ms)
        // spatially culled variant: the whole cloud, the cell histogram and this CTA's half of the sorted order fit in shared memory
        if (N > 8192 && N <= 16384 && fps_cull_enabled()) return launch_fps_cull<512, 16>(xyz, temp, idx, B, N, M, log2T, st);
        if (N <= 16384) return launch_fps_cluster<512, 16, false>(xyz, w, temp, idx, B, N, M, log2T, 2, st);
        if (N <= 32768) return launch_fps_cluster<512, 16, false>(xyz, w, temp, idx, B, N, M, log2T, 4, st);
        if (N <= 65536) return launch_fps_cluster<512, 16, false>(xyz, w, temp, idx, B, N, M, log2T, 8, st);
        if (N <= 131072) {                              // 16-CTA clusters (non-portable size); else the streaming kernel below
            const int rc = launch_fps_cluster<512, 16, false>(xyz, w, temp, idx, B, N, M, log2T, 16, st);
            if (rc == HRN_OK) return rc;
            (void)cudaGetLastError();
        }
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
