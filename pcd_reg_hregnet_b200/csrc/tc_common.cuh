// Device helpers shared by the tcgen05 kernels: mbarrier, bulk copy, UMMA descriptors / issue, TMEM loads,
// bf16 hi/lo operand splitting, activations.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>
#include "rows.cuh"

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
// same, for waits that are expected to be long (ring-full back-pressure): back off between polls so that the spinning
// warp does not take issue slots from the warps on the critical path
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
    uint32_t done;
    for (;;) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(64);
    }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// UMMA shared-memory descriptor, K-major, no swizzle: core matrix = 8 rows x 16 bytes (128 contiguous bytes);
// LBO = byte distance between the two K-halves (core matrices adjacent in K), SBO = between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// The same descriptor as  fixed part | 14-bit start address: the issuing thread keeps the fixed part in registers and
// only adds to the address field (its instruction stream, not the tensor pipe, limits small MMAs).
__device__ __forceinline__ uint64_t umma_desc_fixed(uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
        "%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// The two transcendental activations (sigma heads: softplus + eps, confidence heads: sigmoid) are ONE out-of-line copy: inlined
// into loops unrolled over the 32 columns of an accumulator block they were ~2 K instructions per kernel, and the short head
// launches stalled on instruction fetches (ncu: no_inst).  Same arithmetic, same bits.
static __device__ __noinline__ float act_fn_slow(float v, int act) {
    if (act == HRN_ACT_SOFTPLUS_EPS) return (v > 20.f ? v : log1pf(expf(v))) + 0.001f;
    if (act == HRN_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
    return v;
}
__device__ __forceinline__ float act_fn(float v, int act) {
    if (act == HRN_ACT_RELU) return fmaxf(v, 0.f);
    if (act == HRN_ACT_NONE) return v;
    return act_fn_slow(v, act);
}

// bf16 hi / lo split of a pair: hi = rn_bf16(x) packed, lo = rn_bf16(x - float(hi)) packed; the two residuals are taken
// with one packed fp32x2 subtract (each half rounds like the scalar FADD)
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
    const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h);
    // bf16 -> fp32 is a 16-bit shift: low half << 16, high half masked in place
    const unsigned long long hf = ((unsigned long long)(hb & 0xffff0000u) << 32) | (unsigned long long)(hb << 16);
    unsigned long long xv, rv;
    asm("mov.b64 %0, {%1, %2};" : "=l"(xv) : "f"(x0), "f"(x1));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(rv) : "l"(xv), "l"(hf));
    float r0, r1;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(rv));
    const __nv_bfloat162 l = __floats2bfloat162_rn(r0, r1);
    hi = hb;
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// split 8 floats into bf16 hi / lo and store both 16-byte core-matrix rows.  The residuals x - float(hi) are taken two
// at a time with the packed fp32x2 subtract (each half rounds like the scalar FADD): the conversion loops are bound by
// their instruction count.
__device__ __forceinline__ void split_store8(const float (&x)[8], uint4* dst_hi, uint4* dst_lo) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_pair(x[2 * i], x[2 * i + 1], hi[i], lo[i]);
    *dst_hi = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *dst_lo = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// relu(x) of a pair as bf16 hi / lo with the ReLU folded into the two conversions: hi = rz_bf16(max(x, 0)),
// lo = rn_bf16(max(x - hi, 0)).  hi is TRUNCATED, so the residual of a non-negative x is non-negative; for x < 0
// hi = 0 and the residual x clamps to 0.  |relu(x) - hi - lo| <= 2^-16 |x|.
__device__ __forceinline__ void split_pair_relu(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    uint32_t hb;
    asm("cvt.rz.relu.bf16x2.f32 %0, %1, %2;" : "=r"(hb) : "f"(x1), "f"(x0));
    const unsigned long long hf = ((unsigned long long)(hb & 0xffff0000u) << 32) | (unsigned long long)(hb << 16);
    unsigned long long xv, rv;
    asm("mov.b64 %0, {%1, %2};" : "=l"(xv) : "f"(x0), "f"(x1));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(rv) : "l"(xv), "l"(hf));
    float r0, r1;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(rv));
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
    hi = hb;
}
__device__ __forceinline__ void split_store8_relu(const float* x, uint4* dst_hi, uint4* dst_lo) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_pair_relu(x[2 * i], x[2 * i + 1], hi[i], lo[i]);
    *dst_hi = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *dst_lo = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// ---- precision mode 1: single-pass fp16 operands (one MMA per MAC) ---------------------------------------------------------
// Used where the 1e-3 feature gate leaves room for an 11-bit significand (the correspondence stages: measured
// profiles/r02_precision_emulation.txt); everywhere else the operands are bf16 hi + lo (mode 3, three MMAs per MAC).
// {lo, hi} -> packed f16x2, round to nearest, saturating (a value beyond 65504 stays finite instead of becoming inf)
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// same with ReLU folded into the conversion (F2FP.SATFINITE.RELU: one instruction for two activations)
__device__ __forceinline__ uint32_t pack_f16x2_relu(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// UMMA instruction descriptor, kind::f16, fp32 accumulate, M = 128: operand format bf16 (PREC 3) or f16 (PREC 1)
template <int PREC>
__device__ __forceinline__ uint32_t umma_idesc_m128(int N) {
    const uint32_t fmt = PREC == 3 ? ((1u << 7) | (1u << 10)) : 0u;
    return (1u << 4) | fmt | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
