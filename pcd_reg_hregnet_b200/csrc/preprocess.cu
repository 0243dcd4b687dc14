// Input pipeline step in front of the registration path: what the reference does per sample with
// numpy in DataLoader workers, on the device for a whole batch of sweeps.
//
//   hrn_range_filter   dataset/dataset_utils.py:113-125 (remove_points_by_range): keep the points with ||p|| < max_range,
//                      ORDER PRESERVED (boolean-mask indexing) -> a stable stream compaction, two passes (count per
//                      4096-point segment, then write) so that every sweep is spread over many SMs.  The predicate is evaluated
//                      exactly like numpy's float32 np.linalg.norm(axis=1): sqrt((x*x + y*y) + z*z), every operation
//                      rounded separately (no fma), so the kept set is bit-identical.
//   hrn_se3_exp        transform/rodrigues.py:526-550 (SE3.exp): twist (w, v) -> [R p; 0 1] with the reference's sinc1 /
//                      sinc2 / sinc3 (Taylor series below |t| < 0.01, rodrigues.py:6-18,100-112,132-144); used by
//                      transform/dataset_transforms.py:128-140 to perturb the source cloud.
// Fixed-size resampling (dataset_utils.py:188-223) is a row gather with the (host- or device-drawn) index list:
// hrn_gather_rows; applying the perturbation is hrn_transform_points.
#include "common.cuh"

namespace {

constexpr int RF_THREADS = 1024;
constexpr int RF_SEG = 4096;                       // points per CTA (4 chunks of 1024)

__device__ __forceinline__ bool rf_keep(const float* __restrict__ xyz, long long r, float max_range, float& x, float& y, float& z) {
    x = xyz[r * 3]; y = xyz[r * 3 + 1]; z = xyz[r * 3 + 2];
    // numpy float32 norm: every operation rounded on its own (no fma contraction)
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z))) < max_range;
}

// Sweeps are concatenated, sweep s = rows [offs[s], offs[s+1]); CTA (seg, s) owns RF_SEG consecutive points of sweep s.
// Pass 1: kept points per segment.  Pass 2: base of the segment = sum of the earlier segments of its sweep, then a
// block-wide stable compaction (ballot / popc inside a warp, one scan of the 32 warp totals per 1024-point chunk).
__global__ void __launch_bounds__(RF_THREADS)
rf_count_kernel(const float* __restrict__ xyz, const long long* __restrict__ offs, float max_range, int* __restrict__ seg_count) {
    __shared__ int s_warp[32];
    const int s = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long r0 = offs[s] + (long long)blockIdx.x * RF_SEG, r1 = min(offs[s + 1], r0 + RF_SEG);
    int n = 0;
    for (long long r = r0 + tid; r < r1; r += RF_THREADS) {
        float x, y, z;
        n += rf_keep(xyz, r, max_range, x, y, z) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if (lane == 0) s_warp[warp] = n;
    __syncthreads();
    if (warp == 0) {
        int t = s_warp[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) seg_count[(size_t)s * gridDim.x + blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(RF_THREADS)
rf_write_kernel(const float* __restrict__ xyz, const float* __restrict__ inten, const long long* __restrict__ offs,
                float max_range, const int* __restrict__ seg_count, float* __restrict__ xyz_out,
                float* __restrict__ inten_out, int* __restrict__ count) {
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int s = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long sweep0 = offs[s], sweep1 = offs[s + 1];
    const long long r0 = sweep0 + (long long)blockIdx.x * RF_SEG, r1 = min(sweep1, r0 + RF_SEG);
    const int n_seg = (int)((sweep1 - sweep0 + RF_SEG - 1) / RF_SEG);
    if (warp == 0) {                                   // kept points in the earlier segments of this sweep
        int b = 0;
        for (int j = lane; j < (int)blockIdx.x && j < n_seg; j += 32) b += seg_count[(size_t)s * gridDim.x + j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
        if (lane == 0) s_base = b;
    }
    __syncthreads();
    if (r0 >= sweep1 && !(blockIdx.x == 0 && sweep1 == sweep0)) return;     // segment beyond this sweep
    for (long long c = r0; c < r1; c += RF_THREADS) {
        const long long r = c + tid;
        bool keep = false;
        float x = 0.f, y = 0.f, z = 0.f;
        if (r < r1) keep = rf_keep(xyz, r, max_range, x, y, z);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        const int pre = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        const int base = s_base;
        int wv = s_warp[lane], incl = wv;                      // every warp scans the 32 warp totals
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int wbase = __shfl_sync(0xffffffffu, incl - wv, warp);
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (keep) {
            const long long o = sweep0 + base + wbase + pre;   // compacted into the sweep's own row range
            xyz_out[o * 3] = x; xyz_out[o * 3 + 1] = y; xyz_out[o * 3 + 2] = z;
            if (inten_out) inten_out[o] = inten[r];
        }
        __syncthreads();
        if (tid == 0) s_base = base + total;
        __syncthreads();
    }
    if (tid == 0 && ((int)blockIdx.x == n_seg - 1 || n_seg == 0)) count[s] = s_base;
}

// Fixed-size resampling of a batch of filtered sweeps in one launch (dataset_utils.py:188-223): out[b,i] = sweep_b[j] with
// j = i for i < count[b] when the sweep has <= n points (the cloud itself, then the padding picks), idx[b,i] otherwise.
__global__ void resample_gather_kernel(const float* __restrict__ xyz, const long long* __restrict__ offs,
                                       const int* __restrict__ count, const int* __restrict__ idx, int n,
                                       float* __restrict__ out) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = count[b];
    int j = (c <= n && i < c) ? i : idx[(size_t)b * n + i];
    j = min(max(j, 0), max(c - 1, 0));
    const float* p = xyz + (offs[b] + j) * 3;
    float* o = out + ((size_t)b * n + i) * 3;
    if (c > 0) { o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; } else { o[0] = 0.f; o[1] = 0.f; o[2] = 0.f; }
}

__device__ __forceinline__ float sinc1f(float t) {
    if (fabsf(t) < 0.01f) { const float t2 = t * t; return 1.f - t2 / 6.f * (1.f - t2 / 20.f * (1.f - t2 / 42.f)); }
    return sinf(t) / t;
}
__device__ __forceinline__ float sinc2f(float t) {
    const float t2 = t * t;
    if (fabsf(t) < 0.01f) return 0.5f * (1.f - t2 / 12.f * (1.f - t2 / 30.f * (1.f - t2 / 56.f)));
    return (1.f - cosf(t)) / t2;
}
__device__ __forceinline__ float sinc3f(float t) {
    if (fabsf(t) < 0.01f) { const float t2 = t * t; return (1.f / 6.f) * (1.f - t2 / 20.f * (1.f - t2 / 42.f * (1.f - t2 / 72.f))); }
    return (t - sinf(t)) / (t * t * t);
}

__global__ void se3_exp_kernel(const float* __restrict__ twist, int B, float* __restrict__ g) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float w[3] = {twist[b * 6], twist[b * 6 + 1], twist[b * 6 + 2]}, v[3] = {twist[b * 6 + 3], twist[b * 6 + 4], twist[b * 6 + 5]};
    const float t = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    const float W[9] = {0.f, -w[2], w[1], w[2], 0.f, -w[0], -w[1], w[0], 0.f};
    float S[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) S[i * 3 + j] = W[i * 3] * W[j] + W[i * 3 + 1] * W[3 + j] + W[i * 3 + 2] * W[6 + j];
    const float s1 = sinc1f(t), s2 = sinc2f(t), s3 = sinc3f(t);
    float* o = g + (size_t)b * 16;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float p = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float I = i == j ? 1.f : 0.f;
            o[i * 4 + j] = I + s1 * W[i * 3 + j] + s2 * S[i * 3 + j];
            p += (I + s2 * W[i * 3 + j] + s3 * S[i * 3 + j]) * v[j];
        }
        o[i * 4 + 3] = p;
    }
    o[12] = 0.f; o[13] = 0.f; o[14] = 0.f; o[15] = 1.f;
}

}  // namespace

// xyz [total,3], intensity [total] (nullable) = n_sweeps concatenated sweeps, offsets [n_sweeps+1] (device, int64),
// max_sweep_points >= the largest sweep (host value: sizes the grid), scratch [n_sweeps * ceil(max_sweep_points / 4096)]
// int32.  Kept points of sweep s are written, in their original order, to rows [offsets[s], offsets[s] + count[s]) of
// xyz_out / intensity_out (same layout as the input; in-place operation is NOT supported); count [n_sweeps] int32.
HRN_API int hrn_range_filter(const float* xyz, const float* intensity, const long long* offsets, int n_sweeps,
                             long long max_sweep_points, float max_range, float* xyz_out, float* intensity_out, int* count,
                             int* scratch, void* stream) {
    if (!xyz || !offsets || !xyz_out || !count || !scratch || n_sweeps < 0 || max_sweep_points < 0 || (intensity_out && !intensity) ||
        xyz == xyz_out)
        return HRN_ERR_BAD_ARG;
    if (n_sweeps == 0) return HRN_OK;
    long long segs = (max_sweep_points + RF_SEG - 1) / RF_SEG;
    if (segs < 1) segs = 1;
    if (segs > 0x7fffffffLL || n_sweeps > 65535) return HRN_ERR_UNSUPPORTED;
    const dim3 grid((unsigned)segs, (unsigned)n_sweeps);
    cudaStream_t st = (cudaStream_t)stream;
    rf_count_kernel<<<grid, RF_THREADS, 0, st>>>(xyz, offsets, max_range, scratch);
    HRN_LAUNCH_CHECK();
    rf_write_kernel<<<grid, RF_THREADS, 0, st>>>(xyz, intensity, offsets, max_range, scratch, xyz_out, intensity_out, count);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}

// twist [B,6] = (w, v) -> g [B,16] row-major 4x4.
HRN_API int hrn_se3_exp(const float* twist, int B, float* g, void* stream) {
    if (!twist || !g || B < 0) return HRN_ERR_BAD_ARG;
    if (B == 0) return HRN_OK;
    se3_exp_kernel<<<hrn_divup(B, 128), 128, 0, (cudaStream_t)stream>>>(twist, B, g);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}

// xyz [total,3] filtered sweeps (hrn_range_filter layout), offsets [B+1], count [B], idx [B,n] int32 -> out [B,n,3].
HRN_API int hrn_resample_gather(const float* xyz, const long long* offsets, const int* count, const int* idx, int B, int n,
                                float* out, void* stream) {
    if (!xyz || !offsets || !count || !idx || !out || B < 0 || n <= 0) return HRN_ERR_BAD_ARG;
    if (B == 0) return HRN_OK;
    resample_gather_kernel<<<dim3(hrn_divup(n, 256), B), 256, 0, (cudaStream_t)stream>>>(xyz, offsets, count, idx, n, out);
    HRN_LAUNCH_CHECK();
    return HRN_OK;
}
