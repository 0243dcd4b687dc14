// "Rows" view shared by the shared-MLP kernels: a [rows, K] activation matrix that is never materialised.
// It is the concatenation, along K, of up to 4 column segments, each read from a channels-last matrix
//   DIRECT    : source row = r
//   BROADCAST : source row = r / group          (the reference's  x.unsqueeze(2).repeat(1,1,k,1), layers.py:281-282)
//   GATHER    : source row = b*src_rows_per_batch + gather_idx[r],  b = r / rows_per_batch
//                                                (the reference's  knn_gather(x, idx), layers.py:25,279,437)
// optionally multiplied by a per-row scalar (the attention weight, layers.py:157-158).
// The struct mirrors hrn_rows_t in include/hregnet_b200.h byte for byte.
#pragma once
#include <stdint.h>

enum { HRN_SEG_DIRECT = 0, HRN_SEG_BROADCAST = 1, HRN_SEG_GATHER = 2 };
enum { HRN_ACT_NONE = 0, HRN_ACT_RELU = 1, HRN_ACT_SOFTPLUS_EPS = 2, HRN_ACT_SIGMOID = 3 };

struct hrn_seg_t {
    const float* ptr;        // [src_rows, ld] row-major
    const float* row_scale;  // nullable, [rows]
    int32_t channels;        // columns taken from this source
    int32_t ld;              // leading dimension of the source (floats)
    int32_t col0;            // first source column
    int32_t mode;            // HRN_SEG_*
};

struct hrn_rows_t {
    hrn_seg_t seg[4];
    const int32_t* gather_idx;   // [rows], for GATHER segments
    int32_t n_seg;
    int32_t group;               // rows per group (k), for BROADCAST segments
    int32_t rows_per_batch;      // M*k
    int32_t src_rows_per_batch;  // N
};

#ifdef __CUDACC__
__device__ __forceinline__ long long hrn_src_row(const hrn_rows_t& in, int mode, long long r) {
    if (mode == HRN_SEG_DIRECT) return r;
    if (mode == HRN_SEG_BROADCAST) return r / in.group;
    const long long b = r / in.rows_per_batch;
    return b * in.src_rows_per_batch + in.gather_idx[r];
}
#endif
