/*
 * hregnet_b200.h -- C ABI of libhregnet_b200.so: hand-written sm_100a CUDA kernels for the HRegNet
 * registration forward path (reference: UpendraArun/pcd_reg_hregnet).
 *
 * Conventions
 *   - extern "C", plain device pointers + sizes; no torch / ATen types.  `stream` is a cudaStream_t passed as void*.
 *   - every buffer is CALLER-ALLOCATED device memory, contiguous, fp32 unless noted (the reference's wrappers
 *     also leave allocation to the caller: models/utils.py:24-25,73,84).
 *   - asynchronous on `stream`, no host synchronisation, no global state (re-entrant per stream).
 *   - return 0 on success, a cudaError_t (< 1000) or HRN_ERR_* (>= 1000).  Never exits the process (the reference
 *     kernels launchers call exit(-1): furthest_point_sampling_gpu.cu:34-38,247-251).
 *
 * Each entry point cites the reference interface it replaces (paths relative to the reference repo root).
 * The binding a reference maintainer would add is shown in INTEGRATION.md.
 */
#ifndef HREGNET_B200_H
#define HREGNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HRN_OK 0
#define HRN_ERR_BAD_ARG 1001
#define HRN_ERR_UNSUPPORTED 1002

/* ---------------------------------------------------------------------------------------------------------
 * 1. models/PointUtils native module (pybind `point_utils_cuda`, point_utils_api.cpp:6-13)
 * ------------------------------------------------------------------------------------------------------- */

/* Replaces furthest_point_sampling_wrapper (furthest_point_sampling.cpp:33-43, kernel .cu:84-252) and, with
 * `weights != NULL`, weighted_furthest_point_sampling_wrapper (.cpp:45-56, kernel .cu:254-419).
 *   xyz [B,N,3]; weights [B,N] or NULL; temp [B,N] scratch: if non-NULL its contents are the initial
 *   min-distances (the reference's callers fill 1e10, models/utils.py:25,49) and it receives the final ones;
 *   NULL = start from 1e10 and keep the array on chip (requires N <= 16384).  idx [B,M] int32 out.
 * Bit-exact with the reference kernels (rounding fma(dz,dz,fma(dx,dx,dy*dy)), tie-break = reference's
 * block-reduction order for T = opt_n_threads(N), cuda_utils.h:22-27). */
int hrn_fps(const float* xyz, const float* weights, float* temp, int32_t* idx, int B, int N, int M, void* stream);

/* Replaces gather_points_wrapper (furthest_point_sampling.cpp:10-19, kernel .cu:7-39).
 *   points [B,C,N], idx [B,M] int32 -> out [B,C,M]. */
int hrn_gather_points(const float* points, const int32_t* idx, float* out, int B, int C, int N, int M, void* stream);

/* Replaces gather_points_grad_wrapper (furthest_point_sampling.cpp:21-31, kernel .cu:41-73).
 *   grad_out [B,C,M], idx [B,M] -> grad_points [B,C,N] (caller pre-zeroes; atomically accumulated). */
int hrn_gather_points_grad(const float* grad_out, const int32_t* idx, float* grad_points, int B, int C, int N, int M,
                           void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * 2. pytorch3d.ops (third-party; imported at models/HRegNet/layers.py:7)
 * ------------------------------------------------------------------------------------------------------- */

/* Replaces pytorch3d.ops.knn_points(p1, p2, K=K, return_nn=...) -- call sites layers.py:20,278,316,322,434.
 *   p1 [B,M,D] queries, p2 [B,N,D] references.  If q_idx [B,M] int32 is non-NULL (D must be 3) the queries are
 *   p2[b, q_idx[b,m], :] and p1 is ignored (fuses layers.py:139-143 into the search); q_out [B,M,3] receives them.
 *   Outputs, each nullable: dists [B,M,K] squared L2 ascending; idx64 [B,M,K] int64 (pytorch3d dtype);
 *   idx32 [B,M,K] int32 (internal consumers); nn [B,M,K,D] = p2[idx].
 *   Contract: dist accumulated d = 0..D-1 with fma in fp32; order (dist asc, index asc).  K <= min(N, 64). */
int hrn_knn(const float* p1, const int32_t* q_idx, const float* p2, int B, int M, int N, int D, int K, float* dists,
            int64_t* idx64, int32_t* idx32, float* nn, float* q_out, void* stream);

/* Same contract and bit-identical results as hrn_knn for D = 3, 128 <= N <= 32768 (clouds below 1024 points are padded to 1024 sorted positions), with spatial culling: the
 * cloud is Morton-sorted once per call (scratch_pts [B*N2] float4, scratch_boxes [B*(N2/32)*6] floats, N2 = next
 * power of two >= N, caller-allocated) and each query opens only the 32-point boxes whose fp32 lower bound can still
 * beat its current K-th distance. */
int hrn_knn3_sorted(const float* p1, const int32_t* q_idx, const float* p2, int B, int M, int N, int K, void* scratch_pts,
                    float* scratch_boxes, float* dists, int64_t* idx64, int32_t* idx32, float* nn, float* q_out,
                    void* stream);

/* The two halves of hrn_knn3_sorted on their own (same argument rules).  The sort depends on the reference cloud only:
 * a caller may run it on a second stream while the queries are still being chosen (furthest-point sampling), then
 * search on the stream that holds the queries once both are done. */
int hrn_knn3_sort(const float* p2, int B, int N, void* scratch_pts, float* scratch_boxes, void* stream);
int hrn_knn3_search(const float* p1, const int32_t* q_idx, const float* p2, int B, int M, int N, int K,
                    const void* sorted_pts, const float* sorted_boxes, float* dists, int64_t* idx64, int32_t* idx32,
                    float* nn, float* q_out, void* stream);

/* Replaces pytorch3d.ops.knn_gather(x, idx) -- call sites layers.py:25,279,288,303,309,317,323,352,358,437,443.
 *   x [B,N,U], idx [B,M,K] int64 -> out [B,M,K,U]. */
int hrn_knn_gather(const float* x, const int64_t* idx, float* out, int B, int N, int M, int K, int U, void* stream);

/* Backward passes of the two pytorch3d stand-ins (pytorch3d's autograd functions provide them to the
 * reference's training scripts, train/train_reg_v0.py:281-296):
 *   hrn_knn_gather_grad: grad_x [B,N,U] (pre-zeroed) += scatter of grad_out [B,M,K,U] through idx [B,M,K] (fp32 atomics);
 *   hrn_knn_dists_grad:  for dists[b,m,k] = |p1[b,m] - p2[b,idx]|^2 and grad_dists [B,M,K]:
 *                        grad_p1 [B,M,D] (nullable) = sum_k 2 g (p1 - p2[idx]), grad_p2 [B,N,D] (nullable, pre-zeroed)
 *                        -= 2 g (p1 - p2[idx]) (fp32 atomics). */
int hrn_knn_gather_grad(const float* grad_out, const int64_t* idx, float* grad_x, int B, int N, int M, int K, int U,
                        void* stream);
int hrn_knn_dists_grad(const float* p1, const float* p2, const int64_t* idx, const float* grad_dists, float* grad_p1,
                       float* grad_p2, int B, int M, int N, int D, int K, void* stream);

/* Row gather with int32 indices: out[b,m,:] = x[b, idx[b,m], :]  (the gather_operation(x^T, idx)^T idiom of
 * layers.py:140,143 without its two permute+contiguous passes). */
int hrn_gather_rows(const float* x, const int32_t* idx, float* out, int B, int N, int M, int U, void* stream);

/* [B,R,C] -> [B,C,R]: the permute(0,2,1).contiguous() between the reference's [B,C,N] API layout and the
 * channels-last rows used internally (layers.py:27,274-275,435-436). */
int hrn_transpose(const float* in, float* out, int B, int R, int C, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * 3. Shared-MLP building blocks (models/HRegNet/layers.py KeypointDetector / DescExtractor / CoarseReg / FineReg)
 * ------------------------------------------------------------------------------------------------------- */

enum { HRN_SEG_DIRECT = 0, HRN_SEG_BROADCAST = 1, HRN_SEG_GATHER = 2 };
enum { HRN_ACT_NONE = 0, HRN_ACT_RELU = 1, HRN_ACT_SOFTPLUS_EPS = 2, HRN_ACT_SIGMOID = 3 };

/* One column segment of a virtual [rows, K] activation matrix (see pcd_reg_hregnet_b200/csrc/rows.cuh). */
typedef struct {
    const float* ptr;        /* [src_rows, ld] row-major (channels-last) */
    const float* row_scale;  /* NULL, or [rows] per-row multiplier (attention weights, layers.py:157-158) */
    int32_t channels;        /* columns taken */
    int32_t ld;              /* leading dimension in floats */
    int32_t col0;            /* first source column */
    int32_t mode;            /* HRN_SEG_DIRECT: src row r; BROADCAST: r / group (layers.py:281-282 repeat);
                                GATHER: b*src_rows_per_batch + gather_idx[r] (knn_gather, layers.py:25,279,437) */
} hrn_seg_t;

typedef struct {
    hrn_seg_t seg[4];
    const int32_t* gather_idx;   /* [rows] */
    int32_t n_seg;               /* 1..4 */
    int32_t group;               /* k */
    int32_t rows_per_batch;      /* M*k */
    int32_t src_rows_per_batch;  /* N */
} hrn_rows_t;

/* Y[r,n] = act( sum_k X[r,k] W[n,k] + bias[n] ), X = concatenation of in->seg[*]; W [Cout,K] row-major with
 * BatchNorm(eval) folded in.  Replaces Conv2d/Conv1d(1x1)+BatchNorm+ReLU triples (layers.py:118-121,186-198,
 * 249-268,420-431) and the cat/repeat/knn_gather/permute tensors feeding them.  Exact-fp32 CUDA-core variant. */
int hrn_layer_fp32(const hrn_rows_t* in, const float* W, const float* bias, int act, float* Y, int ldy,
                   long long rows, int Cout, void* stream);

/* Tensor-core variant of hrn_layer_fp32 (tcgen05.mma kind::f16, bf16 hi/lo split of both operands, three products
 * per MAC, fp32 accumulation in TMEM; relative error ~2^-16).  Wp = weights pre-split / pre-tiled on the host as
 * [n_stage][hi|lo][4][NP][8] bf16 (pcd_reg_hregnet_b200/engine_tc.py), NP = Cout padded to a multiple of 16
 * (512 if > 256), K padded per segment to a multiple of 8 and in total to 32 * n_stage.
 * prec = 3: as above.  prec = 1: single-pass fp16 operands (one product per MAC, relative error ~2^-11 per layer), Wp =
 * [n_stage][4][NP][8] fp16 -- for the stages whose tolerance allows it (the correspondence stages; measured deltas in
 * profiles/r02_precision_emulation.txt); 16-byte aligned segments only. */
int hrn_layer_tc(const hrn_rows_t* in, const void* Wp, const float* bias, int act, float* Y, int ldy, long long rows,
                 int Cout, int NP, int n_stage, int prec, void* stream);

/* The same layer with the reference's max over the k neighbours (x.max(dim=3), layers.py:208) taken in the epilogue:
 * G [rows / k, ldg] = max over each k consecutive rows of act(W x + b); the per-row result is never written.
 * k = 8, 16 or 32; act = none or ReLU; Cout a multiple of 32 (= NP); 16-byte aligned segments, G and bias --
 * HRN_ERR_UNSUPPORTED otherwise (callers then run hrn_layer_tc + hrn_group_max). */
int hrn_layer_tc_groupmax(const hrn_rows_t* in, const void* Wp, const float* bias, int act, float* G, int ldg,
                          long long rows, int Cout, int NP, int n_stage, int k, int prec, void* stream);

/* Two or three fused shared-MLP layers with the activations kept in shared / tensor memory: the reference's conv
 * stacks `convs`, `convs_1`, `convs_2` (layers.py:118-121,249-260,420-423), the descriptor head mlp1+mlp2
 * (layers.py:193-198,207) and the per-keypoint heads mlp1/mlp2/mlp3 (layers.py:124-130,161-163,425-431,451-452) where
 * the widths are <= 256, plus the reduction that consumes them: mode 0 = store Y rows; 1 = G[g,:] = max over the kseg
 * rows of each group (+ Y rows if Y != NULL, layers.py:202,208); 2 = a = softmax_k(max_c Y), G[g,:] = sum_k a*Y, Y
 * (optional) = the rows Y*a (layers.py:150-159,329-332,384-390,446-450).  Hidden activations ReLU, last = `act`.
 * W = K=16 weight pieces of the layers in execution order (pcd_reg_hregnet_b200/engine_tc.pack_chain), bias = b1|b2|b3,
 * n1..n3 = issued widths (multiples of 16; the last = cout padded), chunks0 = 8-wide K chunks of the virtual input.
 * prec = 3: bf16 hi/lo operands (pieces [hi|lo][2][N][8]); prec = 1: single-pass fp16 (pieces [2][N][8] fp16), see hrn_layer_tc.
 * Zb [rows / kseg, ldz] (nullable; needs kseg = 8), Zg [source rows, ldz] (nullable): fp32 rows ADDED to the first layer's
 * pre-activations -- row r takes Zb[r / kseg] + Zg[b * src_rows_per_batch + gather_idx[r]] -- the part of the (linear) first
 * layer that is constant inside a group / depends on the gathered row only, applied once per point by the caller; `in` then
 * only holds the per-row segments (n1 must be a multiple of 32). */
int hrn_chain_tc(const hrn_rows_t* in, const void* W, const float* bias, int nl, int n1, int n2, int n3, int cout, int act,
                 int chunks0, int mode, int kseg, float* Y, int ldy, float* G, float* a, long long rows, int prec,
                 const float* Zb, const float* Zg, int ldz, void* stream);

/* CoarseReg's conv stack convs_1 (528 -> 512 -> 512 -> 512, models/HRegNet/layers.py:364-375) + its attention tail
 * (layers.py:384-390: a = softmax_k(max_c Y), attentive feature = sum_k a Y) on a CLUSTER OF TWO CTAs per 128-row tile: each
 * CTA computes one half of every layer's output columns, the hidden activations travel between the two SMs as 32-column
 * operand blocks through distributed shared memory (cp.async.bulk shared::cta -> shared::cluster) and never reach HBM.
 *   W: per column half (rank 0, rank 1; w_rank_bytes each) the K=16 weight pieces of the three layers, bias: per half
 *   b1|b2|b3, both as laid out by pcd_reg_hregnet_b200/engine_tc.pack_chain_wide; n1..n3 full widths (multiples of 64,
 *   <= 512), chunks0 = 8-wide K chunks of the virtual input (even), kseg = 8, rows % 128 == 0, prec as hrn_chain_tc.
 *   Zb [rows / kseg, ldz], Zg [source rows, ldz] (both or neither): fp32 rows ADDED to the first layer's pre-activations,
 *   row r takes Zb[r / kseg] + Zg[b * src_rows_per_batch + gather_idx[r]] -- the part of the (linear) first layer that
 *   is constant inside a group / depends on the gathered row only, applied once per point by the caller
 *   (engine_tc._split_first_layer); `in` then only holds the per-row segments.
 *   Outputs: G [rows / kseg, n3] attentive feature, a [rows] attention weights (nullable). */
int hrn_chain_wide(const hrn_rows_t* in, const void* W, long long w_rank_bytes, const float* bias, int n1, int n2, int n3,
                   int chunks0, int kseg, const float* Zb, const float* Zg, int ldz, float* G, float* a, long long rows,
                   int prec, void* stream);

/* Per-keypoint confidence head of CoarseReg (512 -> 512 -> 512 -> 1, sigmoid; models/HRegNet/layers.py:391-394) as ONE
 * launch of the same 2-CTA-cluster kernel: Y[r] = act(column 0 of the third layer + bias).  W / bias as for hrn_chain_wide
 * (engine_tc.pack_chain_wide) with the last layer zero-padded to n3 = 64 columns; n1, n2 multiples of 64, <= 512;
 * rows % 128 == 0. */
int hrn_chain_wide_head(const hrn_rows_t* in, const void* W, long long w_rank_bytes, const float* bias, int n1, int n2,
                        int n3, int chunks0, int act, float* Y, long long rows, int prec, void* stream);

/* Level 1 of HierFeatureExtraction (models/HRegNet/models.py:27-28: detector_1 + desc_extractor_1; in_channels 0,
 * k = 64, widths 32/32/64, mlp 192->32->64) as ONE persistent tcgen05 kernel: grouping (layers.py:9-27), the two conv
 * stacks, attention / keypoints / attentive feature (layers.py:150-159) and the descriptor head (layers.py:200-209)
 * with every per-neighbour tensor kept in shared / tensor memory; weights resident in shared memory, four tiles in
 * flight per SM; the repeated max_k(X1) input of mlp1 (layers.py:203-205) enters as a per-keypoint bias evaluated once
 * per keypoint in fp32.  (Levels 2 and 3: hrn_level_ws; any other level returns HRN_ERR_UNSUPPORTED.)
 *   q [B*M,3] sampled keypoints, xyz [B,N,3], feat = NULL, idx [B*M*k] int32;
 *   Wpack (hrn_level_pack_bytes(level) bytes) and biases (hrn_level_bias_count(level) floats) as laid out by
 *   pcd_reg_hregnet_b200/engine_tc.pack_level.  Outputs per keypoint: out_xyz [B*M,3], out_af [B*M,CO], out_desc [B*M,CD]. */
int hrn_level_fused(int level, const float* q, const float* xyz, const float* feat, const int32_t* idx, const void* Wpack,
                    const float* biases, float* out_xyz, float* out_af, float* out_desc, int B, int M, int N, int k,
                    void* stream);
int hrn_level_pack_bytes(int level);
int hrn_level_bias_count(int level);

/* Levels 2 and 3 of HierFeatureExtraction (models/HRegNet/models.py:33-34,39-40; level 3: in_channels 128, k = 16,
 * widths 128/128/256, mlp 768->128->256) as ONE persistent, warp-specialised tcgen05 kernel: same stage as
 * hrn_level_fused, for the levels whose weights have to stream (K=16 pieces through a cp.async.bulk ring); the next
 * layer's MMAs are issued block by block while the current accumulator is still being drained, and the reference's
 * repeated max_k(X1) input of mlp1 (layers.py:203-205) enters as a per-keypoint bias evaluated once per keypoint in fp32.
 *   Wpack (hrn_level_ws_pack_bytes(level) bytes), WaT [2C, C] fp32 and biases (hrn_level_ws_bias_count(level) floats)
 *   as laid out by pcd_reg_hregnet_b200/engine_tc.pack_level_ws: the biases of the conv stacks and of mlp2 travel INSIDE Wpack
 *   (d1 / x1 as the weight column of a constant-1 input channel, d2 x2 d3 x3 mlp2 as one bias piece in front of the layer's
 *   K=16 pieces -- one extra MMA against a block of ones); of `biases` only mlp1's entries are read.  Arguments and outputs
 *   as hrn_level_fused. */
int hrn_level_ws(int level, const float* q, const float* xyz, const float* feat, const int32_t* idx, const void* Wpack,
                 const float* WaT, const float* biases, float* out_xyz, float* out_af, float* out_desc, int B, int M,
                 int N, int k, void* stream);
int hrn_level_ws_pack_bytes(int level);
int hrn_level_ws_bias_count(int level);

/* a[g*k+j] = softmax_j( max_c E[g*k+j, c] )   (layers.py:151-152,330-331,385-386,447-448).  k <= 64. */
int hrn_group_attention(const float* E, int ldE, int C, long long groups, int k, float* a, void* stream);

/* out[g,c] = sum_j a[g*k+j] * V[row, c], row = g*k+j (idx NULL) or b*N + idx[g*k+j]
 * (layers.py:154-159,332,337,388-390,449-450). */
int hrn_group_weighted_sum(const float* a, const float* V, int ldV, int C, long long groups, int k, const int32_t* idx,
                           int groups_per_batch, int N, float* out, int ldo, void* stream);

/* hrn_group_attention + the two hrn_group_weighted_sum calls of the correspondence heads in one pass over the rows
 * (reference layers.py:385-390, 447-450): a [groups*k] (nullable) = softmax_k(max_c E), af [groups, ldaf] = sum_j a_j
 * E[g*k+j,:], cor [groups,3] (nullable) = sum_j a_j xyz[b*N + idx[g*k+j],:] with b = g / groups_per_batch.
 * C % 4 == 0, 16-byte aligned rows, k*C*4 <= 96 KB -- HRN_ERR_UNSUPPORTED otherwise. */
int hrn_group_attend(const float* E, int ldE, int C, long long groups, int k, float* a, float* af, int ldaf,
                     const float* xyz, const int32_t* idx, int groups_per_batch, int N, float* cor, void* stream);

/* out[g,c] = max_j X[g*k+j, c]   (layers.py:202,208). */
int hrn_group_max(const float* X, int ldX, int C, long long groups, int k, float* out, int ldo, void* stream);

/* Geometry channels of a grouped tensor (layers.py:20-23; 284-288,364-365; 318-319; 438-445).
 *   q [B,M,3], p [B,N,3], idx [B,M,k] int32.  out [B*M*k, ldo]: cols 0..2 = p[idx]-q, 3 = |.|; if wq and wp are
 *   given also 4..6 = q, 7..9 = p[idx], 10 = wq[b,m], 11 = wp[b,idx].  nn (nullable) [B*M*k,3] = p[idx]. */
int hrn_group_geometry(const float* q, const float* p, const int32_t* idx, const float* wq, const float* wp, int B,
                       int M, int k, int N, float* out, int ldo, float* nn, void* stream);

/* w = (1/(sigma+1e-5)) / mean_m(1/(sigma+1e-5))   (models/HRegNet/models.py:30-32,36-38).  sigma, w [B,M]. */
int hrn_sigma_to_weights(const float* sigma, float* w, int B, int M, void* stream);

/* out[b,n,:] = R[b] x[b,n,:] + t[b]   (models.py:91-92,113-114).  x,out [B,N,3]; R [B,9]; t [B,3]. */
int hrn_transform_points(const float* x, const float* R, const float* t, float* out, int B, int N, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * 4. CoarseReg similarity features (layers.py:29-41,290-313,339-362)
 * ------------------------------------------------------------------------------------------------------- */

/* S [B,N1,C], D [B,N2,C] channels-last descriptors -> cosm [B,N2,N1] = <D,S>/(|D||S|+1e-6), rowmax [B,N2]
 * (max over n1), colmax [B,N1] (max over n2); nS [B,N1], nD [B,N2] scratch norms. */
int hrn_cosine_matrix(const float* S, const float* D, int B, int N1, int N2, int C, float* nS, float* nD, float* cosm,
                      float* rowmax, float* colmax, void* stream);

/* out[r, col_src_dst] = cosm[b,idx[r],i]/(colmax[b,i]+1e-6); out[r, col_dst_src] = cosm[b,idx[r],i]/(rowmax[b,idx[r]]
 * +1e-6), r = (b,i,j) -- the diagonal picks of layers.py:303-313,352-362 without the Python loops. */
int hrn_cosine_pick(const float* cosm, const float* rowmax, const float* colmax, const int32_t* idx, int B, int N1,
                    int N2, int k, float* out, int ldo, int col_src_dst, int col_dst_src, void* stream);

/* The same two feature columns as hrn_cosine_matrix + hrn_cosine_pick in ONE tcgen05 kernel per call (layers.py:29-41,
 * 292-313, 341-362): the [N1 x C].[C x N2] contraction on the tensor cores (bf16 hi/lo operands, all four partial
 * products, fp32 accumulation: ~1e-6 on the cosine), both families of maxima and the k picks taken from the accumulators --
 * the cosine matrix itself is never written.  N1 = 128 or 256, N2 <= 256 (multiple of 16), C multiple of 32, k = 8,
 * 16-byte aligned S / D; HRN_ERR_UNSUPPORTED otherwise. */
int hrn_cosine_features_tc(const float* S, const float* D, const int32_t* idx, int B, int N1, int N2, int C, int k,
                           float* out, int ldo, int col_src_dst, int col_dst_src, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * 5. Pose head (layers.py:456-504 WeightedSVDHead; models.py:100-110,120-127 pose composition)
 * ------------------------------------------------------------------------------------------------------- */

/* src, cor [B,N,3]; w [B,N] -> R [B,9] row-major, t [B,3].  With R_prev/t_prev also R_cmp = R R_prev,
 * t_cmp = R t_prev + t.  Degenerate covariance -> R = I, t = 0 (the reference's SVD-failure fallback).
 * pose12 (nullable) [B,12]: the final pose of the call (composed when R_prev is given) as packed rows [R | t] -- the
 * message of the multi-GPU pose gather. */
int hrn_weighted_kabsch(const float* src, const float* cor, const float* w, int B, int N, const float* R_prev,
                        const float* t_prev, float* R, float* t, float* R_cmp, float* t_cmp, float* pose12, void* stream);

/* RegressionHead (models/model_v2/layers.py:625-668; the pose head of model_v3): w' = w / (sum w + 1e-4); the two weighted
 * means [sum w' src | sum w' cor] go through two perceptrons 6 -> H1 -> H2 -> n_rot / 3 (ReLU between).  params = 12 device
 * pointers: {W1,b1,W2,b2,W3,b3} of the rotation branch, then of the translation branch (nn.Linear layout [out,in], fp32). */
int hrn_regression_head(const float* src, const float* cor, const float* w, int B, int N, const float* const* params,
                        int H1, int H2, int n_rot, float* rot_out, float* trans_out, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * 6. Steps around the path: input pipeline in front of it, evaluation metrics behind it
 * ------------------------------------------------------------------------------------------------------- */

/* dataset/dataset_utils.py:113-125 remove_points_by_range for a batch of sweeps: stable compaction of the points with
 * ||p|| < max_range (numpy float32 norm semantics, bit-identical kept set).  xyz [total,3], intensity [total]
 * (nullable), offsets [n_sweeps+1] int64 on the device; max_sweep_points >= the largest sweep (host value);
 * scratch [n_sweeps * ceil(max_sweep_points/4096)] int32.  Kept points of sweep s land in rows
 * [offsets[s], offsets[s]+count[s]) of the outputs; count [n_sweeps] int32.  Not in place. */
int hrn_range_filter(const float* xyz, const float* intensity, const long long* offsets, int n_sweeps,
                     long long max_sweep_points, float max_range, float* xyz_out, float* intensity_out, int* count,
                     int* scratch, void* stream);

/* dataset/dataset_utils.py:188-223 PointCloudResampler for a batch of filtered sweeps (hrn_range_filter layout):
 * out[b,i,:] = sweep_b[j], j = i for i < count[b] when count[b] <= n (the cloud, then its padding picks idx[b,i]),
 * j = idx[b,i] otherwise (the subset drawn without replacement).  idx [B,n] int32; out [B,n,3]. */
int hrn_resample_gather(const float* xyz, const long long* offsets, const int* count, const int* idx, int B, int n,
                        float* out, void* stream);

/* transform/rodrigues.py:526-550 SE3.exp: twist [B,6] = (w, v) -> g [B,16] row-major 4x4 (the perturbation applied by
 * transform/dataset_transforms.py:128-140; apply it with hrn_transform_points). */
int hrn_se3_exp(const float* twist, int B, float* g, void* stream);

/* Pose-error metrics fused after the pose head / the pose all-gather.  Replaces
 * losses/losses.py:138-164 (calc_rot_rre_err, calc_tran_rte_err), models/utils.py:132-138 (calc_error_np) [mode 0:
 * R_err = pred_R^T gt_R, t_err = pred_t - gt_t] and metrics/calibeval.py:72-106,172-196 (add_batch, geodesic_distance)
 * [mode 1: error = pred_tf . gt_tf].  pred_R, gt_R [B,9] row-major, pred_t, gt_t [B,3].  Outputs, each nullable:
 * geo [B] geodesic angle in degrees, eucl [B] = ||t_err||, euler [B,3] Euler XYZ angles of R_err in degrees, terr [B,3],
 * sums [8] = sum over the batch of (geo, eucl, |euler| xyz, |terr| xyz), ACCUMULATED (zero it before the first batch). */
int hrn_pose_errors(const float* pred_R, const float* pred_t, const float* gt_R, const float* gt_t, int B, int mode,
                    float* geo, float* eucl, float* euler, float* terr, float* sums, void* stream);

/* Host-side evaluation of the closed form used by hrn_weighted_kabsch (H row-major 3x3, fp64); for tests. */
int hrn_pose_from_covariance_host(const double* H9, const double* xbar, const double* ybar, double* R9, double* t3);

/* Library / build identification: returns a static string "hregnet_b200 <version> sm_100a". */
const char* hrn_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HREGNET_B200_H */
