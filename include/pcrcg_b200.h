/*
 * pcrcg_b200.h -- C ABI of libpcrcg_b200.so: the B200 (sm_100a) implementation of PCR-CG's
 * KPConv feature-extraction hot path.  Plain pointers and sizes only.
 *
 * Every function returns 0 on success, non-zero on failure; pcrcg_last_error() then returns a
 * message (thread local).  "_host" entry points take HOST buffers, copy in, compute on the current
 * CUDA device and copy out: they are what the reference's FFI for this path would bind.  "_dev"
 * entry points take DEVICE pointers plus a CUDA stream (cudaStream_t passed as void*) and a caller
 * provided workspace; they never synchronise and never allocate.
 *
 * Reference interfaces replaced (paths relative to the PCR-CG tree; "zip!" = cpp_wrappers.zip!cpp_wrappers/):
 *   pcrcg_subsample_batch_*  : zip!cpp_subsampling/wrapper.cpp:62-333  subsample_batch(points, batches, sampleDl, max_p)
 *                              -> zip!cpp_subsampling/grid_subsampling/grid_subsampling.cpp:109-211
 *   pcrcg_radius_*           : cpp_wrappers/cpp_neighbors/wrapper.cpp:58-238  batch_query(queries, supports, q_batches, s_batches, radius)
 *                              -> cpp_wrappers/cpp_neighbors/neighbors/neighbors.cpp:211-332
 *   (further entry points are declared below, next to the kernels they expose)
 */
#ifndef PCRCG_B200_H
#define PCRCG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* pcrcg_stream_t;   /* cudaStream_t */

const char* pcrcg_last_error(void);
int pcrcg_version(void);
void pcrcg_free(void* host_ptr);            /* frees buffers returned by *_host entry points */

/* Measurement hooks: per kernel-class device time from CUDA events recorded on the launching stream
 * around each class's launches (enable, run, report: ms[c] / counts[c] for c < pcrcg_profile_classes()),
 * and the running count of kernels this library launched. */
void pcrcg_profile_enable(int32_t on);
int32_t pcrcg_profile_classes(void);
const char* pcrcg_profile_class_name(int32_t c);
int pcrcg_profile_report(double* ms, int64_t* counts);
uint64_t pcrcg_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Grid subsampling.  points [n,3] fp32 (stacked clouds), lens [nb] int32.  Output order and
 * barycentre bits are those of the reference (libstdc++ unordered_map iteration order, fp32
 * sequential sums).  max_p <= 0 : keep all.
 * ------------------------------------------------------------------------------------------- */
size_t pcrcg_subsample_ws_bytes(int64_t n, int32_t nb);
/* out_points must hold n*3 floats (upper bound), out_lens nb int32 (device). */
int pcrcg_subsample_batch_dev(const float* points, int64_t n, const int32_t* lens, int32_t nb, float sampleDl,
                              int32_t max_p, float* out_points, int32_t* out_lens, void* ws, size_t ws_bytes,
                              pcrcg_stream_t stream);
/* Host buffers in, freshly malloc'ed *out_points ([*out_m,3]) out; out_lens [nb] is caller provided. */
int pcrcg_subsample_batch_host(const float* points, int64_t n, const int32_t* lens, int32_t nb, float sampleDl,
                               int32_t max_p, float** out_points, int64_t* out_m, int32_t* out_lens);

/* The same with per-point features [n,fdim] fp32 and / or integer classes [n,ldim] (either may be NULL; its dim is then
 * ignored): zip!cpp_subsampling/wrapper.cpp:62-333 subsample_batch(points, batches, features=, classes=, ...) ->
 * zip!cpp_subsampling/grid_subsampling/grid_subsampling.cpp:34-102.  Feature rows are the fp32 sums in point order divided by
 * (float)count; a class is the first maximal vote in the iteration order of libstdc++'s unordered_map<int,int>
 * (csrc/label_vote.h).  Not on the KPConv pyramid's path (datasets/dataloader.py:289 passes neither).
 * ldim > 1 needs nb == 1 (grid_subsampling.cpp:157-158 mis-slices the classes of later clouds: no reference behaviour).
 * out_features [n,fdim], out_classes [n,ldim]: upper bounds like out_points.  status: device int32, required with classes,
 * set to 1 if a voxel held more than 64 distinct labels in one column (its vote is then unreliable); read it after
 * synchronising.  The _host form checks it and fails. */
size_t pcrcg_subsample_ex_ws_bytes(int64_t n, int32_t nb, int32_t fdim, int32_t ldim);
int pcrcg_subsample_batch_ex_dev(const float* points, int64_t n, const int32_t* lens, int32_t nb, float sampleDl, int32_t max_p,
                                 const float* features, int32_t fdim, const int32_t* classes, int32_t ldim, float* out_points,
                                 int32_t* out_lens, float* out_features, int32_t* out_classes, int32_t* status, void* ws,
                                 size_t ws_bytes, pcrcg_stream_t stream);
/* *out_features ([*out_m,fdim]) / *out_classes ([*out_m,ldim]) are malloc'ed like *out_points (pcrcg_free). */
int pcrcg_subsample_batch_ex_host(const float* points, int64_t n, const int32_t* lens, int32_t nb, float sampleDl, int32_t max_p,
                                  const float* features, int32_t fdim, const int32_t* classes, int32_t ldim, float** out_points,
                                  int64_t* out_m, int32_t* out_lens, float** out_features, int32_t** out_classes);

/* ---------------------------------------------------------------------------------------------
 * Radius search.  queries [nq,3], supports [ns,3] fp32; q_lens / s_lens [nb] int32.
 * Rows: neighbours in ascending (d2, index), global support indices, padded with ns.
 * ------------------------------------------------------------------------------------------- */
/* Row starts of groups of `group` consecutive clouds (group = 2: fragment pairs = InstanceNorm segments):
 * out[k] = sum of lens[0 .. k*group), k in [0, nb/group]; *total (device, may be NULL) = sum of all lens. */
int pcrcg_group_starts_dev(const int32_t* lens, int32_t nb, int32_t group, int32_t* out, int32_t* total, pcrcg_stream_t stream);

size_t pcrcg_radius_ws_bytes(int64_t nq, int64_t ns, int32_t nb);
/* Step 1: bin the supports (grid state is kept inside ws). */
int pcrcg_radius_build_dev(const float* supports, int64_t ns, const int32_t* s_lens, int32_t nb, float radius,
                           void* ws, size_t ws_bytes, pcrcg_stream_t stream);
/* Step 2: query.  rows may be NULL (count only).  rows is [nq,row_stride] int32 of which the first
 * `width` columns are written; counts [nq] (may be NULL) gets the un-truncated neighbour count;
 * max_count (device int32, may be NULL) its maximum. */
int pcrcg_radius_query_dev(const float* queries, int64_t nq, const int32_t* q_lens, int64_t ns, int32_t nb, float radius,
                           int32_t width, int32_t row_stride, int32_t* rows, int32_t* counts, int32_t* max_count,
                           void* ws, size_t ws_bytes, pcrcg_stream_t stream);
/* Step 2, cell-centric (the default of the Python layer): one warp per occupied query cell; the 27 adjacent support cells
 * are resolved and their records staged in shared memory once per cell, every query of the cell tests the staged candidates.
 * queries_are_supports != 0: `queries` is the point set the grid was built from (conv lists; qws unused).  Otherwise the
 * queries are first binned into the support grid's cells inside qws (pcrcg_radius_query_ws_bytes(nq, nb) bytes).  Same
 * results as pcrcg_radius_query_dev, bit for bit. */
size_t pcrcg_radius_query_ws_bytes(int64_t nq, int32_t nb);
int pcrcg_radius_query_cells_dev(const float* queries, int64_t nq, const int32_t* q_lens, int64_t ns, int32_t nb, float radius,
                                 int32_t width, int32_t row_stride, int32_t* rows, int32_t* counts, int32_t* max_count,
                                 void* ws, size_t ws_bytes, void* qws, size_t qws_bytes, int32_t queries_are_supports,
                                 pcrcg_stream_t stream);
/* Host buffers in; *out_rows is malloc'ed [nq,*out_width] with *out_width = max_count when limit<=0
 * (the reference's output) or min(limit, max_count). */
int pcrcg_batch_query_host(const float* queries, int64_t nq, const float* supports, int64_t ns, const int32_t* q_lens,
                           const int32_t* s_lens, int32_t nb, float radius, int32_t limit, int32_t** out_rows,
                           int32_t* out_width);

/* ---------------------------------------------------------------------------------------------
 * KPConv forward -- models/blocks.py:229-374 KPConv.forward(q_pts, s_pts, neighb_inds, x), rigid
 * kernel, KP_influence 'linear', aggregation 'sum'.  neighb_inds [nq,H] (row stride idx_stride) is
 * int32 or int64 (idx_is_i64), shadow index = ns.  kernel_points [K,3], weights [K,cin,cout]
 * (the reference's Parameter layouts), out [nq,cout].  All device pointers.
 * ------------------------------------------------------------------------------------------- */
size_t pcrcg_kpconv_ws_bytes(int64_t nq, int64_t ns, int32_t cin, int32_t K);
int pcrcg_kpconv_forward_dev(const float* q_pts, int64_t nq, const float* s_pts, int64_t ns, const void* neighb_inds,
                             int32_t idx_is_i64, int32_t H, int32_t idx_stride, const float* x, int32_t cin,
                             const float* kernel_points, int32_t K, float KP_extent, const float* weights, int32_t cout,
                             float* out, void* ws, size_t ws_bytes, pcrcg_stream_t stream);

/* Same, with the features ALSO available as bf16 (hi, lo) planes [ns, ldxs] (x = hi + lo; emitted by
 * pcrcg_norm_act_dev): the aggregation then runs its bf16x3 ldmatrix / mma.m16n8k16 kernel.  x may then be NULL if
 * row_positive is given, cin % 64 == 0 and cout % 16 == 0 (features that exist as planes only). */
int pcrcg_kpconv_forward_split_dev(const float* q_pts, int64_t nq, const float* s_pts, int64_t ns, const void* neighb_inds,
                                   int32_t idx_is_i64, int32_t H, int32_t idx_stride, const float* x, const void* x_hi,
                                   const void* x_lo, int32_t ldxs, const uint8_t* row_positive /* may be NULL */, int32_t cin,
                                   const float* kernel_points, int32_t K,
                                   float KP_extent, const float* weights, int32_t cout, float* out, void* ws, size_t ws_bytes,
                                   pcrcg_stream_t stream);

/* Same, and the contraction epilogue ALSO accumulates the InstanceNorm statistics of `out` (the BatchNormBlock that
 * follows every KPConv, models/blocks.py:590,662): stats_acc [nseg][2][cout] fp64, zeroed by the caller, receives per
 * (segment, column) the sum and the sum of squares; pcrcg_colstats_final_dev turns them into mean / rstd.  Saves the
 * statistics pass over `out`.  Tensor-core path only (cout % 16 == 0); x_hi / x_lo may be NULL. */
int pcrcg_kpconv_forward_stats_dev(const float* q_pts, int64_t nq, const float* s_pts, int64_t ns, const void* neighb_inds,
                                   int32_t idx_is_i64, int32_t H, int32_t idx_stride, const float* x, const void* x_hi,
                                   const void* x_lo, int32_t ldxs, const uint8_t* row_positive, int32_t cin,
                                   const float* kernel_points, int32_t K, float KP_extent, const float* weights, int32_t cout,
                                   float* out, void* ws, size_t ws_bytes, const int32_t* seg_starts, int32_t nseg,
                                   double* stats_acc, pcrcg_stream_t stream);

/* Dense contraction C[M,N] = A[M,K] * B (* row_scale[m] if not NULL).  B is [K,N] (b_is_nk = 0) or
 * [N,K] (b_is_nk = 1, nn.Linear.weight of models/blocks.py:490).  Row-major, leading dims in elements. */
int pcrcg_gemm_dev(const float* A, int32_t lda, const float* B, int32_t ldb, int32_t b_is_nk, float* C, int32_t ldc,
                   int32_t M, int32_t N, int32_t K, const float* row_scale, pcrcg_stream_t stream);
/* The tensor-core contraction with operands already split into bf16 (hi, lo) planes
 * (x = hi + lo, hi = bf16(x), lo = bf16(x - hi)):  a_* [M,ldk], b_* [N,ldk], K contiguous, ldk % 8 == 0,
 * N % 16 == 0.  pcrcg_split_bf16_dev produces such planes from fp32 rows. */
int pcrcg_split_bf16_dev(const float* x, int32_t ldx, int64_t rows, int32_t cols, void* hi, void* lo, int32_t ldo,
                         pcrcg_stream_t stream);
int pcrcg_gemm_bf16x3_dev(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, int32_t ldk, float* C,
                          int32_t ldc, int32_t M, int32_t N, int32_t K, const float* row_scale, pcrcg_stream_t stream);
/* Same with the column statistics of C accumulated by the epilogue (see pcrcg_kpconv_forward_stats_dev; stats_acc is
 * cleared by the call, as in every *_stats_dev entry point):
 * UnaryBlock = Linear -> InstanceNorm (models/blocks.py:497-499) without a statistics pass over C. */
int pcrcg_gemm_bf16x3_stats_dev(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, int32_t ldk, float* C,
                                int32_t ldc, int32_t M, int32_t N, int32_t K, const float* row_scale, const int32_t* seg_starts,
                                int32_t nseg, double* stats_acc, pcrcg_stream_t stream);
/* 1: force the fp32 CUDA-core contraction (parity anchor); 0: tcgen05 tensor-core path where shapes allow. */
void pcrcg_gemm_force_simt(int32_t on);
/* A/B switches for measurements: "contraction_simt", "aggregate_simt" (CUDA-core variants of the two KPConv stages),
 * "aggregate_pipelined" (persistent software-pipelined bf16 aggregation, default 1; 0 = one point per warp),
 * "kpconv_fused" (0 = two-kernel KPConv everywhere, 1 = one-kernel KPConv for cin == cout == 64 (default), 2 = for every
 * cin, cout multiple of 64),
 * "kpconv_chunk_mb" (bound of one KPConv intermediate buffer in MiB, 0 = default 8192; query rows beyond it are processed in chunks),
 * "first_layer_fused" (cin <= 4: aggregation + contraction in one kernel; default 0, the two-kernel path measures faster),
 * "norm_vectorised" (float4 InstanceNorm apply kernel, default 1), "norm_variant" (0 = per-mode launch shape, 1..4 fixed),
 * "stats_debug" (contraction-epilogue statistics: 1 = no atomics, 2 = no column sums; timing ablation only). */
int pcrcg_set_option(const char* name, int32_t value);

/* ---------------------------------------------------------------------------------------------
 * InstanceNorm over the rows of each segment (= fragment pair) -- models/blocks.py:448,456-463 --
 * fused with LeakyReLU (:501) and the residual sum of ResnetBottleneckBlock.forward (:678).
 * seg_starts [nseg+1] int32 device.  mean / rstd [nseg,C].
 *   out = act( (x-mean)*rstd + shortcut' ),  shortcut' = (sc-sc_mean)*sc_rstd | sc | 0,
 *   act = LeakyReLU(slope) when slope >= 0, identity when slope < 0;  mean == NULL skips the normalisation.
 * split_hi / split_lo (bf16 [n, split_ld], may be NULL): the result is ALSO emitted as the (hi, lo) planes the
 * next tensor-core contraction consumes, saving that contraction's own split pass.
 * out may be NULL when the planes are requested (a result consumed only by tensor-core contractions: saves the fp32 write).
 * row_positive (uint8 [n], may be NULL; needs a power-of-two C): flag[r] = (sum_c out[r,c] > 0), the per-row
 * predicate of the KPConv neighbour count (models/blocks.py:369-370), saving the consumer a pass over out.
 * ------------------------------------------------------------------------------------------- */
int pcrcg_colstats_dev(const float* x, int64_t n, int32_t C, const int32_t* seg_starts, int32_t nseg, float eps,
                       float* mean, float* rstd, pcrcg_stream_t stream);
/* stats_acc [nseg][2][C] fp64 (sum, sum of squares; from the *_stats_dev contractions) -> mean / rstd [nseg,C]
 * (biased variance, models/blocks.py:448) */
int pcrcg_colstats_final_dev(const double* stats_acc, const int32_t* seg_starts, int32_t nseg, int32_t C, float eps,
                             float* mean, float* rstd, pcrcg_stream_t stream);
int pcrcg_norm_act_dev(const float* x, int64_t n, int32_t C, const int32_t* seg_starts, int32_t nseg, const float* mean,
                       const float* rstd, const float* sc, const float* sc_mean, const float* sc_rstd, float slope,
                       float* out, void* split_hi, void* split_lo, int32_t split_ld, uint8_t* row_positive,
                       pcrcg_stream_t stream);

/* The same with the SHORTCUT given as bf16 (hi, lo) planes [n, sc_ld] (sc = hi + lo): a block output that was emitted as
 * planes only (its fp32 copy never written) feeding the next block's residual sum.  Needs C % 4 == 0. */
int pcrcg_norm_act_planes_dev(const float* x, int64_t n, int32_t C, const int32_t* seg_starts, int32_t nseg, const float* mean,
                              const float* rstd, const void* sc_hi, const void* sc_lo, int32_t sc_ld, const float* sc_mean,
                              const float* sc_rstd, float slope, float* out, void* split_hi, void* split_lo, int32_t split_ld,
                              uint8_t* row_positive, pcrcg_stream_t stream);

/* Descriptor head (models/architectures.py:572-582, "next" row of the scope table): x [n, F+2] ->
 * feats [n,F] = x[:, :F] / max(|x[:, :F]|_2, 1e-12); overlap / saliency [n] = clamp(sigmoid(x[:, F | F+1]), 0, 1), NaN/Inf -> 0 */
int pcrcg_descriptor_head_dev(const float* x, int64_t n, int32_t F, float* feats, float* overlap, float* saliency,
                              pcrcg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Bottleneck overlap-attention GNN ("next" row 3 of the scope table) -- models/gcn.py, models/architectures.py:528-565.
 * Dense contractions of these layers use pcrcg_gemm_dev; these are the remaining operators.  cloud_starts [nb+1] int32
 * device = row starts of the stacked clouds.
 *   pcrcg_knn_dev            get_graph_feature's kNN (models/gcn.py:48-51): out [n,k] int32 global rows = the k+1 smallest
 *                            of clamp(-2 x.y + |x|^2 + |y|^2, 1e-12) inside the query's cloud, first dropped; ties by index
 *   pcrcg_edge_max_stats_dev 1x1 conv over the edge features [f_n ; f_j - f_n] (models/gcn.py:54-66,125-131) after the split
 *                            W [f_n ; f_j - f_n] = u_n + v_j:  out [n,C] = u_n + max_j v[knn[n,j]]  and stats_acc
 *                            [nb][2][C] fp64 += (sum, sum of squares)/k of u_n + v_j over all edges (InstanceNorm2d
 *                            statistics; finish with pcrcg_colstats_final_dev, apply with pcrcg_norm_act_dev)
 *   pcrcg_bias_act_dev       out = act(x + bias[c]) (Conv1d bias; slope < 0: identity, 0: ReLU, > 0: LeakyReLU); in place ok
 *   pcrcg_softmax_rows_dev   x[r, 0:m] <- softmax(scale * x[r, 0:m])   (attention, models/gcn.py:153-157; saliency :561-563)
 *   pcrcg_l2norm_rows_dev    out = x / max(|x|_2, eps) per row          (F.normalize, models/architectures.py:543)
 * ------------------------------------------------------------------------------------------- */
int pcrcg_knn_dev(const float* points, int64_t n, const int32_t* cloud_starts, int32_t nb, int32_t k, int32_t* out,
                  pcrcg_stream_t stream);
int pcrcg_edge_max_stats_dev(const float* u, int32_t ldu, const float* v, int32_t ldv, const int32_t* knn, int64_t n,
                             int32_t C, int32_t k, const int32_t* cloud_starts, int32_t nb, float* out, double* stats_acc,
                             pcrcg_stream_t stream);
int pcrcg_bias_act_dev(const float* x, int64_t n, int32_t C, const float* bias, float slope, float* out,
                       pcrcg_stream_t stream);
/* point2node / node visibility of the collate ("next" row 1; datasets/dataloader.py:91-106, 133-158, 309-322):
 *   out[i] = nearest node of point i inside its cloud (index LOCAL to the cloud's nodes; expanded squared distance of
 *   datasets/dataloader.py:70-90, ties by node index);  total / visible_count [n_nodes] int32 (zeroed by the caller) +=
 *   number of points assigned to each node / of those with visible[i] != 0. */
int pcrcg_point2node_dev(const float* points, int64_t n, const int32_t* point_starts, const float* nodes,
                         const int32_t* node_starts, int32_t nb, int32_t* out, pcrcg_stream_t stream);
int pcrcg_node_counts_dev(const int32_t* point2node, const uint8_t* visible, int64_t n, const int32_t* point_starts,
                          const int32_t* node_starts, int32_t nb, int32_t* total, int32_t* visible_count,
                          pcrcg_stream_t stream);
int pcrcg_softmax_rows_dev(float* x, int64_t n, int32_t m, int32_t ld, float scale, pcrcg_stream_t stream);
int pcrcg_l2norm_rows_dev(const float* x, int64_t n, int32_t C, float eps, float* out, pcrcg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Descriptor matching front-end ("next" row 4) -- lib/benchmark_utils.py:187-224,246-262,270-295.
 *   pcrcg_best_match_dev   best_idx[i] = argmax_j <a[i,:], b[j,:]> (first maximum, as np.argmax / torch.max), best_val[i]
 *                          (may be NULL) its score; a [n,D], b [m,D] row-major, D in {16,32,64}.  The [n,m] score matrix
 *                          the reference builds is never materialised.
 *   pcrcg_mutual_dev       mutual[i] = (col_best[row_best[i]] == i): mutual_selection's {0,1} matrix in sparse form
 * ------------------------------------------------------------------------------------------- */
int pcrcg_best_match_dev(const float* a, int64_t n, const float* b, int64_t m, int32_t D, int32_t* best_idx, float* best_val,
                         pcrcg_stream_t stream);
int pcrcg_mutual_dev(const int32_t* row_best, const int32_t* col_best, int64_t n, uint8_t* mutual, pcrcg_stream_t stream);

/* max_pool (models/blocks.py:86-102) and closest_pool (:71-83): x [ns,C], inds [nq,H] -> out [nq,C] */
int pcrcg_max_pool_dev(const float* x, int64_t ns, int32_t C, const void* inds, int32_t idx_is_i64, int64_t nq, int32_t H,
                       int32_t idx_stride, float* out, pcrcg_stream_t stream);
/* max_pool of features held as bf16 (hi, lo) planes [ns, ldx] -> planes [nq, ldo]: an entry's value is hi + lo, the
 * winner's pair is copied (the output represents the maximum exactly); a shadow index contributes 0.  C % 4 == 0. */
int pcrcg_max_pool_planes_dev(const void* x_hi, const void* x_lo, int64_t ns, int32_t C, int32_t ldx, const void* inds,
                              int32_t idx_is_i64, int64_t nq, int32_t H, int32_t idx_stride, void* out_hi, void* out_lo, int32_t ldo,
                              pcrcg_stream_t stream);
int pcrcg_closest_pool_dev(const float* x, int64_t ns, int32_t C, const void* inds, int32_t idx_is_i64, int64_t nq,
                           int32_t idx_stride, float* out, pcrcg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Colour path.  pcrcg_projection_dev = projection.py:31-61 Projection.projection(points, depth_map,
 * world2camera): index lists inds2d [M,2] (x,y) / inds3d [M] (ascending), int64, *count = M (device).
 * world2camera / intrinsics: HOST pointers to 16 floats (row-major 4x4).  Outputs sized for n.
 * pcrcg_project_scatter_dev = projection + the gather/scatter of models/architectures.py:273-307,
 * 360-370 fused: out [n, C+1]; views are given in the reference's WRITE order (the last view that
 * sees a point wins); view v applies to points [row_lo[v], row_hi[v]); unseen points get base[i]
 * (1 when base is NULL) in every column.  depth / feat ([C,H,W]) / valid ([H,W], entries may be
 * NULL) are HOST arrays of DEVICE pointers; w2c / k4 HOST arrays of nviews x 16 floats.
 * ------------------------------------------------------------------------------------------- */
size_t pcrcg_projection_ws_bytes(int64_t n);
int pcrcg_projection_dev(const float* points, int64_t n, const float* depth, int32_t H, int32_t W, const float* world2camera,
                         const float* intrinsics, float thresh, int64_t* inds2d, int64_t* inds3d, int32_t* count, void* ws,
                         size_t ws_bytes, pcrcg_stream_t stream);
int pcrcg_project_scatter_dev(const float* points, int64_t n, int32_t nviews, const float* const* depth, const float* const* feat,
                              const float* const* valid, const float* w2c, const float* k4, const int32_t* row_lo,
                              const int32_t* row_hi, int32_t H, int32_t W, int32_t C, float thresh, const float* base, float* out,
                              pcrcg_stream_t stream);
/* The same for a stacked batch (any number of clouds and views): everything in DEVICE memory.  cloud_starts [nb+1] = row
 * starts of the clouds; the views of cloud c are views[view_starts[c] .. view_starts[c+1]) in write order; one view is
 * { const float* depth; const float* feat; const float* valid; float w2c[12]; float k4[12]; } (rows 0-2 of the 4x4
 * matrices, row-major; 120 bytes, 8-byte aligned). */
int pcrcg_project_scatter_batch_dev(const float* points, int64_t n, const int32_t* cloud_starts, int32_t nb,
                                    const int32_t* view_starts, const void* views, int32_t H, int32_t W, int32_t C, float thresh,
                                    const float* base, float* out, pcrcg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
