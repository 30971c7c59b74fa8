/*
 * de6d_b200.h -- C ABI of libde6d_b200.so: B200-native (sm_100a) point-set-abstraction and box ops for
 * Det6D / SASA / 3DSSD style detectors.
 *
 * This is the drop-in boundary.  Every entry point replaces one function of the reference's three pybind11
 * extension modules (paths relative to core/pcdet/ops/ of HITSZ-NRSL/De6D):
 *   pointnet2_batch_cuda   pointnet2/pointnet2_batch/src/pointnet2_api.cpp:11-30
 *   iou3d_nms_cuda         iou3d_nms/src/iou3d_nms_api.cpp:11-17
 *   roiaware_pool3d_cuda   roiaware_pool3d/src/roiaware_pool3d.cpp:172-177
 * with the same argument order and meaning, minus the at::Tensor wrappers: plain device pointers, sizes,
 * scalars and an explicit cudaStream_t (the reference launches on the legacy default stream).
 *
 * Conventions
 *   - all pointers are DEVICE pointers to contiguous arrays on the current device, unless marked HOST;
 *   - the caller owns and pre-initialises every output exactly like the reference Python wrappers do
 *     (temp = 1e10, idx = 0, idx_cnt = 0, box index = -1, grads = 0); nothing is allocated inside;
 *   - every function returns DE6D_OK (0) or an error code; de6d_last_error_string() describes the last
 *     failure of the calling thread.  (The reference prints to stderr and exit(-1)s instead.)
 *   - functions are asynchronous with respect to the host and re-entrant; no global state except the
 *     per-kernel "max dynamic shared memory" attribute set on first use;
 *   - results: indices / counts / keep lists are bit-identical to the reference kernels; floating-point
 *     outputs are identical for the copy ops and three_interpolate, within 1e-5 relative for IoUs.
 */
#ifndef DE6D_B200_H
#define DE6D_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_API_H__
typedef struct CUstream_st *cudaStream_t;
#endif

#define DE6D_OK 0
#define DE6D_ERR_INVALID 1
#define DE6D_ERR_CUDA 2

const char *de6d_last_error_string(void);
int de6d_version(void);
const char *de6d_build_info(void);
/* kernels launched through this library since it was loaded (diagnostic; bench.py reports it) */
long long de6d_launch_count(void);

/* ---- pointnet2_batch: sampling ------------------------------------------------------------------------- */

/* farthest_point_sampling_wrapper(b, n, m, xyz, temp, idx)      sampling.cpp:41-50, sampling_gpu.cu:101-266
 * xyz (b,n,3) f32; temp (b,n) f32 in/out running min squared distances (caller fills 1e10); idx (b,m) i32. */
int de6d_furthest_point_sampling(int b, int n, int m, const float *xyz, float *temp, int *idx, cudaStream_t stream);

/* furthest_point_sampling_matrix_wrapper(b, n, m, matrix, temp, idx)   sampling.cpp:52-61, sampling_gpu.cu:268-417
 * matrix (b,n,n) f32 pairwise distances. */
int de6d_furthest_point_sampling_matrix(int b, int n, int m, const float *matrix, float *temp, int *idx,
                                        cudaStream_t stream);

/* calc_dist_matrix_for_sampling(xyz, features, gamma)   pointnet2_utils.py:36-44 (two torch.cdist + scale + add in the
 * reference): out[b,i,j] = |xyz_i - xyz_j| + gamma * |features_i - features_j|, (b,n,n) f32, one kernel, one write.
 * xyz (b,n,3); features: element (b,i,ch) at features[b*stride_b + i*stride_n + ch*stride_c] (strides in elements,
 * so both a (b,n,c) tensor and the permuted view of a (b,c,n) tensor are accepted), or NULL / c == 0 for none. */
int de6d_dist_matrix(int b, int n, int c, const float *xyz, const float *features, long long stride_b,
                     long long stride_n, long long stride_c, float gamma, float *out, cudaStream_t stream);

/* Fused F-FPS: the indices de6d_dist_matrix + de6d_furthest_point_sampling_matrix would give (bit for bit), without the
 * (b,n,n) matrix: one thread-block cluster (6 or 8 CTAs) per cloud keeps the features in distributed shared memory and
 * evaluates only the m selected rows (pointnet2_modules.py:383-388 is the call pair this replaces).  features as in
 * de6d_dist_matrix (element strides).  ..._fits(n, c) == 0: shape does not fit on chip, use the two-call form. */
int de6d_furthest_point_sampling_features_fits(int n, int c);
int de6d_furthest_point_sampling_features(int b, int n, int c, int m, const float *xyz, const float *features,
                                          long long stride_b, long long stride_n, long long stride_c, float gamma,
                                          float *temp, int *idx, cudaStream_t stream);
/* The same with the thread-block cluster size pinned (0 = automatic, 4, 6 or 8) and the kernel form pinned (prune: 0 =
 * automatic (dense wherever it fits: measured faster), 1 = dense: every distance of every selected row, 2 = pruned: a warp skips its 64-point bucket when the
 * bucket's bounding box proves that no running min-distance can change; 3 = the same with the surviving buckets evaluated by the
 * whole CTA; an error if the shape is not covered).  cluster_size 4 (64 channels, 3073..4096 points) applies to the dense form.
 * Identical results in every combination -- tests and tuning. */
int de6d_furthest_point_sampling_features_impl(int b, int n, int c, int m, const float *xyz, const float *features,
                                               long long stride_b, long long stride_n, long long stride_c, float gamma,
                                               float *temp, int *idx, int cluster_size, int prune, cudaStream_t stream);

/* furthest_point_sampling_weights_wrapper(b, n, m, xyz, weights, temp, idx)   sampling.cpp:63-73, sampling_gpu.cu:419-585
 * weights (b,n) f32; first index = argmax(weights). */
int de6d_furthest_point_sampling_weights(int b, int n, int m, const float *xyz, const float *weights, float *temp,
                                         int *idx, cudaStream_t stream);

/* Same results, explicit kernel choice (testing / benchmarking): impl 0 = default, 1 = on-chip kernel with bucket
 * pruning disabled, 2 = generic global-memory kernel, 3 = twice the warps, 4 = pruned, one sample per barrier round,
 * 5 = pruned, up to four samples per round (exact speculation; the default above 4096 points), 6 = register-resident
 * kernel without pruning for clouds of 32..4096 points (the default there; other sizes fall through to the default). */
int de6d_furthest_point_sampling_impl(int b, int n, int m, const float *xyz, float *temp, int *idx, int impl,
                                      cudaStream_t stream);
int de6d_furthest_point_sampling_weights_impl(int b, int n, int m, const float *xyz, const float *weights,
                                              float *temp, int *idx, int impl, cudaStream_t stream);

/* gather_points_wrapper(b, c, n, npoints, points, idx, out)   sampling.cpp:18-26, sampling_gpu.cu:16-52
 * points (b,c,n), idx (b,npoints) -> out (b,c,npoints). */
int de6d_gather_points(int b, int c, int n, int npoints, const float *points, const int *idx, float *out,
                       cudaStream_t stream);
/* gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points)   sampling.cpp:29-38, sampling_gpu.cu:54-91
 * accumulates into grad_points (b,c,n) (caller zeroes). */
int de6d_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out, const int *idx,
                            float *grad_points, cudaStream_t stream);

/* ---- pointnet2_batch: ball query ---------------------------------------------------------------------- */

/* ball_query_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx)   ball_query.cpp:32-42, ball_query_gpu.cu:15-51
 * NOTE argument order: new_xyz (b,m,3) before xyz (b,n,3).  idx (b,m,nsample) i32, caller zeroes. */
int de6d_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz, int *idx,
                    cudaStream_t stream);
/* ball_query_cnt_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx_cnt, idx)   ball_query.cpp:44-55, ball_query_gpu.cu:93-130 */
int de6d_ball_query_cnt(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz,
                        int *idx_cnt, int *idx, cudaStream_t stream);
/* ball_query_dilated_wrapper(b, n, m, radius_in, radius_out, nsample, new_xyz, xyz, idx_cnt, idx)
 * ball_query.cpp:57-67, ball_query_gpu.cu:53-91 */
int de6d_ball_query_dilated(int b, int n, int m, float radius_in, float radius_out, int nsample, const float *new_xyz,
                            const float *xyz, int *idx_cnt, int *idx, cudaStream_t stream);

/* The three calls above pick a kernel by size: clouds of >= 2048 points go through a per-cloud uniform grid (same
 * results, see csrc/ball_query.cu) and need device scratch, which they take from cudaMallocAsync on `stream`.
 * Hosts with their own allocator (the torch layer) use the explicit form instead:
 *   mode 0 = ball_query, 1 = ball_query_cnt, 2 = ball_query_dilated;  impl 0 = automatic, 1 = brute-force kernel,
 *   2 = grid kernel;  workspace = de6d_ball_query_workspace_bytes(b, n) bytes (0 = none needed) or NULL;
 *   impl 3 = grid kernel over a grid that de6d_ball_query_grid_build already left in `workspace`. */
size_t de6d_ball_query_workspace_bytes(int b, int n);
int de6d_ball_query_ex(int mode, int impl, int b, int n, int m, float radius_in, float radius_out, int nsample,
                       const float *new_xyz, const float *xyz, int *idx_cnt, int *idx, void *workspace,
                       size_t workspace_bytes, cudaStream_t stream);
/* One search grid per cloud shared by several queries of the same `xyz`: the radius scales of an SA layer
 * (pointnet2_modules.py:462-463 runs one QueryWithCntAndGroup per scale over the same cloud; the reference rescans all
 * n points per scale).  radius = the smallest radius that will be queried; workspace = de6d_ball_query_grid_bytes(b, n)
 * bytes.  Then call de6d_ball_query_ex(mode, 3, ...) once per scale with the same workspace. */
size_t de6d_ball_query_grid_bytes(int b, int n);
int de6d_ball_query_grid_build(int b, int n, float radius, const float *xyz, void *workspace, size_t workspace_bytes,
                               cudaStream_t stream);

/* ---- pointnet2_batch: grouping ------------------------------------------------------------------------ */

/* group_points_wrapper(b, c, n, npoints, nsample, points, idx, out)   group_points.cpp:30-39, group_points_gpu.cu:53-92
 * points (b,c,n), idx (b,npoints,nsample) -> out (b,c,npoints,nsample). */
int de6d_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx, float *out,
                      cudaStream_t stream);
/* impl 0 = auto, 1 = direct-gather kernel, 2 = TMA-staged shared-memory kernel (identical results). */
int de6d_group_points_impl(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx,
                           float *out, int impl, cudaStream_t stream);
/* Fused tail of QueryAndGroup / QueryWithCntAndGroup / QueryAndGroupDilated with use_xyz (pointnet2_utils.py:368-387,
 * 410-424, 449-463): out (b, 3+c, npoints, nsample) = cat(xyz[idx] - new_xyz, features[idx]) in one pass.
 * xyz (b,n,3), new_xyz (b,npoints,3), features (b,c,n) (NULL when c == 0), idx (b,npoints,nsample). */
int de6d_group_concat(int b, int c, int n, int npoints, int nsample, const float *xyz, const float *new_xyz,
                      const float *features, const int *idx, float *out, cudaStream_t stream);
/* The same with an optional transposed copy of the cloud, xyz_t (b,3,n) (`xyz_flipped` of pointnet2_modules.py:374):
 * coordinate rows are then staged by TMA exactly like feature rows.  NULL = de6d_group_concat. */
int de6d_group_concat_t(int b, int c, int n, int npoints, int nsample, const float *xyz, const float *xyz_t,
                        const float *new_xyz, const float *features, const int *idx, float *out, cudaStream_t stream);
/* new_xyz = xyz[sample_idx] in (b,m,3) and/or transposed (b,3,m) layout, one launch: replaces the SA module's
 * transpose(1,2).contiguous() -> gather_operation -> transpose(1,2).contiguous() (pointnet2_modules.py:374,451-454).
 * idx (b,m) or NULL (identity, m == n: plain transposition); either output may be NULL. */
int de6d_gather_xyz(int b, int n, int m, const float *xyz, const int *idx, float *new_xyz, float *new_xyz_t,
                    cudaStream_t stream);
/* group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points)   group_points.cpp:18-27, group_points_gpu.cu:14-51 */
int de6d_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out, const int *idx,
                           float *grad_points, cudaStream_t stream);

/* ---- pointnet2_batch: interpolation ------------------------------------------------------------------- */

/* three_nn_wrapper(b, n, m, unknown, known, dist2, idx)   interpolate.cpp:21-30, interpolate_gpu.cu:16-81
 * unknown (b,n,3), known (b,m,3) -> dist2 (b,n,3) SQUARED distances, idx (b,n,3). */
int de6d_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx,
                  cudaStream_t stream);
/* Explicit form: impl 0 = automatic (grid search over `known` for m >= 2048, identical results), 1 = brute-force scan,
 * 2 = grid; workspace = de6d_ball_query_workspace_bytes(b, m) bytes of device scratch, or NULL (cudaMallocAsync). */
int de6d_three_nn_ex(int impl, int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx,
                     void *workspace, size_t workspace_bytes, cudaStream_t stream);
/* three_interpolate_wrapper(b, c, m, n, points, idx, weight, out)   interpolate.cpp:33-45, interpolate_gpu.cu:84-124 */
int de6d_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx, const float *weight,
                           float *out, cudaStream_t stream);
/* three_interpolate_grad_wrapper(b, c, n, m, grad_out, idx, weight, grad_points)   interpolate.cpp:48-58, interpolate_gpu.cu:127-168
 * NOTE (b,c,n,m) order, as in the reference. */
int de6d_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx, const float *weight,
                                float *grad_points, cudaStream_t stream);

/* ---- iou3d_nms ---------------------------------------------------------------------------------------- */

/* boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap)   iou3d_nms.cpp:49-68, iou3d_nms_kernel.cu:236-249
 * boxes (N,7) [x,y,z,dx,dy,dz,heading] -> (na,nb) f32. */
int de6d_boxes_overlap_bev(int na, const float *boxes_a, int nb, const float *boxes_b, float *ans_overlap,
                           cudaStream_t stream);
/* boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou)   iou3d_nms.cpp:70-88, iou3d_nms_kernel.cu:251-265
 * Also serves boxes_iou_bev_cpu (iou3d_cpu.cpp:232-252): same arithmetic, run on the device. */
int de6d_boxes_iou_bev(int na, const float *boxes_a, int nb, const float *boxes_b, float *ans_iou, cudaStream_t stream);

/* boxes_iou_bev_cpu(boxes_a, boxes_b, ans_iou)   iou3d_cpu.cpp:232-252 -- HOST pointers, evaluated on the calling
 * thread (rows split over `nthreads` std::threads when > 1) with the reference's host arithmetic (libm trig): the
 * reference calls this from forked DataLoader workers (database_sampler.py:232-233) where no CUDA context can exist.
 * Bit-identical to the reference function.  Never touches the device. */
int de6d_boxes_iou_bev_host(int na, const float *boxes_a, int nb, const float *boxes_b, float *ans_iou, int nthreads);
/* Fused boxes_iou3d_gpu (python composition iou3d_nms_utils.py:48-81: BEV overlap x height overlap / union volume). */
int de6d_boxes_iou3d(int na, const float *boxes_a, int nb, const float *boxes_b, float *ans_iou, cudaStream_t stream);

/* nms_gpu / nms_normal_gpu (iou3d_nms.cpp:90-186, kernels iou3d_nms_kernel.cu:267-372), batched and
 * host-synchronisation free.  boxes (frames,n,7) sorted by descending score per frame; nvalid (frames) i32 or
 * NULL (= n each): boxes at positions >= nvalid[f] are ignored.  normal != 0 selects the axis-aligned IoU.
 * Outputs: keep (frames,n) int64 kept positions in ascending order (first num_keep[f] entries valid),
 * num_keep (frames) i32.  workspace: de6d_nms_workspace_bytes(frames,n) bytes, initialised once by
 * de6d_nms_workspace_init (the kernel leaves it reusable).  The reference's single-frame
 * nms_gpu(boxes, keep HOST int64, thresh) -> num_to_keep is this call with frames = 1 plus a D2H copy. */
size_t de6d_nms_workspace_bytes(int frames, int n);
int de6d_nms_workspace_init(int frames, void *workspace, cudaStream_t stream);
int de6d_nms_batched(int frames, int n, const float *boxes, const int *nvalid, float thresh, int mode,
                     long long *keep, int *num_keep, void *workspace, size_t workspace_bytes, cudaStream_t stream);
/* mode: 0 = rotated BEV IoU (nms_gpu), 1 = axis-aligned BEV IoU (nms_normal_gpu), boxes (frames, n, 7);
 *       2 = full-pose 3-D IoU, boxes (frames, n, 9) [x, y, z, dx, dy, dz, rz, ry, rx] (see de6d_boxes_iou3d9). */

/* Full-pose (9-DoF) IoU -- not in the reference, which evaluates boxes_iou3d_gpu / nms_gpu on boxes[:, 0:7] and so
 * ignores the pitch and roll Det6D predicts (point_head_box6d_vote.py:355, model_nms_utils.py:18).
 * boxes (n, 9) [x, y, z, dx, dy, dz, rz, ry, rx] with R = Rx Ry Rz as pcdet/utils/box_utils.py:59-72 builds corners
 * (scipy from_euler('zyx')); ans (na, nb) = volume(A n B) / max(vol A + vol B - volume(A n B), 1e-6). */
int de6d_boxes_iou3d9(int na, const float *boxes_a, int nb, const float *boxes_b, float *ans_iou, cudaStream_t stream);

/* ---- roiaware_pool3d ---------------------------------------------------------------------------------- */

/* points_in_boxes_gpu(boxes, pts, box_idx_of_points)   roiaware_pool3d.cpp:98-118, roiaware_pool3d_kernel.cu:313-359
 * NOTE boxes (b,t,7) before pts (b,m,3).  out (b,m) i32, caller fills -1; value = first containing box. */
int de6d_points_in_boxes(int b, int t, int m, const float *boxes, const float *pts, int *box_idx_of_points,
                         cudaStream_t stream);
/* points_in_boxes_cpu(boxes, pts, pts_indices)   roiaware_pool3d.cpp:143-168, run on the device:
 * boxes (t,7), pts (m,3) -> (t,m) i32 0/1 mask, MARGIN 1e-2, host (unfused) arithmetic. */
int de6d_points_in_boxes_mask(int t, int m, const float *boxes, const float *pts, int *point_indices,
                              cudaStream_t stream);

/* points_in_boxes_cpu(boxes, pts, pts_indices)   roiaware_pool3d.cpp:143-168 -- HOST pointers, calling thread
 * (+ optional row threads), MARGIN 1e-2, (t,m) 0/1 mask; for the reference's DataLoader-worker callers
 * (kitti_dataset.py:248, box_utils.py:104).  Bit-identical to the reference function.  Never touches the device. */
int de6d_points_in_boxes_mask_host(int t, int m, const float *boxes, const float *pts, int *point_indices, int nthreads);

/* ---- next to the path: one SA radius scale with the shared MLP fused behind the grouper ----------- */

/* Replaces, per scale, groupers[i] -> mlps[i] -> idx_cnt mask -> max_pool2d of _PointnetSAModuleFSBase.forward
 * (pointnet2_modules.py:461-478) after the ball query: gather, up to 4 [1x1 conv + folded BatchNorm + ReLU] layers on the
 * tensor cores (tcgen05.mma kind::tf32, activations in tensor memory), empty-ball mask, max over nsample -- the grouped
 * tensor (b, 3+c, m, nsample) and the hidden activations never reach HBM.
 * widths (HOST int array, n_layers + 1): [3 + c_feat, c_1, ..., c_L]; every c_l a multiple of 16, <= 256; nsample a power
 * of two in 4..128; all layers' tf32 weights must fit in one SM's shared memory (de6d_sa_mlp_fits returns 1) or, with widths
 * that are multiples of 32, in two SMs' (returns 2: the kernel then runs as cta_group::2 pairs); 0 = not supported.
 * de6d_sa_mlp_pack: weights_cat (device) = the layers' BatchNorm-folded matrices, row-major [c_{l+1} x c_l], concatenated,
 * layer 0's columns in the reference order (dx, dy, dz, features...); packed (device) = de6d_sa_mlp_packed_floats floats.
 * de6d_sa_mlp_fused: xyz (b,n,3), new_xyz (b,m,3), feats_pm (b,n,c_feat) POINT-major features (NULL when c_feat == 0),
 * idx (b,m,nsample), idx_cnt (b,m) or NULL, bias (device) = concatenated folded biases, out (b, c_L, m);
 * status (device int, may be NULL) is set to 1 if a tensor-core wait timed out (never in a correct build). */
int de6d_sa_mlp_fits(int n_layers, const int *widths, int nsample);
size_t de6d_sa_mlp_packed_floats(int n_layers, const int *widths);
int de6d_sa_mlp_pack(int n_layers, const int *widths, const float *weights_cat, float *packed, cudaStream_t stream);
int de6d_sa_mlp_fused(int b, int n, int m, int nsample, int c_feat, const float *xyz, const float *new_xyz,
                      const float *feats_pm, const int *idx, const int *idx_cnt, int n_layers, const int *widths,
                      const float *packed, const float *bias, float *out, int *status, cudaStream_t stream);
/* The same, writing channels [out_channel_offset, out_channel_offset + c_L) of an out tensor (b, out_channels, m): an MLP whose
 * LAST layer is too wide for the shared memory of an SM pair (131 -> 128 -> 256 -> 256: 464 KB of tf32 weights) runs as several
 * launches over row blocks of the last layer's weight matrix, each with its own packed image (widths[n_layers] = rows of the block). */
int de6d_sa_mlp_fused_slice(int b, int n, int m, int nsample, int c_feat, const float *xyz, const float *new_xyz,
                            const float *feats_pm, const int *idx, const int *idx_cnt, int n_layers, const int *widths,
                            const float *packed, const float *bias, float *out, int out_channels, int out_channel_offset,
                            int *status, cudaStream_t stream);

/* ---- next to the path: full-pose boxes -------------------------------------------------- */

/* box_utils.points_in_boxes3d(points, boxes3d)   pcdet/utils/box_utils.py:110-124 (host numpy + scipy Delaunay per box in
 * the reference; called by the Det6D head's target assignment, point_head_box6d_vote.py:198-209,284-286), batched:
 * boxes (b,t,9) [x,y,z,dx,dy,dz,rz,ry,rx] (scipy 'zyx' Euler convention), pts (b,m,3) -> out (b,m) int64 = index of the
 * LAST box containing the point (later boxes overwrite earlier ones, like the reference loop), -1 if none. */
int de6d_points_in_boxes9(int b, int t, int m, const float *boxes, const float *pts, long long *out, cudaStream_t stream);

/* ---- next to the path: input staging ----------------------------------------------------- */

/* DataProcessor.sample_points gather (datasets/processor/data_processor.py:145-177: points[choice]) + break_up_pc and
 * the per-frame count / view / permute copies of PointNet2FSMSG.forward (backbones_3d/pointnet2_backbone.py:193-222) in
 * one pass: src (total_rows, lead+3+c) f32 rows [batch_idx (lead=1 only), x, y, z, c features]; choice NULL (row bs*n+i,
 * requires total_rows == b*n) or (b,n) i32 global row indices -> xyz (b,n,3), features (b,c,n) (NULL allowed when c == 0),
 * batch_idx (b,n) f32 or NULL, status i32[2] or NULL (zeroed by the call): [0] = rows whose batch column differs from the
 * frame slot they were written to (the equal-count assertion of pointnet2_backbone.py:214-218 holds iff it is 0),
 * [1] = choice entries outside [0,total_rows) (those points are written as zeros). */
int de6d_stage_points(int b, int n, int c, int lead, long long total_rows, const float *src, const int *choice, float *xyz,
                      float *features, float *batch_idx, int *status, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DE6D_B200_H */
