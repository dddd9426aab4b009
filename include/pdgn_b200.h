/*
 * pdgn_b200.h -- C ABI of libpdgn_b200.so: the B200 (sm_100a) implementation of PDGN's nearest-neighbour /
 * distance hot path.  This is the drop-in boundary: each entry point replaces one native function the
 * reference binds (cited per function, paths under the reference tree fpthink/PDGN).
 *
 * Conventions (all entry points)
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer on the current device unless the
 *     name ends in _host; tensors are dense row-major ("contiguous") FP32 / int32 / int64 as stated;
 *   - the caller allocates every output (as lib/pointops/functions/pointops.py does); accumulating
 *     outputs (grad_*) are ADDED into, so the caller zero-fills them first (pointops.py:146,114);
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*, NULL = legacy default stream); no
 *     entry point synchronises the host, allocates device memory or exits the process;
 *   - return value: 0 on success, PDGN_ERR_* (negative) for argument errors, or a positive cudaError_t
 *     from the launch.  pdgn_error_string() turns either into text.
 *   - environment: PDGN_B200_VERIFY=1 range-checks every index tensor before it is used (one extra kernel and a stream
 *     synchronisation per gather call; PDGN_ERR_INDEX on failure).  Every entry point opens an NVTX range of its own name.
 *     Kernel-variant tuning hooks are inert unless PDGN_B200_TUNE=1.
 */
#ifndef PDGN_B200_H
#define PDGN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDGN_ABI_VERSION 1

#define PDGN_OK 0
#define PDGN_ERR_BAD_ARG (-1)      /* null pointer / negative size */
#define PDGN_ERR_UNSUPPORTED (-2)  /* size outside what the kernels implement (stated per function) */
#define PDGN_ERR_WORKSPACE (-3)    /* workspace too small */
#define PDGN_ERR_INDEX (-4)        /* PDGN_B200_VERIFY=1 only: an index tensor holds a value outside [0, n) */

int pdgn_abi_version(void);
const char *pdgn_error_string(int code);

/* ---- k nearest neighbours in xyz ------------------------------------------------------------------
 * Replaces knnquery_cuda_launcher(b,n,m,nsample,xyz,new_xyz,idx,dist2,stream)
 *   (lib/pointops/src/knnquery/knnquery_cuda_kernel.h:14, kernel knnquery_cuda_kernel.cu:6-50).
 * xyz [b,n,3] references, new_xyz [b,m,3] queries -> idx int32 [b,m,k], dist2 f32 [b,m,k] (may be NULL).
 * Result order: ascending (d2, index) with d2 = fma(dz,dz, fma(dx,dx, dy*dy)) -- bit-identical to the
 * reference kernel as compiled by nvcc 12.9 -O2.  n < k: trailing idx 0 / dist2 +inf.  NaN or +inf
 * distances are never selected.  1 <= k <= 200 (the reference's own limit: best_dist[200], knnquery_cuda_kernel.cu:21-22). */
int pdgn_knn_xyz(const float *xyz, const float *new_xyz, int b, int n, int m, int k, int *idx, float *dist2,
                 void *stream);

/* ---- three nearest neighbours ---------------------------------------------------------------------
 * Replaces nearestneighbor_cuda_launcher_fast(b,n,m,unknown,known,dist2,idx)
 *   (lib/pointops/src/interpolation/interpolation_cuda_kernel.h:26, kernel .cu:134-176).
 * unknown [b,n,3], known [b,m,3] -> dist2 f32 [b,n,3] (SQUARED, the caller takes sqrt as pointops.py:77
 * does), idx int32 [b,n,3]. */
int pdgn_nn3(const float *unknown, const float *known, int b, int n, int m, float *dist2, int *idx, void *stream);

/* ---- grouping (neighbour gather) and its backward scatter-add --------------------------------------
 * Replace grouping_forward_cuda_launcher_fast / grouping_backward_cuda_launcher
 *   (lib/pointops/src/grouping/grouping_cuda_kernel.h:17-19, kernels .cu:60-75 and :28-46).
 * fwd: points [b,c,n], idx int32 [b,m,k] -> out [b,c,m,k];  out[b,c,j,s] = points[b,c,idx[b,j,s]]
 * bwd: grad_out [b,c,m,k], idx -> grad_points [b,c,n] += scatter (FP32 atomics, like the reference). */
int pdgn_group_fwd(const float *points, const int *idx, int b, int c, int n, int m, int k, float *out, void *stream);
int pdgn_group_bwd(const float *grad_out, const int *idx, int b, int c, int n, int m, int k, float *grad_points,
                   void *stream);
/* Same result through a deterministic pull (inverse index built in `workspace`, pdgn_group_bwd_workspace(b,n,m,k)
 * bytes; each target sums its contributions in ascending position order, no float atomics).  Falls back to the
 * atomic kernel for c < 4 or rows that do not fit shared memory. */
size_t pdgn_group_bwd_workspace(int b, int n, int m, int k);
int pdgn_group_bwd_ws(const float *grad_out, const int *idx, int b, int c, int n, int m, int k, float *grad_points,
                      void *workspace, size_t workspace_bytes, void *stream);

/* ---- three-point interpolation and its backward -----------------------------------------------------
 * Replace interpolation_forward_cuda_launcher_fast / interpolation_backward_cuda_launcher
 *   (interpolation_cuda_kernel.h:27,24, kernels .cu:181-195 and :90-114).
 * fwd: points [b,c,m], idx int32 [b,n,3], weight [b,n,3] -> out [b,c,n]
 *      out = fma(w2,p2, fma(w0,p0, w1*p1))  (the reference's compiled order)
 * bwd: grad_out [b,c,n] -> grad_points [b,c,m] += g*w_t at idx_t. */
int pdgn_interp_fwd(const float *points, const int *idx, const float *weight, int b, int c, int m, int n, float *out,
                    void *stream);
int pdgn_interp_bwd(const float *grad_out, const int *idx, const float *weight, int b, int c, int n, int m,
                    float *grad_points, void *stream);
/* Deterministic pull form of the backward (see pdgn_group_bwd_ws); workspace = pdgn_interp_bwd_workspace(b,n,m) bytes. */
size_t pdgn_interp_bwd_workspace(int b, int n, int m);
int pdgn_interp_bwd_ws(const float *grad_out, const int *idx, const float *weight, int b, int c, int n, int m,
                       float *grad_points, void *workspace, size_t workspace_bytes, void *stream);

/* ---- directional nearest-neighbour distance (Chamfer building block) --------------------------------
 * Replaces nndistance(b,n,xyz,m,xyz2,result,result_i,result2,result2_i,stream)
 *   (evaluation/pytorch_structural_losses/src/nndistance.cuh:1, kernel nndistance.cu:2-128) for d == 3, and
 *   the bmm/min composition of utils/chamfer_loss.py:13-38 / evaluation_metrics.py:35-45 for any d.
 * x [b,nx,d], y [b,ny,d] -> min_xy f32 [b,nx] = min_j |x_i-y_j|^2, arg_xy int32 [b,nx] (lowest index among
 * equal minima), and the same for y against x.  Either direction's outputs may be NULL (both pointers)
 * to skip it; arg pointers may be NULL alone.  1 <= d <= 16. */
int pdgn_chamfer_min(const float *x, const float *y, int b, int nx, int ny, int d, float *min_xy, int *arg_xy,
                     float *min_yx, int *arg_yx, void *stream);

/* Gradient of sum_i w_xy[i]*min_xy[i] + sum_j w_yx[j]*min_yx[j] with respect to x and y
 * (replaces nndistancegrad, nndistance.cu:129-154, generalised to any d).  grad_x [b,nx,d] and
 * grad_y [b,ny,d] are ADDED into.  w_* are the upstream gradients [b,nx] / [b,ny]. */
int pdgn_chamfer_bwd(const float *x, const float *y, int b, int nx, int ny, int d, const float *w_xy,
                     const int *arg_xy, const float *w_yx, const int *arg_yx, float *grad_x, float *grad_y,
                     void *stream);

/* ---- all-pairs Chamfer-distance matrix ---------------------------------------------------------------
 * Replaces the Python double loop _pairwise_EMD_CD_ -> distChamfer (evaluation/evaluation_metrics.py:85-121,
 * :35-45), CD half.  A [na,npts,3], B [nb,npts,3] (both resident on this device) ->
 *   out[(s-row0)*ld_out + (r-col0)] = mean_i min_j |A_s,i - B_r,j|^2 + mean_j min_i |A_s,i - B_r,j|^2
 * for s in [row0,row1), r in [col0,col1): one rank's 2-D tile of the pair grid.
 * workspace: pdgn_cd_allpairs_workspace(na, nb, npts) bytes of device scratch (packed cloud copies).
 * 1 <= npts <= 16384. */
size_t pdgn_cd_allpairs_workspace(int na, int nb, int npts);
int pdgn_cd_allpairs(const float *A, const float *B, int na, int nb, int npts, int row0, int row1, int col0, int col1,
                     float *out, long long ld_out, void *workspace, size_t workspace_bytes, void *stream);

/* Host-buffer form of the same call (the end-to-end path: H2D of both cloud sets, the tile, D2H of the
 * scalars; synchronises `stream` before returning).  A_host/B_host/out_host are HOST pointers. */
int pdgn_cd_allpairs_host(const float *A_host, const float *B_host, int na, int nb, int npts, int row0, int row1,
                          int col0, int col1, float *out_host, long long ld_out, void *stream);

/* ---- all-pairs approximate Earth Mover's Distance -------------------------------------------------------
 * Replaces, for evaluation, ApproxMatch + MatchCost (evaluation/pytorch_structural_losses/src/approxmatch.cu:3-224,
 * bound as StructuralLossesBackend.ApproxMatch / MatchCost, pybind/bind.cpp:10-16) as emd_approx composes them
 * (evaluation/evaluation_metrics.py:26-31) inside _pairwise_EMD_CD_ (:110).  A [na,n,3], B [nb,m,3] ->
 *   out[(s-row0)*ld_out + (r-col0)] = match_cost(A_s, B_r) / n.
 * The n x m match matrix is never materialised.  n, m <= 2048.  No gradient (evaluation only).
 * workspace: pdgn_emd_allpairs_workspace(row1-row0, col1-col0, n, m) bytes of 16-byte-aligned device scratch (the tile's
 * clouds in kd-tree leaf order, which lets the kernel skip far-apart point blocks whose weights are exact zeros). */
size_t pdgn_emd_allpairs_workspace(int nrows, int ncols, int n, int m);
int pdgn_emd_allpairs(const float *A, const float *B, int na, int nb, int n, int m, int row0, int row1, int col0,
                      int col1, float *out, long long ld_out, void *workspace, size_t workspace_bytes, void *stream);

/* Paired form of the same kernel: out[i] = match_cost(A_i, B_i) / n for i in [0,b) -- what MatchCostFunction.forward
 * (evaluation/pytorch_structural_losses/match_cost.py:10-24) returns for a batch, divided by n as emd_approx does
 * (evaluation_metrics.py:26-31); used by EMD_CD (:48-82).  One CTA per pair, all pairs in one launch. */
size_t pdgn_emd_paired_workspace(int b, int n, int m);
int pdgn_emd_paired(const float *A, const float *B, int b, int n, int m, float *out, void *workspace,
                    size_t workspace_bytes, void *stream);

/* ---- fused neighbourhood statistics (next row: the loss-side of get_local_pair) -------------------------------
 * Replaces grouping + transpose/view + compute_mean_covariance (lib/pointops/functions/pointops.py:699-703,
 * models/PDGNet_v2.py:127-134,142-147) for given kNN indices: xyz [b,n,3], idx int32 [b,m,k] ->
 * mu [b,m,3] (mean of the k neighbours), cov [b,m,9] (their covariance, row-major 3x3, divided by k).
 * bwd: grad_xyz [b,n,3] += adjoint of (grad_mu, grad_cov); `mu` is the forward output.  1 <= k <= 64. */
int pdgn_local_stats_fwd(const float *xyz, const int *idx, int b, int n, int m, int k, float *mu, float *cov, void *stream);
int pdgn_local_stats_bwd(const float *xyz, const int *idx, const float *mu, const float *grad_mu, const float *grad_cov,
                         int b, int n, int m, int k, float *grad_xyz, void *stream);

/* ---- the whole loss side of get_local_pair in one call (next row) ---------------------------------------------
 * Replaces get_local_pair (models/PDGNet_v2.py:136-155): Gen_QueryAndGroupXYZ on (pt1, pt1) and (pt2, pt1)
 * (lib/pointops/functions/pointops.py:670-703), compute_mean_covariance (:127-134) and the two ChamferLoss calls
 * (utils/chamfer_loss.py:13-38) divided by M.  pt1 [b,3,m], pt2 [b,3,n] (channel-first, as the generator emits them) ->
 * out[0] = like_mu12, out[1] = like_var12 (device scalars).  Every intermediate (transposed clouds, kNN indices,
 * statistics, arg-minima) stays in `workspace` (pdgn_local_pair_workspace(b,m,n,k) bytes, 16-byte aligned), which the
 * backward call reuses: grad_pt1 [b,3,m] and grad_pt2 [b,3,n] are ADDED into, grad_out[2] holds the upstream gradients
 * of the two scalars (device).  1 <= k <= 64; m, n >= 1. */
size_t pdgn_local_pair_workspace(int b, int m, int n, int k);
int pdgn_local_pair_fwd(const float *pt1, const float *pt2, int b, int m, int n, int k, float *out, void *workspace,
                        size_t workspace_bytes, void *stream);
int pdgn_local_pair_bwd(int b, int m, int n, int k, const float *grad_out, float *grad_pt1, float *grad_pt2,
                        void *workspace, size_t workspace_bytes, void *stream);

/* ---- the shape-preserving loss of one generator step, every operator in ONE launch ---------------------------------
 * Replaces the six get_local_pair calls of PDGNet_v2.train (models/PDGNet_v2.py:232-237) on the generator's `levels` outputs
 * pts[l] [b,3,npts[l]] (device pointers, l = 0..levels-1, 2 <= levels <= 4): for every level pair (a < c), in the reference's
 * order (0,1) (0,2) (0,3) (1,2) (1,3) (2,3), out[2p] = like_mu and out[2p+1] = like_var of get_local_pair(pts[a], pts[c]).
 * The 9 kNN, 9 statistics and 24 directional-minimum problems run from problem-descriptor tables: 6 launches forward, 4 +
 * one memset backward.  Same conventions as pdgn_local_pair_*: the workspace (pdgn_shape_loss_workspace bytes, 16-byte
 * aligned) carries indices / statistics / arg-minima to the backward call; grad_pts[l] [b,3,npts[l]] are ADDED into (a
 * NULL entry skips that level); grad_out[2 * pairs] (device) holds the upstream gradients.  1 <= k <= 64. */
size_t pdgn_shape_loss_workspace(int b, int levels, const int *npts, int k);
int pdgn_shape_loss_fwd(const float *const *pts, int b, int levels, const int *npts, int k, float *out, void *workspace,
                        size_t workspace_bytes, void *stream);
int pdgn_shape_loss_bwd(int b, int levels, const int *npts, int k, const float *grad_out, float *const *grad_pts,
                        void *workspace, size_t workspace_bytes, void *stream);

/* ---- feature-space kNN of the generator ---------------------------------------------------------------
 * Replaces bmm + torch.sort + slice in get_edge_features{,_xyz} (models/PDGNet_v2.py:449-459, :492-502).
 * x [b,c,n] -> idx int64 [b,n,k]: ranks skip..skip+k-1 of the ascending (d2, index) order of exact FP32
 * direct distances d2(i,j) = sum_c (x[c,i]-x[c,j])^2 accumulated as an fma chain over c (skip=1 reproduces
 * the reference's "drop rank 0").  dist2 f32 [b,n,k] may be NULL.  skip+k <= min(n, 64). */
int pdgn_knn_feat(const float *x, int b, int c, int n, int k, int skip, int64_t *idx, float *dist2, void *stream);
/* Workspace form of pdgn_knn_feat (same results): for 8 <= c <= 256 (c % 8 == 0), 128 <= n <= 4096 (n % 128 == 0), k + skip <= 20
 * the pairwise work runs on the tensor cores (tcgen05 TF32 Gram tiles as a FILTER with rigorous margins) and only the ~13
 * surviving candidates of a query are re-ranked with the exact FP32 chain; other shapes take pdgn_knn_feat's kernel.
 * Replaces the same torch.bmm + torch.sort (models/PDGNet_v2.py:449-459, :492-502).  workspace: device memory of
 * pdgn_knn_feat_workspace(b, c, n) bytes. */
size_t pdgn_knn_feat_workspace(int b, int c, int n);
int pdgn_knn_feat_ws(const float *x, int b, int c, int n, int k, int skip, int64_t *idx, float *dist2, void *workspace,
                     size_t workspace_bytes, void *stream);

/* Edge-feature gather: replaces the index_select loop + repeat + cat (PDGNet_v2.py:461-477, :505-525).
 * x [b,c,n], idx int64 [b,n,k] -> ee [b,2c,n,k] = cat(x_i broadcast over k, x_idx - x_i) on dim 1.
 * bwd: grad_ee [b,2c,n,k] -> grad_x [b,c,n] ADDED into. */
int pdgn_edge_feat_fwd(const float *x, const int64_t *idx, int b, int c, int n, int k, float *ee, void *stream);
int pdgn_edge_feat_bwd(const float *grad_ee, const int64_t *idx, int b, int c, int n, int k, float *grad_x,
                       void *stream);
/* Deterministic pull form (see pdgn_group_bwd_ws); workspace = pdgn_edge_feat_bwd_workspace(b,n,k) bytes. */
size_t pdgn_edge_feat_bwd_workspace(int b, int n, int k);
int pdgn_edge_feat_bwd_ws(const float *grad_ee, const int64_t *idx, int b, int c, int n, int k, float *grad_x,
                          void *workspace, size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PDGN_B200_H */
