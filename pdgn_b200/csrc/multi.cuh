// multi.cuh -- problem-descriptor launches: ONE kernel launch runs several independent (shape, pointer) problems of the same
// operator.  Used by shape_loss.cu to run all the kNN / neighbourhood-statistics / Chamfer-minimum problems of one generator
// step (models/PDGNet_v2.py:232-237: six get_local_pair calls = 9 distinct kNN problems, 9 statistics, 24 directional minima)
// in one launch per operator.  The descriptor table travels as a __grid_constant__ kernel parameter (no device allocation);
// CTAs are laid out problem after problem along grid.x (`cta0` = first CTA of the problem), batch elements along grid.y.
#pragma once
#include "common.cuh"

namespace pdgn {

constexpr int MULTI_MAX = 24;

struct KnnProb { const float* xyz; const float* q; int* idx; int n, m, log2ss, gsz, cta0; };
struct KnnTable { KnnProb p[12]; int count; };

struct StatProb { const float* xyz; const int* idx; float* mu; float* cov; float* gxyz; const float* gmu; const float* gcov; int n, m, cta0; };
struct StatTable { StatProb p[12]; int count; };

// directional minimum x -> y (for every x point the nearest y point); `gx`, `gy`, `gscale` are used by the backward only:
// d/dx += 2 * (*gscale) * inv_m * (x - y[arg]), d/dy -= the same
struct MinProb { const float* x; const float* y; float* mind; int* argm; float* gx; float* gy; const float* gscale; float inv_m; int nx, ny, cta0; };
struct MinTable { MinProb p[MULTI_MAX]; int count; };

template <typename Table>
__device__ __forceinline__ int multi_find(const Table& tb, int bx) {
    int i = 0;
#pragma unroll 1
    while (i + 1 < tb.count && bx >= tb.p[i + 1].cta0) ++i;
    return i;
}

// knn_xyz.cu: every problem must satisfy 256 <= n <= 2048 (single resident tile) and k <= 24; returns PDGN_ERR_UNSUPPORTED
// otherwise (the caller then falls back to one pdgn_knn_xyz call per problem).
int knn_multi_launch(KnnTable& tb, int b, int k, cudaStream_t st);
// local_stats.cu: the multi launches write / read the covariance in the packed 6-channel layout (see local_stats_fwd_body)
int local_stats_multi_fwd(StatTable& tb, int b, int k, cudaStream_t st);
int local_stats_fwd_launch(const float* xyz, const int* idx, int b, int n, int m, int k, float* mu, float* cov, int cw, cudaStream_t st);
int local_stats_bwd_launch(const float* xyz, const int* idx, const float* mu, const float* grad_mu, const float* grad_cov, int b, int n,
                           int m, int k, float* grad_xyz, int cw, cudaStream_t st);
int local_stats_multi_bwd(StatTable& tb, int b, int k, cudaStream_t st);
// chamfer.cu: all problems of one call share the channel count d (3, 6 or 9)
int nn_min_multi_launch(MinTable& tb, int b, int d, cudaStream_t st);
int chamfer_bwd_multi_launch(MinTable& tb, int b, int d, cudaStream_t st);

}  // namespace pdgn
