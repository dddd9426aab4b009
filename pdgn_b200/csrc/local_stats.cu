// local_stats.cu -- fused neighbourhood statistics of get_local_pair (next row SURVEY.md 8f-2).
//
// The reference's shape-preserving loss (models/PDGNet_v2.py:136-155) groups the k nearest neighbours of every query
// (Gen_QueryAndGroupXYZ: knnquery -> transpose -> grouping, pointops.py:682-703), reshapes the [B,3,M,k] tensor to
// [(B*M),3,k] and computes mean and covariance with dense torch ops (compute_mean_covariance, :127-134: mean, repeat,
// subtract, bmm, divide) -- ten small launches and a [B,3,M,k] round trip through HBM per call, twelve calls per G step.
// Here one kernel reads the indices and writes mu [B,M,3] and cov [B,M,9] directly; the backward is its exact adjoint
// (scatter-add of d(mu,cov)/d(point) through the same indices).
//   mu_a    = (1/k) sum_s x[idx_s][a]
//   cov_ab  = (1/k) sum_s (x[idx_s][a] - mu_a)(x[idx_s][b] - mu_b)
//   dL/dx[idx_s][c] += gmu_c / k + (1/k) sum_b (gcov_cb + gcov_bc) (x[idx_s][b] - mu_b)      (sum_s (x_s - mu) = 0)
#include "common.cuh"
#include "multi.cuh"

namespace pdgn {

constexpr int LS_T = 128;
constexpr int LS_KMAX = 64;

// cw = 9: covariance as the reference lays it out (row-major 3x3).  cw = 6 (fused paths only): the symmetric matrix packed as
// (xx, yy, zz, sqrt2 xy, sqrt2 xz, sqrt2 yz) -- the same Frobenius distances in 6 channels instead of 9, a third less work for
// the Chamfer minima on the covariances.
__device__ __forceinline__ void local_stats_fwd_body(const float* __restrict__ xyz, const int* __restrict__ idx, int n, int m, int k,
                                                     float* __restrict__ mu, float* __restrict__ cov, int cw, int bx, int bz) {
    const int j = bx * LS_T + threadIdx.x;
    if (j >= m) return;
    const float* pb = xyz + (size_t)bz * n * 3;
    const int* ip = idx + ((size_t)bz * m + j) * k;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int s = 0; s < k; ++s) {
        const float* p = pb + (size_t)ip[s] * 3;
        sx += __ldg(p); sy += __ldg(p + 1); sz += __ldg(p + 2);
    }
    const float inv = 1.0f / (float)k;
    const float mx = sx * inv, my = sy * inv, mz = sz * inv;
    float cxx = 0.f, cxy = 0.f, cxz = 0.f, cyy = 0.f, cyz = 0.f, czz = 0.f;
    for (int s = 0; s < k; ++s) {
        const float* p = pb + (size_t)ip[s] * 3;
        const float tx = __ldg(p) - mx, ty = __ldg(p + 1) - my, tz = __ldg(p + 2) - mz;
        cxx = fmaf(tx, tx, cxx); cxy = fmaf(tx, ty, cxy); cxz = fmaf(tx, tz, cxz);
        cyy = fmaf(ty, ty, cyy); cyz = fmaf(ty, tz, cyz); czz = fmaf(tz, tz, czz);
    }
    float* mo = mu + ((size_t)bz * m + j) * 3;
    mo[0] = mx; mo[1] = my; mo[2] = mz;
    if (cw == 6) {
        constexpr float kSqrt2 = 1.41421356237309505f;
        float* co = cov + ((size_t)bz * m + j) * 6;
        co[0] = cxx * inv; co[1] = cyy * inv; co[2] = czz * inv;
        co[3] = kSqrt2 * (cxy * inv); co[4] = kSqrt2 * (cxz * inv); co[5] = kSqrt2 * (cyz * inv);
        return;
    }
    float* co = cov + ((size_t)bz * m + j) * 9;
    co[0] = cxx * inv; co[1] = cxy * inv; co[2] = cxz * inv;
    co[3] = cxy * inv; co[4] = cyy * inv; co[5] = cyz * inv;
    co[6] = cxz * inv; co[7] = cyz * inv; co[8] = czz * inv;
}

__global__ void __launch_bounds__(LS_T) local_stats_fwd_kernel(const float* __restrict__ xyz, const int* __restrict__ idx, int n, int m,
                                                              int k, float* __restrict__ mu, float* __restrict__ cov, int cw) {
    local_stats_fwd_body(xyz, idx, n, m, k, mu, cov, cw, blockIdx.x, blockIdx.y);
}

__device__ __forceinline__ void local_stats_bwd_body(const float* __restrict__ xyz, const int* __restrict__ idx,
                                                     const float* __restrict__ mu, const float* __restrict__ gmu,
                                                     const float* __restrict__ gcov, int n, int m, int k, float* __restrict__ gxyz,
                                                     int cw, int bx, int bz) {
    const int j = bx * LS_T + threadIdx.x;
    if (j >= m) return;
    const float* pb = xyz + (size_t)bz * n * 3;
    float* gb = gxyz + (size_t)bz * n * 3;
    const int* ip = idx + ((size_t)bz * m + j) * k;
    const float* mo = mu + ((size_t)bz * m + j) * 3;
    const float* gm = gmu + ((size_t)bz * m + j) * 3;
    const float inv = 1.0f / (float)k;
    const float mx = mo[0], my = mo[1], mz = mo[2];
    // symmetrised covariance gradient: S_cb = gcov_cb + gcov_bc (packed layout: d/d(xy) = sqrt2 * d/d(packed xy))
    float sxx, sxy, sxz, syy, syz, szz;
    if (cw == 6) {
        constexpr float kSqrt2 = 1.41421356237309505f;
        const float* gc = gcov + ((size_t)bz * m + j) * 6;
        sxx = 2.f * gc[0]; syy = 2.f * gc[1]; szz = 2.f * gc[2];
        sxy = kSqrt2 * gc[3]; sxz = kSqrt2 * gc[4]; syz = kSqrt2 * gc[5];
    } else {
        const float* gc = gcov + ((size_t)bz * m + j) * 9;
        sxx = 2.f * gc[0]; sxy = gc[1] + gc[3]; sxz = gc[2] + gc[6]; syy = 2.f * gc[4]; syz = gc[5] + gc[7]; szz = 2.f * gc[8];
    }
    const float gx0 = gm[0] * inv, gy0 = gm[1] * inv, gz0 = gm[2] * inv;
    for (int s = 0; s < k; ++s) {
        const int pi = ip[s];
        const float* p = pb + (size_t)pi * 3;
        const float tx = __ldg(p) - mx, ty = __ldg(p + 1) - my, tz = __ldg(p + 2) - mz;
        atomicAdd(gb + (size_t)pi * 3, gx0 + inv * (sxx * tx + sxy * ty + sxz * tz));
        atomicAdd(gb + (size_t)pi * 3 + 1, gy0 + inv * (sxy * tx + syy * ty + syz * tz));
        atomicAdd(gb + (size_t)pi * 3 + 2, gz0 + inv * (sxz * tx + syz * ty + szz * tz));
    }
}

__global__ void __launch_bounds__(LS_T) local_stats_bwd_kernel(const float* __restrict__ xyz, const int* __restrict__ idx,
                                                              const float* __restrict__ mu, const float* __restrict__ gmu,
                                                              const float* __restrict__ gcov, int n, int m, int k,
                                                              float* __restrict__ gxyz, int cw) {
    local_stats_bwd_body(xyz, idx, mu, gmu, gcov, n, m, k, gxyz, cw, blockIdx.x, blockIdx.y);
}

// problem-descriptor launches (multi.cuh)
__global__ void __launch_bounds__(LS_T) local_stats_multi_fwd_kernel(const __grid_constant__ StatTable tb, int k) {
    const StatProb& pr = tb.p[multi_find(tb, blockIdx.x)];
    local_stats_fwd_body(pr.xyz, pr.idx, pr.n, pr.m, k, pr.mu, pr.cov, 6, blockIdx.x - pr.cta0, blockIdx.y);
}
__global__ void __launch_bounds__(LS_T) local_stats_multi_bwd_kernel(const __grid_constant__ StatTable tb, int k) {
    const StatProb& pr = tb.p[multi_find(tb, blockIdx.x)];
    local_stats_bwd_body(pr.xyz, pr.idx, pr.mu, pr.gmu, pr.gcov, pr.n, pr.m, k, pr.gxyz, 6, blockIdx.x - pr.cta0, blockIdx.y);
}

// single problem, either covariance layout (the fused per-call path uses the packed one)
int local_stats_fwd_launch(const float* xyz, const int* idx, int b, int n, int m, int k, float* mu, float* cov, int cw, cudaStream_t st) {
    if (b < 0 || n < 0 || m < 0 || k < 1 || (cw != 6 && cw != 9)) return PDGN_ERR_BAD_ARG;
    if (k > LS_KMAX || b > 65535) return PDGN_ERR_UNSUPPORTED;
    if (b == 0 || m == 0) return PDGN_OK;
    if (!xyz || !idx || !mu || !cov || n == 0) return PDGN_ERR_BAD_ARG;
    PDGN_VERIFY_IDX32(idx, (size_t)b * m * k, n, st);
    local_stats_fwd_kernel<<<dim3((m + LS_T - 1) / LS_T, b), LS_T, 0, st>>>(xyz, idx, n, m, k, mu, cov, cw);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}
int local_stats_bwd_launch(const float* xyz, const int* idx, const float* mu, const float* grad_mu, const float* grad_cov, int b, int n,
                           int m, int k, float* grad_xyz, int cw, cudaStream_t st) {
    if (b < 0 || n < 0 || m < 0 || k < 1 || (cw != 6 && cw != 9)) return PDGN_ERR_BAD_ARG;
    if (k > LS_KMAX || b > 65535) return PDGN_ERR_UNSUPPORTED;
    if (b == 0 || m == 0) return PDGN_OK;
    if (!xyz || !idx || !mu || !grad_mu || !grad_cov || !grad_xyz || n == 0) return PDGN_ERR_BAD_ARG;
    PDGN_VERIFY_IDX32(idx, (size_t)b * m * k, n, st);
    local_stats_bwd_kernel<<<dim3((m + LS_T - 1) / LS_T, b), LS_T, 0, st>>>(xyz, idx, mu, grad_mu, grad_cov, n, m, k, grad_xyz, cw);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}

static int stat_table_ctas(StatTable& tb) {
    int ctas = 0;
    for (int i = 0; i < tb.count; ++i) {
        tb.p[i].cta0 = ctas;
        ctas += (tb.p[i].m + LS_T - 1) / LS_T;
    }
    return ctas;
}
int local_stats_multi_fwd(StatTable& tb, int b, int k, cudaStream_t st) {
    if (k < 1 || k > LS_KMAX || tb.count < 1 || tb.count > 12) return PDGN_ERR_UNSUPPORTED;
    local_stats_multi_fwd_kernel<<<dim3(stat_table_ctas(tb), b), LS_T, 0, st>>>(tb, k);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}
int local_stats_multi_bwd(StatTable& tb, int b, int k, cudaStream_t st) {
    if (k < 1 || k > LS_KMAX || tb.count < 1 || tb.count > 12) return PDGN_ERR_UNSUPPORTED;
    local_stats_multi_bwd_kernel<<<dim3(stat_table_ctas(tb), b), LS_T, 0, st>>>(tb, k);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}

}  // namespace pdgn

using namespace pdgn;

extern "C" int pdgn_local_stats_fwd(const float* xyz, const int* idx, int b, int n, int m, int k, float* mu, float* cov, void* stream) {
    PDGN_RANGE("pdgn_local_stats_fwd");
    return local_stats_fwd_launch(xyz, idx, b, n, m, k, mu, cov, 9, (cudaStream_t)stream);
}

extern "C" int pdgn_local_stats_bwd(const float* xyz, const int* idx, const float* mu, const float* grad_mu, const float* grad_cov,
                                    int b, int n, int m, int k, float* grad_xyz, void* stream) {
    PDGN_RANGE("pdgn_local_stats_bwd");
    return local_stats_bwd_launch(xyz, idx, mu, grad_mu, grad_cov, b, n, m, k, grad_xyz, 9, (cudaStream_t)stream);
}
