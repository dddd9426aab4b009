// local_pair.cu -- the whole loss side of one get_local_pair call (models/PDGNet_v2.py:136-155) behind ONE C call per
// direction (next row SURVEY.md 8f-2).
//
// pdgn_b200.local_pair.get_local_pair already replaced the reference's ~45 launches by ~12, but issued from Python it stays
// host bound at the training shapes (B=35, 256..2048 points: ~0.7 ms of wall clock per call for ~0.3 ms of kernels, twelve
// calls per G step).  Here the chain
//   transpose -> kNN(pt1|pt1), kNN(pt2|pt1) -> neighbourhood mean/covariance -> both Chamfer minima on (mu, cov) -> two sums
// is enqueued by one entry point on the caller's stream, every intermediate lives in a caller-provided workspace that the
// backward entry point reuses (indices, statistics, arg-minima), and nothing synchronises the host.
#include <cstdint>
#include "common.cuh"
#include "multi.cuh"

namespace pdgn {

struct LpLayout {  // offsets in 4-byte words into the workspace
    size_t p1, p2, idx1, idx2, mu1, cov1, mu2, cov2, mn[8], w, gmu1, gcov1, gmu2, gcov2, gp1, gp2, total;
};

static LpLayout lp_layout(int b, int m, int n, int k) {
    LpLayout L{};
    size_t o = 0;
    auto take = [&](size_t words) { const size_t at = o; o += (words + 3) & ~(size_t)3; return at; };  // 16-byte aligned blocks
    const size_t bm = (size_t)b * m, bn = (size_t)b * n;
    L.p1 = take(bm * 3); L.p2 = take(bn * 3);
    L.idx1 = take(bm * k); L.idx2 = take(bm * k);
    L.mu1 = take(bm * 3); L.cov1 = take(bm * 9); L.mu2 = take(bm * 3); L.cov2 = take(bm * 9);
    for (int i = 0; i < 8; ++i) L.mn[i] = take(bm);  // (min, arg) x (mu2->mu1, mu1->mu2, cov2->cov1, cov1->cov2)
    // backward scratch (contiguous: zeroed with one memset)
    L.gmu1 = take(bm * 3); L.gcov1 = take(bm * 9); L.gmu2 = take(bm * 3); L.gcov2 = take(bm * 9);
    L.gp1 = take(bm * 3); L.gp2 = take(bn * 3);
    L.w = take(2 * bm);
    L.total = o;
    return L;
}

// [b,3,n] -> [b,n,3]
__global__ void lp_transpose_in_kernel(const float* __restrict__ src, int n, float* __restrict__ dst) {
    const int bz = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float* s = src + (size_t)bz * 3 * n;
    float* d = dst + ((size_t)bz * n + j) * 3;
    d[0] = s[j]; d[1] = s[n + j]; d[2] = s[2 * (size_t)n + j];
}

// grad [b,3,n] += g [b,n,3]
__global__ void lp_transpose_add_kernel(const float* __restrict__ g, int n, float* __restrict__ grad) {
    const int bz = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float* s = g + ((size_t)bz * n + j) * 3;
    float* d = grad + (size_t)bz * 3 * n;
    d[j] += s[0]; d[n + j] += s[1]; d[2 * (size_t)n + j] += s[2];
}

// out[0] = (sum a0 + sum a1) / m, out[1] = (sum a2 + sum a3) / m; one CTA per output, fixed summation order (deterministic).
// Four independent partial sums per thread keep 8 loads in flight (the arrays are L2 resident: the loop is latency bound).
__global__ void __launch_bounds__(1024) lp_sums_kernel(const float* __restrict__ a0, const float* __restrict__ a1, const float* __restrict__ a2,
                                                      const float* __restrict__ a3, size_t count, float inv_m, float* __restrict__ out) {
    __shared__ float red[32];
    const float* p = blockIdx.x == 0 ? a0 : a2;
    const float* q = blockIdx.x == 0 ? a1 : a3;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    size_t i = threadIdx.x;
    for (; i + 3 * 1024 < count; i += 4 * 1024) {
#pragma unroll
        for (int u = 0; u < 4; ++u) s[u] += p[i + u * 1024] + q[i + u * 1024];
    }
    for (; i < count; i += 1024) s[0] += p[i] + q[i];
    float t = warp_sum((s[0] + s[1]) + (s[2] + s[3]));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = t;
    __syncthreads();
    if (warp == 0) {
        t = warp_sum(red[lane]);
        if (lane == 0) out[blockIdx.x] = t * inv_m;
    }
}

// w[0..count) = gout[0] / m, w[count..2 count) = gout[1] / m
__global__ void lp_fill_w_kernel(const float* __restrict__ gout, size_t count, float inv_m, float* __restrict__ w) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2 * count) w[i] = gout[i < count ? 0 : 1] * inv_m;
}

}  // namespace pdgn

using namespace pdgn;

extern "C" size_t pdgn_local_pair_workspace(int b, int m, int n, int k) {
    if (b < 0 || m < 0 || n < 0 || k < 1) return 0;
    return lp_layout(b, m, n, k).total * 4 + 256;
}

#define PDGN_LP_TRY(call)              \
    do {                               \
        const int rc_ = (call);        \
        if (rc_ != PDGN_OK) return rc_; \
    } while (0)

extern "C" int pdgn_local_pair_fwd(const float* pt1, const float* pt2, int b, int m, int n, int k, float* out, void* workspace,
                                   size_t workspace_bytes, void* stream) {
    PDGN_RANGE("pdgn_local_pair_fwd");
    if (b < 0 || m < 0 || n < 0 || k < 1) return PDGN_ERR_BAD_ARG;
    if (k > 64 || b > 65535) return PDGN_ERR_UNSUPPORTED;
    if (!out) return PDGN_ERR_BAD_ARG;
    if (b == 0 || m == 0) return PDGN_ERR_BAD_ARG;  // the reference divides by m and takes minima over empty sets
    if (!pt1 || !pt2 || n <= 0) return PDGN_ERR_BAD_ARG;
    const LpLayout L = lp_layout(b, m, n, k);
    if (!workspace || workspace_bytes < L.total * 4 || (reinterpret_cast<uintptr_t>(workspace) & 15)) return PDGN_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    float* W = reinterpret_cast<float*>(workspace);
    int* WI = reinterpret_cast<int*>(workspace);
    lp_transpose_in_kernel<<<dim3((m + 255) / 256, b), 256, 0, st>>>(pt1, m, W + L.p1);
    PDGN_CHECK_LAUNCH();
    lp_transpose_in_kernel<<<dim3((n + 255) / 256, b), 256, 0, st>>>(pt2, n, W + L.p2);
    PDGN_CHECK_LAUNCH();
    // Gen_QueryAndGroupXYZ(pt1, pt1) and (pt2, pt1): the queries are pt1's points in both (PDGNet_v2.py:139-146)
    PDGN_LP_TRY(pdgn_knn_xyz(W + L.p1, W + L.p1, b, m, m, k, WI + L.idx1, nullptr, stream));
    PDGN_LP_TRY(pdgn_knn_xyz(W + L.p2, W + L.p1, b, n, m, k, WI + L.idx2, nullptr, stream));
    // covariances in the packed 6-channel layout (local_stats.cu): same Frobenius distances, a third less Chamfer work
    PDGN_LP_TRY(local_stats_fwd_launch(W + L.p1, WI + L.idx1, b, m, m, k, W + L.mu1, W + L.cov1, 6, st));
    PDGN_LP_TRY(local_stats_fwd_launch(W + L.p2, WI + L.idx2, b, n, m, k, W + L.mu2, W + L.cov2, 6, st));
    // ChamferLoss(preds = stats of pt2, gts = stats of pt1): both directional minima (chamfer_loss.py:13-20)
    PDGN_LP_TRY(pdgn_chamfer_min(W + L.mu2, W + L.mu1, b, m, m, 3, W + L.mn[0], WI + L.mn[1], W + L.mn[2], WI + L.mn[3], stream));
    PDGN_LP_TRY(pdgn_chamfer_min(W + L.cov2, W + L.cov1, b, m, m, 6, W + L.mn[4], WI + L.mn[5], W + L.mn[6], WI + L.mn[7], stream));
    lp_sums_kernel<<<2, 1024, 0, st>>>(W + L.mn[0], W + L.mn[2], W + L.mn[4], W + L.mn[6], (size_t)b * m, 1.0f / (float)m, out);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}

extern "C" int pdgn_local_pair_bwd(int b, int m, int n, int k, const float* grad_out, float* grad_pt1, float* grad_pt2,
                                   void* workspace, size_t workspace_bytes, void* stream) {
    PDGN_RANGE("pdgn_local_pair_bwd");
    if (b <= 0 || m <= 0 || n <= 0 || k < 1 || !grad_out || !grad_pt1 || !grad_pt2) return PDGN_ERR_BAD_ARG;
    if (k > 64 || b > 65535) return PDGN_ERR_UNSUPPORTED;
    const LpLayout L = lp_layout(b, m, n, k);
    if (!workspace || workspace_bytes < L.total * 4 || (reinterpret_cast<uintptr_t>(workspace) & 15)) return PDGN_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    float* W = reinterpret_cast<float*>(workspace);
    int* WI = reinterpret_cast<int*>(workspace);
    const size_t bm = (size_t)b * m;
    PDGN_CUDA(cudaMemsetAsync(W + L.gmu1, 0, (L.w - L.gmu1) * 4, st));  // gmu1 .. gp2 are contiguous
    lp_fill_w_kernel<<<(unsigned)((2 * bm + 255) / 256), 256, 0, st>>>(grad_out, bm, 1.0f / (float)m, W + L.w);
    PDGN_CHECK_LAUNCH();
    PDGN_LP_TRY(pdgn_chamfer_bwd(W + L.mu2, W + L.mu1, b, m, m, 3, W + L.w, WI + L.mn[1], W + L.w, WI + L.mn[3], W + L.gmu2, W + L.gmu1, stream));
    PDGN_LP_TRY(pdgn_chamfer_bwd(W + L.cov2, W + L.cov1, b, m, m, 6, W + L.w + bm, WI + L.mn[5], W + L.w + bm, WI + L.mn[7], W + L.gcov2,
                                 W + L.gcov1, stream));
    PDGN_LP_TRY(local_stats_bwd_launch(W + L.p1, WI + L.idx1, W + L.mu1, W + L.gmu1, W + L.gcov1, b, m, m, k, W + L.gp1, 6, st));
    PDGN_LP_TRY(local_stats_bwd_launch(W + L.p2, WI + L.idx2, W + L.mu2, W + L.gmu2, W + L.gcov2, b, n, m, k, W + L.gp2, 6, st));
    lp_transpose_add_kernel<<<dim3((m + 255) / 256, b), 256, 0, st>>>(W + L.gp1, m, grad_pt1);
    PDGN_CHECK_LAUNCH();
    lp_transpose_add_kernel<<<dim3((n + 255) / 256, b), 256, 0, st>>>(W + L.gp2, n, grad_pt2);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}
