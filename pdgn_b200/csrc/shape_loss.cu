// shape_loss.cu -- the whole shape-preserving loss side of one generator step (models/PDGNet_v2.py:232-237) in SIX launches.
//
// The trainer calls get_local_pair (PDGNet_v2.py:136-155) on the six level pairs (1,2) (1,3) (1,4) (2,3) (2,4) (3,4) of the
// generator's four outputs.  Each call runs kNN(pt1|pt1), kNN(pt2|pt1), two neighbourhood statistics and two Chamfer losses;
// the self problems kNN(pt1|pt1) + statistics repeat for every pair with the same first level, so one step holds 9 distinct
// kNN problems, 9 statistics problems and 24 directional-minimum problems.  pdgn_local_pair_fwd (local_pair.cu) already put one
// call behind one C entry point (11 launches); here every operator runs ALL its problems of the step from one descriptor
// table (multi.cuh):
//   forward : transpose x4 | kNN x9 | statistics x9 | minima d=3 x12 | minima d=6 x12 | 12 sums        = 6 launches
//   backward: (memset) | Chamfer adjoint d=3 x12 | d=6 x12 | statistics adjoint x9 | transpose-add x4  = 4 launches
// (d=6: the symmetric covariances travel packed as (xx, yy, zz, sqrt2 xy, sqrt2 xz, sqrt2 yz): same Frobenius distance, 6 channels)
// Values are those of the per-call path (same kernels' bodies); the shared self statistics are computed once.
#include <cstdint>
#include "common.cuh"
#include "multi.cuh"

namespace pdgn {

constexpr int SL_MAXL = 4;
constexpr int SL_MAXP = SL_MAXL * (SL_MAXL - 1) / 2;

struct SlLayout {  // offsets in 4-byte words
    size_t P[SL_MAXL], idxS[SL_MAXL], idxC[SL_MAXP], muS[SL_MAXL], covS[SL_MAXL], muC[SL_MAXP], covC[SL_MAXP], mn[SL_MAXP][8];
    size_t zero0, gmuS[SL_MAXL], gcovS[SL_MAXL], gmuC[SL_MAXP], gcovC[SL_MAXP], gP[SL_MAXL], zero1, total;
    int pa[SL_MAXP], pb[SL_MAXP], pairs;
};

static SlLayout sl_layout(int b, int levels, const int* n, int k) {
    SlLayout L{};
    size_t o = 0;
    auto take = [&](size_t words) { const size_t at = o; o += (words + 3) & ~(size_t)3; return at; };
    L.pairs = 0;
    for (int a = 0; a < levels; ++a)
        for (int c = a + 1; c < levels; ++c) { L.pa[L.pairs] = a; L.pb[L.pairs] = c; ++L.pairs; }
    for (int l = 0; l < levels; ++l) L.P[l] = take((size_t)b * n[l] * 3);
    for (int a = 0; a + 1 < levels; ++a) { L.idxS[a] = take((size_t)b * n[a] * k); L.muS[a] = take((size_t)b * n[a] * 3); L.covS[a] = take((size_t)b * n[a] * 9); }
    for (int p = 0; p < L.pairs; ++p) {
        const size_t bm = (size_t)b * n[L.pa[p]];
        L.idxC[p] = take(bm * k); L.muC[p] = take(bm * 3); L.covC[p] = take(bm * 9);
        for (int i = 0; i < 8; ++i) L.mn[p][i] = take(bm);
    }
    L.zero0 = o;
    for (int a = 0; a + 1 < levels; ++a) { L.gmuS[a] = take((size_t)b * n[a] * 3); L.gcovS[a] = take((size_t)b * n[a] * 9); }
    for (int p = 0; p < L.pairs; ++p) { const size_t bm = (size_t)b * n[L.pa[p]]; L.gmuC[p] = take(bm * 3); L.gcovC[p] = take(bm * 9); }
    for (int l = 0; l < levels; ++l) L.gP[l] = take((size_t)b * n[l] * 3);
    L.zero1 = o;
    L.total = o;
    return L;
}

struct SlSets { const float* src[SL_MAXL]; float* dst[SL_MAXL]; int n[SL_MAXL]; };

// [b,3,n] -> [b,n,3] for every level (blockIdx.z = level)
__global__ void sl_transpose_in_kernel(const __grid_constant__ SlSets s) {
    const int l = blockIdx.z, bz = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x, n = s.n[l];
    if (j >= n) return;
    const float* src = s.src[l] + (size_t)bz * 3 * n;
    float* d = s.dst[l] + ((size_t)bz * n + j) * 3;
    d[0] = src[j]; d[1] = src[n + j]; d[2] = src[2 * (size_t)n + j];
}
// grad [b,3,n] += g [b,n,3] for every level (src = workspace gradients, dst = caller's)
__global__ void sl_transpose_add_kernel(const __grid_constant__ SlSets s) {
    const int l = blockIdx.z, bz = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x, n = s.n[l];
    if (j >= n || !s.dst[l]) return;
    const float* g = s.src[l] + ((size_t)bz * n + j) * 3;
    float* d = s.dst[l] + (size_t)bz * 3 * n;
    d[j] += g[0]; d[n + j] += g[1]; d[2 * (size_t)n + j] += g[2];
}

struct SlSums { const float* a[2 * SL_MAXP]; const float* c[2 * SL_MAXP]; size_t count[2 * SL_MAXP]; float inv_m[2 * SL_MAXP]; };
// out[o] = (sum a[o] + sum c[o]) / m; one CTA per output, fixed summation order (deterministic)
__global__ void __launch_bounds__(1024) sl_sums_kernel(const __grid_constant__ SlSums s, float* __restrict__ out) {
    __shared__ float red[32];
    const int o = blockIdx.x;
    const float* a = s.a[o];
    const float* c = s.c[o];
    float acc = 0.f;
    for (size_t i = threadIdx.x; i < s.count[o]; i += 1024) acc += a[i] + c[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = warp_sum(red[threadIdx.x]);
        if (threadIdx.x == 0) out[o] = acc * s.inv_m[o];
    }
}

}  // namespace pdgn

using namespace pdgn;

static bool sl_args_ok(int b, int levels, const int* n, int k) {
    if (b <= 0 || levels < 2 || levels > SL_MAXL || !n || k < 1) return false;
    for (int l = 0; l < levels; ++l)
        if (n[l] <= 0) return false;
    return true;
}

extern "C" size_t pdgn_shape_loss_workspace(int b, int levels, const int* npts, int k) {
    if (!sl_args_ok(b, levels, npts, k)) return 0;
    return sl_layout(b, levels, npts, k).total * 4 + 256;
}

#define PDGN_SL_TRY(call)               \
    do {                                \
        const int rc_ = (call);         \
        if (rc_ != PDGN_OK) return rc_; \
    } while (0)

extern "C" int pdgn_shape_loss_fwd(const float* const* pts, int b, int levels, const int* npts, int k, float* out, void* workspace,
                                   size_t workspace_bytes, void* stream) {
    PDGN_RANGE("pdgn_shape_loss_fwd");
    if (!sl_args_ok(b, levels, npts, k) || !pts || !out) return PDGN_ERR_BAD_ARG;
    if (k > 64 || b > 65535) return PDGN_ERR_UNSUPPORTED;
    for (int l = 0; l < levels; ++l)
        if (!pts[l]) return PDGN_ERR_BAD_ARG;
    const SlLayout L = sl_layout(b, levels, npts, k);
    if (!workspace || workspace_bytes < L.total * 4 || (reinterpret_cast<uintptr_t>(workspace) & 15)) return PDGN_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    float* W = reinterpret_cast<float*>(workspace);
    int* WI = reinterpret_cast<int*>(workspace);
    int nmax = 0;
    SlSets sets{};
    for (int l = 0; l < levels; ++l) { sets.src[l] = pts[l]; sets.dst[l] = W + L.P[l]; sets.n[l] = npts[l]; nmax = npts[l] > nmax ? npts[l] : nmax; }
    sl_transpose_in_kernel<<<dim3((nmax + 255) / 256, b, levels), 256, 0, st>>>(sets);
    PDGN_CHECK_LAUNCH();
    // kNN: self problems (level a against itself) then cross problems (queries = level a's points, candidates = level c's)
    KnnTable kt{};
    StatTable stt{};
    for (int a = 0; a + 1 < levels; ++a) {
        kt.p[kt.count++] = KnnProb{W + L.P[a], W + L.P[a], WI + L.idxS[a], npts[a], npts[a], 0, 0, 0};
        stt.p[stt.count++] = StatProb{W + L.P[a], WI + L.idxS[a], W + L.muS[a], W + L.covS[a], nullptr, nullptr, nullptr, npts[a], npts[a], 0};
    }
    for (int p = 0; p < L.pairs; ++p) {
        const int a = L.pa[p], c = L.pb[p];
        kt.p[kt.count++] = KnnProb{W + L.P[c], W + L.P[a], WI + L.idxC[p], npts[c], npts[a], 0, 0, 0};
        stt.p[stt.count++] = StatProb{W + L.P[c], WI + L.idxC[p], W + L.muC[p], W + L.covC[p], nullptr, nullptr, nullptr, npts[c], npts[a], 0};
    }
    int rc = knn_multi_launch(kt, b, k, st);
    if (rc == PDGN_ERR_UNSUPPORTED) {  // a level outside the descriptor kernel's range (< 256 or > 2048 points): one launch per problem
        for (int i = 0; i < kt.count; ++i) PDGN_SL_TRY(pdgn_knn_xyz(kt.p[i].xyz, kt.p[i].q, b, kt.p[i].n, kt.p[i].m, k, kt.p[i].idx, nullptr, stream));
    } else if (rc != PDGN_OK) {
        return rc;
    }
    PDGN_SL_TRY(local_stats_multi_fwd(stt, b, k, st));
    // ChamferLoss(preds = cross statistics, gts = self statistics): both directional minima (chamfer_loss.py:13-20)
    MinTable m3{}, m9{};
    SlSums sums{};
    for (int p = 0; p < L.pairs; ++p) {
        const int a = L.pa[p], m = npts[a];
        const float inv_m = 1.0f / (float)m;
        m3.p[m3.count++] = MinProb{W + L.muC[p], W + L.muS[a], W + L.mn[p][0], WI + L.mn[p][1], nullptr, nullptr, nullptr, inv_m, m, m, 0};
        m3.p[m3.count++] = MinProb{W + L.muS[a], W + L.muC[p], W + L.mn[p][2], WI + L.mn[p][3], nullptr, nullptr, nullptr, inv_m, m, m, 0};
        m9.p[m9.count++] = MinProb{W + L.covC[p], W + L.covS[a], W + L.mn[p][4], WI + L.mn[p][5], nullptr, nullptr, nullptr, inv_m, m, m, 0};
        m9.p[m9.count++] = MinProb{W + L.covS[a], W + L.covC[p], W + L.mn[p][6], WI + L.mn[p][7], nullptr, nullptr, nullptr, inv_m, m, m, 0};
        sums.a[2 * p] = W + L.mn[p][0]; sums.c[2 * p] = W + L.mn[p][2];
        sums.a[2 * p + 1] = W + L.mn[p][4]; sums.c[2 * p + 1] = W + L.mn[p][6];
        sums.count[2 * p] = sums.count[2 * p + 1] = (size_t)b * m;
        sums.inv_m[2 * p] = sums.inv_m[2 * p + 1] = inv_m;
    }
    PDGN_SL_TRY(nn_min_multi_launch(m3, b, 3, st));
    PDGN_SL_TRY(nn_min_multi_launch(m9, b, 6, st));   // covariances in the packed 6-channel layout
    sl_sums_kernel<<<2 * L.pairs, 1024, 0, st>>>(sums, out);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}

extern "C" int pdgn_shape_loss_bwd(int b, int levels, const int* npts, int k, const float* grad_out, float* const* grad_pts,
                                   void* workspace, size_t workspace_bytes, void* stream) {
    PDGN_RANGE("pdgn_shape_loss_bwd");
    if (!sl_args_ok(b, levels, npts, k) || !grad_out || !grad_pts) return PDGN_ERR_BAD_ARG;
    if (k > 64 || b > 65535) return PDGN_ERR_UNSUPPORTED;
    const SlLayout L = sl_layout(b, levels, npts, k);
    if (!workspace || workspace_bytes < L.total * 4 || (reinterpret_cast<uintptr_t>(workspace) & 15)) return PDGN_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    float* W = reinterpret_cast<float*>(workspace);
    int* WI = reinterpret_cast<int*>(workspace);
    PDGN_CUDA(cudaMemsetAsync(W + L.zero0, 0, (L.zero1 - L.zero0) * 4, st));
    MinTable m3{}, m9{};
    for (int p = 0; p < L.pairs; ++p) {
        const int a = L.pa[p], m = npts[a];
        const float inv_m = 1.0f / (float)m;
        const float* gmu = grad_out + 2 * p;
        const float* gcv = grad_out + 2 * p + 1;
        m3.p[m3.count++] = MinProb{W + L.muC[p], W + L.muS[a], nullptr, WI + L.mn[p][1], W + L.gmuC[p], W + L.gmuS[a], gmu, inv_m, m, m, 0};
        m3.p[m3.count++] = MinProb{W + L.muS[a], W + L.muC[p], nullptr, WI + L.mn[p][3], W + L.gmuS[a], W + L.gmuC[p], gmu, inv_m, m, m, 0};
        m9.p[m9.count++] = MinProb{W + L.covC[p], W + L.covS[a], nullptr, WI + L.mn[p][5], W + L.gcovC[p], W + L.gcovS[a], gcv, inv_m, m, m, 0};
        m9.p[m9.count++] = MinProb{W + L.covS[a], W + L.covC[p], nullptr, WI + L.mn[p][7], W + L.gcovS[a], W + L.gcovC[p], gcv, inv_m, m, m, 0};
    }
    PDGN_SL_TRY(chamfer_bwd_multi_launch(m3, b, 3, st));
    PDGN_SL_TRY(chamfer_bwd_multi_launch(m9, b, 6, st));
    StatTable stt{};
    for (int a = 0; a + 1 < levels; ++a)
        stt.p[stt.count++] = StatProb{W + L.P[a], WI + L.idxS[a], W + L.muS[a], nullptr, W + L.gP[a], W + L.gmuS[a], W + L.gcovS[a], npts[a], npts[a], 0};
    for (int p = 0; p < L.pairs; ++p) {
        const int a = L.pa[p], c = L.pb[p];
        stt.p[stt.count++] = StatProb{W + L.P[c], WI + L.idxC[p], W + L.muC[p], nullptr, W + L.gP[c], W + L.gmuC[p], W + L.gcovC[p], npts[c], npts[a], 0};
    }
    PDGN_SL_TRY(local_stats_multi_bwd(stt, b, k, st));
    int nmax = 0;
    SlSets sets{};
    for (int l = 0; l < levels; ++l) { sets.src[l] = W + L.gP[l]; sets.dst[l] = grad_pts[l]; sets.n[l] = npts[l]; nmax = npts[l] > nmax ? npts[l] : nmax; }
    sl_transpose_add_kernel<<<dim3((nmax + 255) / 256, b, levels), 256, 0, st>>>(sets);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}
