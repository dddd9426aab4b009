// knn_feat.cu -- feature-space kNN graph of the generator (C = 32..256 channels, k = 10, "ranks 1..k").
//
// Replaces, in get_edge_features / get_edge_features_xyz (models/PDGNet_v2.py:449-459, :492-502), the
// [B,N,N] cuBLAS Gram matrix + full torch.sort of every row + slice.  Nothing N x N is materialised: a CTA owns
// 64 query points, walks the candidates in tiles of 64, accumulates the 64x64 block of squared distances in
// registers (4x4 per thread, channels streamed through shared memory) and folds each finished block into a
// per-query sorted list of skip+k entries.
//
// Arithmetic contract (SURVEY.md section 7 "Gram-form parity"; the reference's own indices depend on cuBLAS
// summation order and an unstable sort, so they are not a reproducible target): exact FP32 direct distances
// d2(i,j) = sum_c (x[c,i]-x[c,j])^2 accumulated as ONE fma chain over c = 0..C-1 per pair, total order
// (d2, index), ranks skip..skip+k-1.  Bit-exact against oracle_knn_feat (oracle/pdgn_oracle.c).
// FP32 SIMT: 2 FMA-pipe instructions per pair per channel.
#include "common.cuh"

namespace pdgn {

constexpr int FT = 256;   // threads
constexpr int FB = 64;    // block edge (queries per CTA, candidates per tile)
constexpr int FCK = 32;   // channels staged per step

// flags != nullptr: only the queries with flags[b*n + i] < 0 are computed and written (the tensor-core path's overflow list);
// a CTA without one returns at once.
__global__ void __launch_bounds__(FT) knn_feat_kernel(const float* __restrict__ x, int c, int n, int k, int skip,
                                                     long long* __restrict__ idx, float* __restrict__ dist2,
                                                     const int* __restrict__ flags, const int* __restrict__ nflag, int brute_max,
                                                     int brute_n_max) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* xi = reinterpret_cast<float*>(smem_raw);  // [FCK][FB]
    float* xj = xi + FCK * FB;                       // [FCK][FB]
    float* ds = xj + FCK * FB;                       // [FB][FB+1]
    const int kk = k + skip;
    float* ld = ds + FB * (FB + 1);                  // [kk][FB]
    int* li = reinterpret_cast<int*>(ld + (size_t)kk * FB);

    const int bz = blockIdx.y, t = threadIdx.x;
    const int i0 = blockIdx.x * FB;
    const int ty = t >> 4, tx = t & 15;
    const float* xb = x + (size_t)bz * c * n;
    bool mine = true;
    if (flags) {
        // the tensor-core path's re-rank kernel has already recomputed the flagged queries when they were few (and the cloud small)
        if (*nflag == 0 || (*nflag <= brute_max && n <= brute_n_max)) return;
        mine = t < FB && i0 + t < n && flags[(size_t)bz * n + i0 + t] < 0;
        if (!__syncthreads_or(mine ? 1 : 0)) return;
    }

    if (t < FB)
        for (int e = 0; e < kk; ++e) {
            ld[e * FB + t] = kInf;
            li[e * FB + t] = 0;
        }
    float thr = kInf;

    for (int j0 = 0; j0 < n; j0 += FB) {
        float acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int s = 0; s < 4; ++s) acc[r][s] = 0.f;

        for (int c0 = 0; c0 < c; c0 += FCK) {
            const int cc = min(FCK, c - c0);
            __syncthreads();
            for (int e = t; e < cc * FB; e += FT) {
                const int ch = e >> 6, p = e & (FB - 1);
                const float* row = xb + (size_t)(c0 + ch) * n;
                xi[e] = (i0 + p < n) ? row[i0 + p] : 0.f;
                xj[e] = (j0 + p < n) ? row[j0 + p] : 0.f;
            }
            __syncthreads();
#pragma unroll 4
            for (int ch = 0; ch < cc; ++ch) {
                const float4 a = *reinterpret_cast<const float4*>(xi + ch * FB + ty * 4);
                const float4 b = *reinterpret_cast<const float4*>(xj + ch * FB + tx * 4);
                const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const float diff = __fsub_rn(av[r], bv[s]);
                        acc[r][s] = __fmaf_rn(diff, diff, acc[r][s]);
                    }
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int s = 0; s < 4; ++s) ds[(ty * 4 + r) * (FB + 1) + tx * 4 + s] = acc[r][s];
        __syncthreads();
        if (t < FB && i0 + t < n) {
            const int jn = min(FB, n - j0);
            for (int j = 0; j < jn; ++j) {
                const float d = ds[t * (FB + 1) + j];
                if (d < thr) {
                    list_insert(ld, li, FB, t, kk, d, j0 + j);
                    thr = ld[(kk - 1) * FB + t];
                }
            }
        }
    }
    if (t < FB && i0 + t < n && mine) {
        const size_t o = ((size_t)bz * n + i0 + t) * k;
        for (int e = 0; e < k; ++e) {
            idx[o + e] = li[(skip + e) * FB + t];
            if (dist2) dist2[o + e] = ld[(skip + e) * FB + t];
        }
    }
}

}  // namespace pdgn

using namespace pdgn;

extern "C" int pdgn_knn_feat(const float* x, int b, int c, int n, int k, int skip, int64_t* idx, float* dist2, void* stream) {
    PDGN_RANGE("pdgn_knn_feat");
    if (!x || !idx || b < 0 || c < 1 || n < 0 || k < 1 || skip < 0) return PDGN_ERR_BAD_ARG;
    if (k + skip > 64 || b > 65535) return PDGN_ERR_UNSUPPORTED;
    if (k + skip > n) return PDGN_ERR_BAD_ARG;  // the reference's slice [1:k+1] would come up short
    if (b == 0) return PDGN_OK;
    const size_t smem = (size_t)(2 * FCK * FB + FB * (FB + 1)) * 4 + (size_t)(k + skip) * FB * 8;
    PDGN_CUDA(cudaFuncSetAttribute(knn_feat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((n + FB - 1) / FB, b);
    knn_feat_kernel<<<grid, FT, smem, (cudaStream_t)stream>>>(x, c, n, k, skip, reinterpret_cast<long long*>(idx), dist2, nullptr, nullptr, 0, 0);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}

namespace pdgn {
bool knn_feat_tc_eligible(int c, int n, int k, int skip);
int knn_feat_tc_brute_max();
size_t knn_feat_tc_workspace(int b, int c, int n);
int knn_feat_tc_launch(const float* x, int b, int c, int n, int k, int skip, long long* idx, float* dist2, void* ws, const int** flags,
                       const int** nflag_out, int* brute_n_max, cudaStream_t st);
}  // namespace pdgn

// Workspace form: the tensor-core filter + exact re-rank of knn_feat_tc.cu for the shapes it takes (8 <= c <= 256, c % 8 == 0,
// 128 <= n <= 4096, n % 128 == 0, k + skip <= 20), the kernel above otherwise.  Same results either way.
extern "C" size_t pdgn_knn_feat_workspace(int b, int c, int n) {
    if (b < 0 || c < 1 || n < 0) return 0;
    return knn_feat_tc_workspace(b, c, n);
}

extern "C" int pdgn_knn_feat_ws(const float* x, int b, int c, int n, int k, int skip, int64_t* idx, float* dist2, void* workspace,
                                size_t workspace_bytes, void* stream) {
    PDGN_RANGE("pdgn_knn_feat_ws");
    if (!x || !idx || b < 0 || c < 1 || n < 0 || k < 1 || skip < 0) return PDGN_ERR_BAD_ARG;
    static const char* impl = tune_env("PDGN_KNN_FEAT_IMPL");      // "simt": always the FP32 SIMT kernel (A/B runs)
    if (!workspace || !knn_feat_tc_eligible(c, n, k, skip) || b > 65535 || b == 0 || (impl && impl[0] == 's'))
        return pdgn_knn_feat(x, b, c, n, k, skip, idx, dist2, stream);
    if (workspace_bytes < knn_feat_tc_workspace(b, c, n) - 256) return PDGN_ERR_WORKSPACE;
    const int *flags = nullptr, *nflag = nullptr;
    int brute_n_max = 0;
    const int rc = knn_feat_tc_launch(x, b, c, n, k, skip, reinterpret_cast<long long*>(idx), dist2, workspace, &flags, &nflag, &brute_n_max,
                                      (cudaStream_t)stream);
    if (rc != PDGN_OK || !flags) return rc;
    const size_t smem = (size_t)(2 * FCK * FB + FB * (FB + 1)) * 4 + (size_t)(k + skip) * FB * 8;
    PDGN_CUDA(cudaFuncSetAttribute(knn_feat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((n + FB - 1) / FB, b);
    knn_feat_kernel<<<grid, FT, smem, (cudaStream_t)stream>>>(x, c, n, k, skip, reinterpret_cast<long long*>(idx), dist2, flags, nflag, knn_feat_tc_brute_max(), brute_n_max);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}
