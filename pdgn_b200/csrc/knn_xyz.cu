// knn_xyz.cu -- batched brute-force k nearest neighbours in xyz (C = 3), bit-exact with the reference.
//
// Replaces knnquery_cuda_kernel (lib/pointops/src/knnquery/knnquery_cuda_kernel.cu:6-50: one thread per query,
// insertion sort in 2400 B of local memory, FP64 compares) and nearestneighbor_cuda_kernel_fast
// (lib/pointops/src/interpolation/interpolation_cuda_kernel.cu:134-176).  Result contract = the reference's:
// ascending (d2, index), d2 by the native FMUL/FFMA/FFMA chain (d2_xyz), NaN/+inf distances never selected,
// missing neighbours reported as idx 0 / dist2 +inf.
//
// Fast path (knn_fast_kernel), FP32 SIMT, two scans of the candidate set per query, no sorted structure in the
// hot loops:
//   pass 1  every thread keeps KF_R queries in registers and scans the candidates, staged in shared memory as
//           SoA planes (warp-broadcast LDS.128 = 4 candidates per plane per load).  It only tracks the MINIMUM
//           distance inside each of G contiguous candidate groups (FMNMX3, 0.5 instr per pair).
//   bound   the k-th smallest of the G group minima is an upper bound tau on the k-th nearest distance (k
//           distinct groups each hold a candidate <= tau).  It is found with a register-resident bitonic
//           sorting network over the G minima (FMNMX only, branch free, all lanes busy).
//   pass 2  rescan; candidates with d2 <= tau (typically ~1.5 k of them) are appended, in index order, to a small
//           per-query queue of 16-bit indices in shared memory (predicated STS).
//   final   each thread insertion-sorts its queue by (d2, index) into a k-entry list and writes it out.  A queue
//           overflow (adversarial ties / clustering) makes that thread redo its query with the exact generic
//           scan, so the result is always exact.
// Generic path (knn_generic_kernel): any k <= 128, any n; threshold-guarded insertion into a shared-memory list.
#include "common.cuh"

namespace pdgn {

// =====================================================================================================
// generic exact kernel
// =====================================================================================================
constexpr int KG_T = 128;
constexpr int KG_TILE = 1024;

__global__ void __launch_bounds__(KG_T) knn_generic_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, int n,
                                                          int m, int k, int* __restrict__ idx, float* __restrict__ dist2) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tile = reinterpret_cast<float*>(smem_raw);       // [3][KG_TILE]
    float* ld = tile + 3 * KG_TILE;                         // [k][KG_T]
    int* li = reinterpret_cast<int*>(ld + (size_t)k * KG_T);  // [k][KG_T]
    const int bz = blockIdx.y, t = threadIdx.x;
    const int q = blockIdx.x * KG_T + t;
    const int qc = min(q, m - 1);
    const float* qp = new_xyz + ((size_t)bz * m + qc) * 3;
    const float qx = qp[0], qy = qp[1], qz = qp[2];
    for (int e = 0; e < k; ++e) {
        ld[e * KG_T + t] = kInf;
        li[e * KG_T + t] = 0;
    }
    float thr = kInf;
    const float* pb = xyz + (size_t)bz * n * 3;
    for (int j0 = 0; j0 < n; j0 += KG_TILE) {
        const int cnt = min(KG_TILE, n - j0);
        __syncthreads();
        for (int e = t; e < cnt * 3; e += KG_T) {
            const int j = e / 3, c = e - j * 3;
            tile[c * KG_TILE + j] = pb[(size_t)j0 * 3 + e];
        }
        __syncthreads();
        for (int j = 0; j < cnt; ++j) {
            const float d = d2_xyz(qx, qy, qz, tile[j], tile[KG_TILE + j], tile[2 * KG_TILE + j]);
            if (d < thr) {
                list_insert(ld, li, KG_T, t, k, d, j0 + j);
                thr = ld[(k - 1) * KG_T + t];
            }
        }
    }
    if (q < m) {
        const size_t o = ((size_t)bz * m + q) * k;
        for (int e = 0; e < k; ++e) {
            idx[o + e] = li[e * KG_T + t];
            if (dist2) dist2[o + e] = ld[e * KG_T + t];
        }
    }
}

// =====================================================================================================
// fast two-pass kernel
// =====================================================================================================
constexpr int KF_T = 128;            // threads per CTA
constexpr int KF_R = 2;              // queries per thread
constexpr int KF_Q = KF_T * KF_R;    // queries per CTA
constexpr int KF_TILE = 2048;        // candidate capacity of the shared tile

template <int N>
__device__ __forceinline__ void bitonic_sort_regs(float (&v)[N]) {
#pragma unroll
    for (int size = 2; size <= N; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const int p = i ^ stride;
                if (p > i) {
                    const bool up = (i & size) == 0;
                    const float lo = fminf(v[i], v[p]), hi = fmaxf(v[i], v[p]);
                    v[i] = up ? lo : hi;
                    v[p] = up ? hi : lo;
                }
            }
        }
    }
}

// Cooperative load of candidates [j0, j0+cnt) of one batch element, AoS global -> SoA shared, tail padded to a
// multiple of 4 with NaN (a NaN distance is never a minimum and never <= tau).
__device__ __forceinline__ void kf_load_tile(float* tile, const float* __restrict__ pb, int j0, int cnt, int t) {
    for (int e = t; e < cnt * 3; e += KF_T) {
        const int j = e / 3, c = e - j * 3;
        tile[c * KF_TILE + j] = pb[(size_t)j0 * 3 + e];
    }
    const int cnt4 = (cnt + 3) & ~3;
    if (t < cnt4 - cnt) {
        const float nanv = __int_as_float(0x7fc00000);
        tile[cnt + t] = nanv;
        tile[KF_TILE + cnt + t] = nanv;
        tile[2 * KF_TILE + cnt + t] = nanv;
    }
}

template <int G, int QCAP>
__global__ void __launch_bounds__(KF_T) knn_fast_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, int n,
                                                       int m, int k, int gs, int* __restrict__ idx, float* __restrict__ dist2) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tile = reinterpret_cast<float*>(smem_raw);                         // [3][KF_TILE]
    unsigned short* queue = reinterpret_cast<unsigned short*>(tile + 3 * KF_TILE);  // [QCAP][KF_Q]
    float* gm = reinterpret_cast<float*>(queue + (size_t)QCAP * KF_Q);         // [G][KF_Q]   (pass 1 / bound)
    float* ld = gm;                                                           // [k][KF_Q]   (final; overlays gm)
    int* li = reinterpret_cast<int*>(ld + (size_t)k * KF_Q);                   // [k][KF_Q]

    const int bz = blockIdx.y, t = threadIdx.x;
    const int qbase = blockIdx.x * KF_Q;
    const float* pb = xyz + (size_t)bz * n * 3;
    float qx[KF_R], qy[KF_R], qz[KF_R];
#pragma unroll
    for (int r = 0; r < KF_R; ++r) {
        const int q = min(qbase + r * KF_T + t, m - 1);  // slot r*KF_T + t  (lane-contiguous shared columns)
        const float* qp = new_xyz + ((size_t)bz * m + q) * 3;
        qx[r] = qp[0]; qy[r] = qp[1]; qz[r] = qp[2];
    }
    const int ng = (n + gs - 1) / gs;             // non-empty groups (<= G)
    const int tile_cands = (KF_TILE / gs) * gs;   // whole groups per tile

    // ---------------- pass 1: group minima
    for (int g = ng; g < G; ++g) {
#pragma unroll
        for (int r = 0; r < KF_R; ++r) gm[g * KF_Q + r * KF_T + t] = kInf;
    }
    for (int j0 = 0; j0 < n; j0 += tile_cands) {
        const int cnt = min(tile_cands, n - j0);
        __syncthreads();
        kf_load_tile(tile, pb, j0, cnt, t);
        __syncthreads();
        for (int gj = 0; gj < cnt; gj += gs) {
            const int gend = min(cnt, gj + gs);
            float mn[KF_R];
#pragma unroll
            for (int r = 0; r < KF_R; ++r) mn[r] = kInf;
#pragma unroll 2
            for (int j = gj; j < gend; j += 4) {
                const float4 X = *reinterpret_cast<const float4*>(tile + j);
                const float4 Y = *reinterpret_cast<const float4*>(tile + KF_TILE + j);
                const float4 Z = *reinterpret_cast<const float4*>(tile + 2 * KF_TILE + j);
#pragma unroll
                for (int r = 0; r < KF_R; ++r) {
                    const float d0 = d2_xyz(qx[r], qy[r], qz[r], X.x, Y.x, Z.x), d1 = d2_xyz(qx[r], qy[r], qz[r], X.y, Y.y, Z.y);
                    const float d2 = d2_xyz(qx[r], qy[r], qz[r], X.z, Y.z, Z.z), d3 = d2_xyz(qx[r], qy[r], qz[r], X.w, Y.w, Z.w);
                    mn[r] = min3(min3(mn[r], d0, d1), d2, d3);
                }
            }
            const int g = (j0 + gj) / gs;
#pragma unroll
            for (int r = 0; r < KF_R; ++r) gm[g * KF_Q + r * KF_T + t] = mn[r];
        }
    }

    // ---------------- bound: tau = k-th smallest group minimum (own columns only: no barrier needed)
    float tau[KF_R];
#pragma unroll
    for (int r = 0; r < KF_R; ++r) {
        float v[G];
#pragma unroll
        for (int g = 0; g < G; ++g) v[g] = gm[g * KF_Q + r * KF_T + t];
        bitonic_sort_regs<G>(v);
        float tv = kInf;
#pragma unroll
        for (int g = 0; g < G; ++g)
            if (g == k - 1) tv = v[g];
        tau[r] = tv;
    }

    // ---------------- pass 2: append every candidate with d2 <= tau (index order)
    int cnt_q[KF_R];
#pragma unroll
    for (int r = 0; r < KF_R; ++r) cnt_q[r] = 0;
    const bool single_tile = n <= tile_cands;
    for (int j0 = 0; j0 < n; j0 += tile_cands) {
        const int cnt = min(tile_cands, n - j0);
        if (!single_tile) {
            __syncthreads();
            kf_load_tile(tile, pb, j0, cnt, t);
            __syncthreads();
        }
        const int cnt4 = (cnt + 3) & ~3;
#pragma unroll 2
        for (int j = 0; j < cnt4; j += 4) {
            const float4 X = *reinterpret_cast<const float4*>(tile + j);
            const float4 Y = *reinterpret_cast<const float4*>(tile + KF_TILE + j);
            const float4 Z = *reinterpret_cast<const float4*>(tile + 2 * KF_TILE + j);
#pragma unroll
            for (int r = 0; r < KF_R; ++r) {
                float d[4];
                d[0] = d2_xyz(qx[r], qy[r], qz[r], X.x, Y.x, Z.x);
                d[1] = d2_xyz(qx[r], qy[r], qz[r], X.y, Y.y, Z.y);
                d[2] = d2_xyz(qx[r], qy[r], qz[r], X.z, Y.z, Z.z);
                d[3] = d2_xyz(qx[r], qy[r], qz[r], X.w, Y.w, Z.w);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (d[u] <= tau[r]) {
                        if (cnt_q[r] < QCAP) queue[cnt_q[r] * KF_Q + r * KF_T + t] = (unsigned short)(j0 + j + u);
                        ++cnt_q[r];
                    }
                }
            }
        }
    }
    __syncthreads();  // every thread is done with gm before the lists (which overlay it) are written

    // ---------------- final: exact (d2, index) order among the survivors
#pragma unroll
    for (int r = 0; r < KF_R; ++r) {
        const int col = r * KF_T + t;
        for (int e = 0; e < k; ++e) {
            ld[e * KF_Q + col] = kInf;
            li[e * KF_Q + col] = 0;
        }
        float thr = kInf;
        if (cnt_q[r] <= QCAP) {
            for (int e = 0; e < cnt_q[r]; ++e) {
                const int j = queue[e * KF_Q + col];
                const float* p = pb + (size_t)j * 3;
                const float d = d2_xyz(qx[r], qy[r], qz[r], __ldg(p), __ldg(p + 1), __ldg(p + 2));
                if (d < thr) {
                    list_insert(ld, li, KF_Q, col, k, d, j);
                    thr = ld[(k - 1) * KF_Q + col];
                }
            }
        } else {  // queue overflow: exact rescan of the whole candidate set for this query only
            for (int j = 0; j < n; ++j) {
                const float* p = pb + (size_t)j * 3;
                const float d = d2_xyz(qx[r], qy[r], qz[r], __ldg(p), __ldg(p + 1), __ldg(p + 2));
                if (d < thr) {
                    list_insert(ld, li, KF_Q, col, k, d, j);
                    thr = ld[(k - 1) * KF_Q + col];
                }
            }
        }
        const int q = qbase + col;
        if (q < m) {
            const size_t o = ((size_t)bz * m + q) * k;
            for (int e = 0; e < k; ++e) {
                idx[o + e] = li[e * KF_Q + col];
                if (dist2) dist2[o + e] = ld[e * KF_Q + col];
            }
        }
    }
}

template <int G, int QCAP>
static int launch_fast(const float* xyz, const float* new_xyz, int b, int n, int m, int k, int gs, int* idx, float* dist2,
                       cudaStream_t st) {
    const size_t overlay = (size_t)KF_Q * 4 * (size_t)max(G, 2 * k);
    const size_t smem = (size_t)3 * KF_TILE * 4 + (size_t)QCAP * KF_Q * 2 + overlay;
    PDGN_CUDA(cudaFuncSetAttribute(knn_fast_kernel<G, QCAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((m + KF_Q - 1) / KF_Q, b);
    knn_fast_kernel<G, QCAP><<<grid, KF_T, smem, st>>>(xyz, new_xyz, n, m, k, gs, idx, dist2);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}

static inline int group_size(int n, int G) { return max(4, ((n + G - 1) / G + 3) & ~3); }

static int knn_dispatch(const float* xyz, const float* new_xyz, int b, int n, int m, int k, int* idx, float* dist2, cudaStream_t st) {
    // Fast path needs: >= k non-empty groups (else tau = +inf), 16-bit candidate indices, whole groups per tile.
    if (n <= 65535) {
        int gs = group_size(n, 32);
        if (k <= 20 && (n + gs - 1) / gs >= k && gs <= KF_TILE) return launch_fast<32, 64>(xyz, new_xyz, b, n, m, k, gs, idx, dist2, st);
        gs = group_size(n, 64);
        if (k <= 40 && (n + gs - 1) / gs >= k && gs <= KF_TILE) return launch_fast<64, 128>(xyz, new_xyz, b, n, m, k, gs, idx, dist2, st);
    }
    const size_t smem = (size_t)3 * KG_TILE * 4 + (size_t)k * KG_T * 8;
    PDGN_CUDA(cudaFuncSetAttribute(knn_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((m + KG_T - 1) / KG_T, b);
    knn_generic_kernel<<<grid, KG_T, smem, st>>>(xyz, new_xyz, n, m, k, idx, dist2);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}

}  // namespace pdgn

using namespace pdgn;

extern "C" int pdgn_knn_xyz(const float* xyz, const float* new_xyz, int b, int n, int m, int k, int* idx, float* dist2, void* stream) {
    if (!xyz || !new_xyz || !idx || b < 0 || n < 0 || m < 0 || k < 1) return PDGN_ERR_BAD_ARG;
    if (k > 128 || b > 65535) return PDGN_ERR_UNSUPPORTED;
    if (b == 0 || m == 0) return PDGN_OK;
    // n == 0 falls through to the generic kernel, which leaves the reference's initial values (idx 0, dist +inf)
    return knn_dispatch(xyz, new_xyz, b, n, m, k, idx, dist2, (cudaStream_t)stream);
}

extern "C" int pdgn_nn3(const float* unknown, const float* known, int b, int n, int m, float* dist2, int* idx, void* stream) {
    if (!unknown || !known || !dist2 || !idx || b < 0 || n < 0 || m <= 0) return PDGN_ERR_BAD_ARG;
    if (b > 65535) return PDGN_ERR_UNSUPPORTED;
    if (b == 0 || n == 0) return PDGN_OK;
    return knn_dispatch(known, unknown, b, m, n, 3, idx, dist2, (cudaStream_t)stream);
}
