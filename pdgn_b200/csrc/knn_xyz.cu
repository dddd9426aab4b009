// knn_xyz.cu -- batched brute-force k nearest neighbours in xyz (C = 3), bit-exact with the reference.
//
// Replaces knnquery_cuda_kernel (lib/pointops/src/knnquery/knnquery_cuda_kernel.cu:6-50: one thread per query,
// insertion sort in 2400 B of local memory, FP64 compares) and nearestneighbor_cuda_kernel_fast
// (lib/pointops/src/interpolation/interpolation_cuda_kernel.cu:134-176).  Result contract = the reference's:
// ascending (d2, index), d2 by the native FMUL/FFMA/FFMA chain (d2_xyz), NaN/+inf distances never selected,
// missing neighbours reported as idx 0 / dist2 +inf.
//
// Fast path (knn_select_kernel): no sorted structure and no divergent insertion in the hot loop.
//   pass 1  (lane = query) every lane scans all candidates, staged in shared memory as SoA planes and read as
//           warp-broadcast LDS.128 (4 candidates per plane per load), and records only MINIMA: one per subgroup of
//           `ss` consecutive candidates (stored as bf16 rounded DOWN) and one per group of `gsz` subgroups (bf16
//           rounded UP).  6 FMA-pipe + 0.5 FMNMX3 instructions per point pair.
//   bound   tau = k-th smallest group minimum, from a register-resident bitonic network over the G group minima
//           (branch free, lane = query).  k distinct groups each hold a candidate <= tau, so tau bounds the k-th
//           nearest distance from above; rounding up keeps it a bound, rounding the subgroup minima down keeps the
//           filter below free of false negatives.
//   pass 2  (warp = query, 32 queries in turn) the warp ballots which subgroups can hold a survivor
//           (subgroup minimum <= tau: about k of them), rescans only those candidates (32 per step, one per lane),
//           compacts the survivors d2 <= tau in index order with ballot/popc, and ranks them: every lane counts how
//           many survivors precede its own in (d2, index) order and, if that rank is < k, stores straight into
//           idx[rank].  Warp-uniform control flow throughout.
//   A survivor overflow (adversarial ties / clustering / fewer than k finite candidates) sends that one query to
//   an exact serial scan, so the result is always exact.
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
// small-k kernel (k <= 4: the 3-NN of nearestneighbor): register-resident sorted list per query
// =====================================================================================================
// With k this small a record-breaking candidate is rare (k + k ln(n/k) per query), so the classic scan is efficient:
// 6 FMA-pipe instructions + one compare per pair, and a short branch when some lane of the warp improves its list.
constexpr int KK_T = 128;
constexpr int KK_R = 2;      // queries per thread
constexpr int KK_TILE = 2048;

template <int K>
__global__ void __launch_bounds__(KK_T) knn_smallk_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, int n,
                                                         int m, int* __restrict__ idx, float* __restrict__ dist2) {
    __shared__ __align__(16) float tile[3 * KK_TILE];
    const int bz = blockIdx.y, t = threadIdx.x;
    const float* pb = xyz + (size_t)bz * n * 3;
    float qx[KK_R], qy[KK_R], qz[KK_R], bd[KK_R][K];
    int bi[KK_R][K];
#pragma unroll
    for (int r = 0; r < KK_R; ++r) {
        const int q = min(blockIdx.x * (KK_T * KK_R) + r * KK_T + t, m - 1);
        const float* qp = new_xyz + ((size_t)bz * m + q) * 3;
        qx[r] = qp[0]; qy[r] = qp[1]; qz[r] = qp[2];
#pragma unroll
        for (int e = 0; e < K; ++e) { bd[r][e] = kInf; bi[r][e] = 0; }
    }
    for (int j0 = 0; j0 < n; j0 += KK_TILE) {
        const int cnt = min(KK_TILE, n - j0);
        __syncthreads();
        for (int e = t; e < cnt * 3; e += KK_T) {
            const int j = e / 3, c = e - j * 3;
            tile[c * KK_TILE + j] = pb[(size_t)j0 * 3 + e];
        }
        if (t < ((cnt + 3) & ~3) - cnt) {  // NaN padding: never strictly smaller than anything
            const float nanv = __int_as_float(0x7fc00000);
            tile[cnt + t] = nanv; tile[KK_TILE + cnt + t] = nanv; tile[2 * KK_TILE + cnt + t] = nanv;
        }
        __syncthreads();
#pragma unroll 2
        for (int j = 0; j < cnt; j += 4) {
            const float4 X = *reinterpret_cast<const float4*>(tile + j);
            const float4 Y = *reinterpret_cast<const float4*>(tile + KK_TILE + j);
            const float4 Z = *reinterpret_cast<const float4*>(tile + 2 * KK_TILE + j);
#pragma unroll
            for (int r = 0; r < KK_R; ++r) {
                float d[4];
                d[0] = d2_xyz(qx[r], qy[r], qz[r], X.x, Y.x, Z.x);
                d[1] = d2_xyz(qx[r], qy[r], qz[r], X.y, Y.y, Z.y);
                d[2] = d2_xyz(qx[r], qy[r], qz[r], X.z, Y.z, Z.z);
                d[3] = d2_xyz(qx[r], qy[r], qz[r], X.w, Y.w, Z.w);
                if (min3(fminf(d[0], d[1]), d[2], d[3]) < bd[r][K - 1]) {  // rare: some candidate of the quad enters the list
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (d[u] < bd[r][K - 1]) {
                            float dv = d[u];
                            int iv = j0 + j + u;
                            bool shifting = false;  // once inserted, the displaced entries shift down unconditionally
#pragma unroll
                            for (int e = 0; e < K; ++e) {  // strict '<': equal distances keep the lower index first
                                if (shifting || dv < bd[r][e]) {
                                    shifting = true;
                                    const float td = bd[r][e];
                                    const int ti = bi[r][e];
                                    bd[r][e] = dv; bi[r][e] = iv;
                                    dv = td; iv = ti;
                                }
                            }
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < KK_R; ++r) {
        const int q = blockIdx.x * (KK_T * KK_R) + r * KK_T + t;
        if (q < m) {
            const size_t o = ((size_t)bz * m + q) * K;
#pragma unroll
            for (int e = 0; e < K; ++e) {
                idx[o + e] = bi[r][e];
                if (dist2) dist2[o + e] = bd[r][e];
            }
        }
    }
}

template <int K>
static int launch_smallk(const float* xyz, const float* new_xyz, int b, int n, int m, int* idx, float* dist2, cudaStream_t st) {
    dim3 grid((m + KK_T * KK_R - 1) / (KK_T * KK_R), b);
    knn_smallk_kernel<K><<<grid, KK_T, 0, st>>>(xyz, new_xyz, n, m, idx, dist2);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}

// =====================================================================================================
// fast kernel: minima scan + warp-cooperative selection
// =====================================================================================================
constexpr int KS_TILE = 2048;   // candidate capacity of the shared tile
constexpr int KS_NSUB = 128;    // subgroup-minimum rows per query
constexpr int KS_SCAP = 128;    // survivor capacity per query

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
// multiple of 4 with NaN (a NaN distance is never a minimum).
template <int THREADS>
__device__ __forceinline__ void ks_load_tile(float* tile, const float* __restrict__ pb, int j0, int cnt, int t) {
    for (int e = t; e < cnt * 3; e += THREADS) {
        const int j = e / 3, c = e - j * 3;
        tile[c * KS_TILE + j] = pb[(size_t)j0 * 3 + e];
    }
    const int cnt4 = (cnt + 3) & ~3;
    if (t < cnt4 - cnt) {
        const float nanv = __int_as_float(0x7fc00000);
        tile[cnt + t] = nanv;
        tile[KS_TILE + cnt + t] = nanv;
        tile[2 * KS_TILE + cnt + t] = nanv;
    }
}

// per-warp shared block
template <int G>
struct alignas(16) KsWarp {
    unsigned short sub[KS_NSUB][32];   // [subgroup][(query + subgroup) & 31]  bf16, rounded down  (swizzled: both the
                                       //   lane=query writes and the lane=subgroup reads are bank-conflict free)
    unsigned short grp[G][32];         // [group][query]  bf16, rounded up
    unsigned long long key[KS_SCAP];   // survivors of the query being selected: (d2 bits << 32) | candidate index --
                                       //   d2 >= 0, so unsigned order of the key == the reference's (d2, index) order
    unsigned short plist[KS_NSUB];     // subgroups that may hold survivors
    int nsurv;                         // survivor counter of the query being selected
};

template <int G, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32) knn_select_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz,
                                                                int n, int m, int k, int log2ss, int gsz, int* __restrict__ idx,
                                                                float* __restrict__ dist2) {
    constexpr int THREADS = NWARPS * 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tile = reinterpret_cast<float*>(smem_raw);  // [3][KS_TILE]
    KsWarp<G>* wsm = reinterpret_cast<KsWarp<G>*>(tile + 3 * KS_TILE) + (threadIdx.x >> 5);

    const int bz = blockIdx.y, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int ss = 1 << log2ss;
    const int nsub = (n + ss - 1) >> log2ss;          // <= KS_NSUB
    const int ng = (nsub + gsz - 1) / gsz;            // <= G
    const float* pb = xyz + (size_t)bz * n * 3;
    const int q0 = blockIdx.x * THREADS + warp * 32;  // first query of this warp
    const int qmine = min(q0 + lane, m - 1);
    const float* qp = new_xyz + ((size_t)bz * m + qmine) * 3;
    const float qx = qp[0], qy = qp[1], qz = qp[2];

    for (int g = ng; g < G; ++g) wsm->grp[g][lane] = 0x7f80;  // +inf for groups that do not exist

    // ---------------- pass 1: subgroup / group minima (lane = query)
    {
        float gmn = kInf;
        int gleft = gsz;                                   // subgroups left in the current group
        unsigned short* subrow = &wsm->sub[0][0];          // row of the current subgroup
        unsigned short* grow = &wsm->grp[0][lane];
        int col = lane;                                    // (lane + sg) & 31
        for (int j0 = 0; j0 < n; j0 += KS_TILE) {
            const int cnt = min(KS_TILE, n - j0);
            __syncthreads();
            ks_load_tile<THREADS>(tile, pb, j0, cnt, t);
            __syncthreads();
            for (int jb = 0; jb < cnt; jb += ss) {
                const int je = min(cnt, jb + ss);
                float mn = kInf;
#pragma unroll 4
                for (int j = jb; j < je; j += 4) {
                    const float4 X = *reinterpret_cast<const float4*>(tile + j);
                    const float4 Y = *reinterpret_cast<const float4*>(tile + KS_TILE + j);
                    const float4 Z = *reinterpret_cast<const float4*>(tile + 2 * KS_TILE + j);
                    const float d0 = d2_xyz(qx, qy, qz, X.x, Y.x, Z.x), d1 = d2_xyz(qx, qy, qz, X.y, Y.y, Z.y);
                    const float d2 = d2_xyz(qx, qy, qz, X.z, Y.z, Z.z), d3 = d2_xyz(qx, qy, qz, X.w, Y.w, Z.w);
                    mn = min3(min3(mn, d0, d1), d2, d3);
                }
                subrow[col] = (unsigned short)(__float_as_uint(mn) >> 16);   // toward zero = down (mn >= 0)
                subrow += 32;
                col = (col + 1) & 31;
                gmn = fminf(gmn, mn);
                if (--gleft == 0) {
                    *grow = (unsigned short)((__float_as_uint(gmn) + 0xffffu) >> 16);  // up
                    grow += 32;
                    gmn = kInf;
                    gleft = gsz;
                }
            }
        }
        if (gleft != gsz) *grow = (unsigned short)((__float_as_uint(gmn) + 0xffffu) >> 16);  // ragged last group
    }

    // ---------------- bound: tau = k-th smallest group minimum (own column only)
    float tau;
    {
        float v[G];
#pragma unroll
        for (int g = 0; g < G; ++g) v[g] = __uint_as_float((unsigned)wsm->grp[g][lane] << 16);
        bitonic_sort_regs<G>(v);
        // v ascending => v[k-1] = max(v[0..k-1]); written as a predicated max so that the compiler cannot turn it into
        // a dynamically indexed (local-memory) array access
        tau = 0.f;
#pragma unroll
        for (int g = 0; g < G; ++g) tau = (g < k) ? fmaxf(tau, v[g]) : tau;
    }
    __syncwarp();

    // ---------------- pass 2: warp-cooperative selection, one query at a time
    const unsigned lt = (1u << lane) - 1u;
    const int nq = min(32, m - q0);
    const bool resident = n <= KS_TILE;  // the single tile of pass 1 is still in shared memory
    for (int qi = 0; qi < nq; ++qi) {
        const float tq = __shfl_sync(kFull, tau, qi);
        const float ax = __shfl_sync(kFull, qx, qi), ay = __shfl_sync(kFull, qy, qi), az = __shfl_sync(kFull, qz, qi);
        // subgroups whose (rounded-down) minimum is <= tau, in ascending order
        int npass = 0;
#pragma unroll
        for (int i = 0; i < KS_NSUB / 32; ++i) {
            const int sg = lane + 32 * i;
            bool p = false;
            if (sg < nsub) p = __uint_as_float((unsigned)wsm->sub[sg][(qi + sg) & 31] << 16) <= tq;
            const unsigned mask = __ballot_sync(kFull, p);
            if (p) wsm->plist[npass + __popc(mask & lt)] = (unsigned short)sg;
            npass += __popc(mask);
        }
        __syncwarp();
        // rescan those subgroups: 4 consecutive candidates per lane, 128 per step.  Survivors are appended in ANY order
        // (shared-memory counter): their 64-bit (d2, index) keys are unique, so the ranking below does not need them sorted.
        const int total = npass << log2ss;
        if (lane == 0) wsm->nsurv = 0;
        __syncwarp();
        for (int p0 = 0; p0 < total; p0 += 128) {
            const int p = p0 + 4 * lane;
            const bool ok = p < total;
            const int sg = wsm->plist[ok ? (p >> log2ss) : 0];
            const int j = (sg << log2ss) + (p & (ss - 1));  // multiple of 4; j..j+3 stay inside the subgroup
            float d[4];
            if (resident) {
                const float4 X = *reinterpret_cast<const float4*>(tile + j);
                const float4 Y = *reinterpret_cast<const float4*>(tile + KS_TILE + j);
                const float4 Z = *reinterpret_cast<const float4*>(tile + 2 * KS_TILE + j);
                d[0] = d2_xyz(ax, ay, az, X.x, Y.x, Z.x);
                d[1] = d2_xyz(ax, ay, az, X.y, Y.y, Z.y);
                d[2] = d2_xyz(ax, ay, az, X.z, Y.z, Z.z);
                d[3] = d2_xyz(ax, ay, az, X.w, Y.w, Z.w);
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    d[u] = kInf;
                    if (ok && j + u < n) {
                        const float* c = pb + (size_t)(j + u) * 3;
                        d[u] = d2_xyz(ax, ay, az, __ldg(c), __ldg(c + 1), __ldg(c + 2));
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                // NaN and +inf never survive; j+u >= n: ragged last subgroup
                if (ok && j + u < n && d[u] <= tq && d[u] < kInf) {
                    const int pos = atomicAdd(&wsm->nsurv, 1);
                    if (pos < KS_SCAP) wsm->key[pos] = ((unsigned long long)__float_as_uint(d[u]) << 32) | (unsigned)(j + u);
                }
            }
        }
        __syncwarp();
        const int nsurv = wsm->nsurv;
        int* oi = idx + ((size_t)bz * m + q0 + qi) * k;
        float* od = dist2 ? dist2 + ((size_t)bz * m + q0 + qi) * k : nullptr;
        if (nsurv <= KS_SCAP) {
            // rank = number of survivors whose (d2, index) key is smaller
            for (int e0 = 0; e0 < nsurv; e0 += 32) {
                const int e = e0 + lane;
                const bool have = e < nsurv;
                const unsigned long long ke = have ? wsm->key[e] : ~0ull;
                int rank = 0;
                int f = 0;
                for (; f + 2 <= nsurv; f += 2) {
                    const ulonglong2 kf = *reinterpret_cast<const ulonglong2*>(wsm->key + f);
                    rank += (kf.x < ke) ? 1 : 0;
                    rank += (kf.y < ke) ? 1 : 0;
                }
                if (f < nsurv) rank += (wsm->key[f] < ke) ? 1 : 0;
                if (have && rank < k) {
                    oi[rank] = (int)(unsigned)ke;
                    if (od) od[rank] = __uint_as_float((unsigned)(ke >> 32));
                }
            }
            for (int e = nsurv + lane; e < k; e += 32) {  // fewer than k finite candidates
                oi[e] = 0;
                if (od) od[e] = kInf;
            }
        } else if (lane == 0) {
            // survivor overflow: exact serial scan of every candidate for this query (the key array doubles as the list)
            float* ldl = reinterpret_cast<float*>(wsm->key);
            int* lil = reinterpret_cast<int*>(wsm->key) + KS_SCAP;
            for (int e = 0; e < k; ++e) {
                ldl[e] = kInf;
                lil[e] = 0;
            }
            float thr = kInf;
            for (int j = 0; j < n; ++j) {
                const float* c = pb + (size_t)j * 3;
                const float d = d2_xyz(ax, ay, az, __ldg(c), __ldg(c + 1), __ldg(c + 2));
                if (d < thr) {
                    list_insert(ldl, lil, 1, 0, k, d, j);
                    thr = ldl[k - 1];
                }
            }
            for (int e = 0; e < k; ++e) {
                oi[e] = lil[e];
                if (od) od[e] = ldl[e];
            }
        }
        __syncwarp();
    }
}

template <int G, int NWARPS>
static int launch_select(const float* xyz, const float* new_xyz, int b, int n, int m, int k, int log2ss, int gsz, int* idx,
                         float* dist2, cudaStream_t st) {
    const size_t smem = (size_t)3 * KS_TILE * 4 + (size_t)NWARPS * sizeof(KsWarp<G>);
    PDGN_CUDA(cudaFuncSetAttribute(knn_select_kernel<G, NWARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((m + NWARPS * 32 - 1) / (NWARPS * 32), b);
    knn_select_kernel<G, NWARPS><<<grid, NWARPS * 32, smem, st>>>(xyz, new_xyz, n, m, k, log2ss, gsz, idx, dist2);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}

// Subgroup size (power of two >= 4 covering n with at most KS_NSUB subgroups) and subgroups per group such that the
// number of groups is in [k, G].  Returns false when no such split exists (tiny n or large k): generic kernel.
static bool select_plan(int n, int k, int G, int* log2ss, int* gsz) {
    int l = 2;
    while (((n + (1 << l) - 1) >> l) > KS_NSUB) ++l;
    if ((1 << l) > KS_TILE) return false;
    const int nsub = (n + (1 << l) - 1) >> l;
    for (int g = 4; g >= 1; g >>= 1) {
        const int ng = (nsub + g - 1) / g;
        if (ng >= k && ng <= G) {
            *log2ss = l;
            *gsz = g;
            return true;
        }
    }
    return false;
}

static int knn_dispatch(const float* xyz, const float* new_xyz, int b, int n, int m, int k, int* idx, float* dist2, cudaStream_t st) {
    int log2ss = 0, gsz = 0;
    switch (k) {  // tiny k: register-resident lists
        case 1: return launch_smallk<1>(xyz, new_xyz, b, n, m, idx, dist2, st);
        case 2: return launch_smallk<2>(xyz, new_xyz, b, n, m, idx, dist2, st);
        case 3: return launch_smallk<3>(xyz, new_xyz, b, n, m, idx, dist2, st);
        case 4: return launch_smallk<4>(xyz, new_xyz, b, n, m, idx, dist2, st);
        default: break;
    }
    // k <= 24 of 32 groups / k <= 48 of 64 groups keeps the expected survivor count (~G/(G-k) * k-ish) well under KS_SCAP
    if (k <= 24 && select_plan(n, k, 32, &log2ss, &gsz)) {
        // warps per CTA: as many as possible (they share the candidate tile) while the grid still covers the SMs --
        // the training shapes have only 256..1024 queries per batch element
        const long long want = 132;  // ~0.9 x SMs: one full wave of the widest CTA beats two waves of narrower ones
        if ((long long)((m + 511) / 512) * b >= want) return launch_select<32, 16>(xyz, new_xyz, b, n, m, k, log2ss, gsz, idx, dist2, st);
        if ((long long)((m + 255) / 256) * b >= want) return launch_select<32, 8>(xyz, new_xyz, b, n, m, k, log2ss, gsz, idx, dist2, st);
        if ((long long)((m + 127) / 128) * b >= want) return launch_select<32, 4>(xyz, new_xyz, b, n, m, k, log2ss, gsz, idx, dist2, st);
        return launch_select<32, 2>(xyz, new_xyz, b, n, m, k, log2ss, gsz, idx, dist2, st);
    }
    if (k <= 48 && select_plan(n, k, 64, &log2ss, &gsz)) return launch_select<64, 8>(xyz, new_xyz, b, n, m, k, log2ss, gsz, idx, dist2, st);
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
    if (b < 0 || n < 0 || m < 0 || k < 1) return PDGN_ERR_BAD_ARG;
    if (k > 128 || b > 65535) return PDGN_ERR_UNSUPPORTED;
    if (b == 0 || m == 0) return PDGN_OK;  // empty query set: nothing to write (pointers may be null)
    if (!new_xyz || !idx || (!xyz && n > 0)) return PDGN_ERR_BAD_ARG;
    // n == 0 falls through to the generic kernel, which leaves the reference's initial values (idx 0, dist +inf)
    return knn_dispatch(xyz, new_xyz, b, n, m, k, idx, dist2, (cudaStream_t)stream);
}

extern "C" int pdgn_nn3(const float* unknown, const float* known, int b, int n, int m, float* dist2, int* idx, void* stream) {
    if (b < 0 || n < 0 || m < 0) return PDGN_ERR_BAD_ARG;
    if (b > 65535) return PDGN_ERR_UNSUPPORTED;
    if (b == 0 || n == 0) return PDGN_OK;
    if (!unknown || !dist2 || !idx || (!known && m > 0)) return PDGN_ERR_BAD_ARG;
    return knn_dispatch(known, unknown, b, m, n, 3, idx, dist2, (cudaStream_t)stream);
}
