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
//   a warp-cooperative exact selection (k rounds of a min-search over all candidates), so the result is always exact.
// Generic path (knn_generic_kernel): any k <= 200, any n; threshold-guarded insertion into a shared-memory list.
#include <cstdlib>
#include "common.cuh"
#include "multi.cuh"

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
// multiple of 16 with NaN (a NaN distance is never a minimum).
template <int THREADS>
__device__ __forceinline__ void ks_load_tile(float* tile, const float* __restrict__ pb, int j0, int cnt, int t) {
    for (int e = t; e < cnt * 3; e += THREADS) {
        const int j = e / 3, c = e - j * 3;
        tile[c * KS_TILE + j] = pb[(size_t)j0 * 3 + e];
    }
    const int cnt4 = (cnt + 15) & ~15;  // pass 1 walks blocks of 16
    if (t < cnt4 - cnt) {
        const float nanv = __int_as_float(0x7fc00000);
        tile[cnt + t] = nanv;
        tile[KS_TILE + cnt + t] = nanv;
        tile[2 * KS_TILE + cnt + t] = nanv;
    }
}

// per-warp shared block
struct KsSelect {                      // scratch of the query being selected (pass 2)
    unsigned long long key[KS_SCAP];   // survivors: (d2 bits << 32) | candidate index -- d2 >= 0, so unsigned order of the
                                       //   key == the reference's (d2, index) order
    unsigned short plist[KS_NSUB];     // subgroups that may hold survivors
};
template <int G, int S>
struct alignas(16) KsWarp {
    unsigned short sub[KS_NSUB][32];   // [subgroup][(query + subgroup) & 31]  bf16, rounded down  (swizzled: both the
                                       //   lane=query writes and the lane=subgroup reads are bank-conflict free)
    union {                            // the group minima are dead once every lane holds its bound tau in a register
        unsigned short grp[G][32];     // [group][query]  bf16, rounded up (pass 1 -> bound)
        KsSelect sel[S];               // pass 2: one scratch per slice (warp) of the query block
    };
};

// S > 1 ("slices"): S warps share one block of 32 queries -- in pass 1 each scans 1/S of the candidates (whole groups), in pass 2
// each selects 32/S of the queries.  Used when the batch has too few queries to fill the chip with one warp per block (the
// training shapes: 256..1024 queries per cloud); needs the single resident tile (n <= KS_TILE) and subgroups of <= 16.
template <int G, int NWARPS, int S>
__device__ __forceinline__ void knn_select_body(const float* __restrict__ xyz, const float* __restrict__ new_xyz, int n, int m, int k,
                                                int log2ss, int gsz, int* __restrict__ idx, float* __restrict__ dist2, int bx, int bz) {
    constexpr int THREADS = NWARPS * 32;
    constexpr int NQ = NWARPS / S;  // query blocks per CTA
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tile = reinterpret_cast<float*>(smem_raw);  // [3][KS_TILE]
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int qblock = warp / S, slice = warp % S;
    KsWarp<G, S>* wsm = reinterpret_cast<KsWarp<G, S>*>(tile + 3 * KS_TILE) + qblock;
    KsSelect& sel = wsm->sel[slice];

    const int ss = 1 << log2ss;
    const int nsub = (n + ss - 1) >> log2ss;          // <= KS_NSUB
    const int ng = (nsub + gsz - 1) / gsz;            // <= G
    const float* pb = xyz + (size_t)bz * n * 3;
    const int q0 = (bx * NQ + qblock) * 32;   // first query of this block
    const int qmine = max(0, min(q0 + lane, m - 1));
    const float* qp = new_xyz + ((size_t)bz * m + qmine) * 3;
    const float qx = qp[0], qy = qp[1], qz = qp[2];

    for (int g = ng; g < G; ++g) wsm->grp[g][lane] = 0x7f80;  // +inf for groups that do not exist
    for (int sg = nsub; sg < KS_NSUB; ++sg) wsm->sub[sg][lane] = 0xffff;  // subgroups that do not exist never pass the bound

    // ---------------- pass 1: subgroup / group minima (lane = query)
    // Candidates are walked in blocks of 16 (4 quads, fully unrolled, each quad loaded one quad ahead); a block yields its
    // four quad minima, which are then emitted at the subgroup granularity (ss = 4, 8, 16, or a multiple of 16).
    {
        // slice `slice` owns the candidate blocks [jb0, jb1) of the (single) tile: whole groups of subgroups
        int jb0 = 0, jb1 = KS_TILE;
        if (S > 1) {
            const int unit = max(16, ss * gsz);
            const int units = (((min(KS_TILE, n) + 15) & ~15) + unit - 1) / unit;
            const int per = (units + S - 1) / S;
            jb0 = slice * per * unit;
            jb1 = jb0 + per * unit;
        }
        const int sg0 = jb0 >> log2ss;                     // first subgroup of the slice (a multiple of gsz)
        float gmn = kInf;
        int gleft = gsz;                                   // subgroups left in the current group
        unsigned short* subrow = &wsm->sub[min(sg0, KS_NSUB - 1)][0];           // row of the current subgroup
        unsigned short* grow = &wsm->grp[min(sg0 / gsz, G - 1)][lane];
        int col = (lane + sg0) & 31;                       // (lane + sg) & 31
        int sgi = sg0;                                     // index of the next subgroup to emit
        auto emit = [&](float mn) {
            if (sgi < nsub) {                              // (the NaN padding of the last block can start an extra one)
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
                ++sgi;
            }
        };
        float acc = kInf;                                  // ss > 16: minimum of the blocks of the current subgroup
        for (int j0 = 0; j0 < n; j0 += KS_TILE) {
            const int cnt = min(KS_TILE, n - j0);
            __syncthreads();
            ks_load_tile<THREADS>(tile, pb, j0, cnt, t);
            __syncthreads();
            const int cnt16 = (cnt + 15) & ~15;
            const int jbeg = min(jb0, cnt16 - 16), jend = min(jb1, cnt16);   // (an empty slice reads a valid quad, then loops 0 times)
            float4 X = *reinterpret_cast<const float4*>(tile + jbeg);
            float4 Y = *reinterpret_cast<const float4*>(tile + KS_TILE + jbeg);
            float4 Z = *reinterpret_cast<const float4*>(tile + 2 * KS_TILE + jbeg);
            for (int jb = jb0; jb < jend; jb += 16) {
                float mq[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int jn = q < 3 ? jb + 4 * q + 4 : min(jb + 16, cnt16 - 4);
                    const float4 Xn = *reinterpret_cast<const float4*>(tile + jn);
                    const float4 Yn = *reinterpret_cast<const float4*>(tile + KS_TILE + jn);
                    const float4 Zn = *reinterpret_cast<const float4*>(tile + 2 * KS_TILE + jn);
                    const float d0 = d2_xyz(qx, qy, qz, X.x, Y.x, Z.x), d1 = d2_xyz(qx, qy, qz, X.y, Y.y, Z.y);
                    const float d2 = d2_xyz(qx, qy, qz, X.z, Y.z, Z.z), d3 = d2_xyz(qx, qy, qz, X.w, Y.w, Z.w);
                    mq[q] = fminf(min3(d0, d1, d2), d3);
                    X = Xn; Y = Yn; Z = Zn;
                }
                if (log2ss == 4) {
                    emit(fminf(min3(mq[0], mq[1], mq[2]), mq[3]));
                } else if (log2ss == 3) {
                    emit(fminf(mq[0], mq[1]));
                    emit(fminf(mq[2], mq[3]));
                } else if (log2ss == 2) {
                    emit(mq[0]);
                    emit(mq[1]);
                    emit(mq[2]);
                    emit(mq[3]);
                } else {
                    acc = fminf(acc, fminf(min3(mq[0], mq[1], mq[2]), mq[3]));
                    if (((j0 + jb + 16) & (ss - 1)) == 0) {
                        emit(acc);
                        acc = kInf;
                    }
                }
            }
        }
        if (log2ss > 4 && (n & (ss - 1)) != 0 && ((((n + 15) & ~15)) & (ss - 1)) != 0) emit(acc);  // ragged last subgroup
        if (gleft != gsz) *grow = (unsigned short)((__float_as_uint(gmn) + 0xffffu) >> 16);  // ragged last group
    }

    if (S > 1) __syncthreads();  // the minima of all slices are in shared memory
    // ---------------- bound: tau = k-th smallest group minimum (own column only; every slice computes all 32 bounds)
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
    if (S > 1) __syncthreads();  // every slice has read the group minima: their storage becomes the selection scratch
    else __syncwarp();

    // ---------------- pass 2: warp-cooperative selection, one query at a time
    const unsigned lt = (1u << lane) - 1u;
    const int nq = min(32, m - q0);
    const bool resident = n <= KS_TILE;  // the single tile of pass 1 is still in shared memory
    for (int qi = slice; qi < nq; qi += S) {
        // NaN and +inf never survive: the bound is clamped to the largest finite float
        const float tq = fminf(__shfl_sync(kFull, tau, qi), 3.402823466e+38f);
        const float ax = __shfl_sync(kFull, qx, qi), ay = __shfl_sync(kFull, qy, qi), az = __shfl_sync(kFull, qz, qi);
        // subgroups whose (rounded-down) minimum is <= tau, in ascending order.  bf16 bit patterns of non-negative floats
        // order like unsigned integers, and (v << 16) <= T  <=>  v <= (T >> 16); rows >= nsub hold 0xffff (never pass)
        const unsigned tq16 = __float_as_uint(tq) >> 16;
        const unsigned short* sp = &wsm->sub[lane][(qi + lane) & 31];  // sub[lane + 32 i][(qi + lane + 32 i) & 31]
        int npass = 0;
#pragma unroll
        for (int i = 0; i < KS_NSUB / 32; ++i) {
            const bool p = (unsigned)sp[i * 32 * 32] <= tq16;
            const unsigned mask = __ballot_sync(kFull, p);
            if (p) sel.plist[npass + __popc(mask & lt)] = (unsigned short)(lane + 32 * i);
            npass += __popc(mask);
        }
        __syncwarp();
        // rescan those subgroups: 4 consecutive candidates per lane, 128 per step; survivors (d2 <= tau) are compacted
        // with ballots into the key array (their 64-bit (d2, index) keys are unique, the ranking below orders them)
        const int total = npass << log2ss;
        int nsurv = 0;
        for (int p0 = 0; p0 < total; p0 += 128) {
            const int p = p0 + 4 * lane;
            const bool ok = p < total;
            const int sg = sel.plist[ok ? (p >> log2ss) : 0];
            const int j = (sg << log2ss) + (p & (ss - 1));  // multiple of 4; j..j+3 stay inside the subgroup
            const float tql = ok ? tq : -1.0f;              // lanes past the end of the list keep nothing
            float d[4];
            if (resident) {  // the tile is NaN-padded past n: a ragged last subgroup needs no index check
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
                const bool keep = d[u] <= tql;  // false for NaN
                const unsigned mk = __ballot_sync(kFull, keep);
                const int pos = nsurv + __popc(mk & lt);
                if (keep && pos < KS_SCAP) sel.key[pos] = ((unsigned long long)__float_as_uint(d[u]) << 32) | (unsigned)(j + u);
                nsurv += __popc(mk);
            }
        }
        __syncwarp();
        int* oi = idx + ((size_t)bz * m + q0 + qi) * k;
        float* od = dist2 ? dist2 + ((size_t)bz * m + q0 + qi) * k : nullptr;
        if (nsurv <= KS_SCAP) {
            // rank = number of survivors whose (d2, index) key is smaller
            for (int e0 = 0; e0 < nsurv; e0 += 32) {
                const int e = e0 + lane;
                const bool have = e < nsurv;
                const unsigned long long ke = have ? sel.key[e] : ~0ull;
                int rank = 0;
                int f = 0;
                for (; f + 2 <= nsurv; f += 2) {
                    const ulonglong2 kf = *reinterpret_cast<const ulonglong2*>(sel.key + f);
                    rank += (kf.x < ke) ? 1 : 0;
                    rank += (kf.y < ke) ? 1 : 0;
                }
                if (f < nsurv) rank += (sel.key[f] < ke) ? 1 : 0;
                if (have && rank < k) {
                    oi[rank] = (int)(unsigned)ke;
                    if (od) od[rank] = __uint_as_float((unsigned)(ke >> 32));
                }
            }
            for (int e = nsurv + lane; e < k; e += 32) {  // fewer than k finite candidates
                oi[e] = 0;
                if (od) od[e] = kInf;
            }
        } else {
            // survivor overflow (duplicates, clusters, fewer than k finite candidates): warp-cooperative exact selection.  k rounds of
            // "smallest (d2, index) key above the previous winner"; every lane scans n/32 candidates per round, one min-reduction
            // per round.  Exact by construction, ~k*n/32 distance evaluations per lane instead of n serial ones on lane 0.
            unsigned long long last = 0;
            for (int e = 0; e < k; ++e) {
                unsigned long long best = ~0ull;
                for (int j = lane; j < n; j += 32) {
                    const float* c = pb + (size_t)j * 3;
                    const float d = d2_xyz(ax, ay, az, __ldg(c), __ldg(c + 1), __ldg(c + 2));
                    const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)j;
                    if (d <= 3.402823466e+38f && (e == 0 || key > last) && key < best) best = key;   // NaN / +inf never selected
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long other = __shfl_xor_sync(kFull, best, o);
                    best = other < best ? other : best;
                }
                if (best == ~0ull) {
                    for (int f = e + lane; f < k; f += 32) {
                        oi[f] = 0;
                        if (od) od[f] = kInf;
                    }
                    break;
                }
                if (lane == 0) {
                    oi[e] = (int)(unsigned)best;
                    if (od) od[e] = __uint_as_float((unsigned)(best >> 32));
                }
                last = best;
            }
        }
        __syncwarp();
    }
}

template <int G, int NWARPS, int S>
__global__ void __launch_bounds__(NWARPS * 32) knn_select_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz,
                                                                int n, int m, int k, int log2ss, int gsz, int* __restrict__ idx,
                                                                float* __restrict__ dist2) {
    knn_select_body<G, NWARPS, S>(xyz, new_xyz, n, m, k, log2ss, gsz, idx, dist2, blockIdx.x, blockIdx.y);
}

// problem-descriptor launch (multi.cuh): CTA -> (problem, query block)
template <int G, int NWARPS, int S>
__global__ void __launch_bounds__(NWARPS * 32) knn_select_multi_kernel(const __grid_constant__ KnnTable tb, int k) {
    const int pi = multi_find(tb, blockIdx.x);
    const KnnProb& pr = tb.p[pi];
    knn_select_body<G, NWARPS, S>(pr.xyz, pr.q, pr.n, pr.m, k, pr.log2ss, pr.gsz, pr.idx, nullptr, blockIdx.x - pr.cta0, blockIdx.y);
}

template <int G, int NWARPS, int S = 1>
static int launch_select(const float* xyz, const float* new_xyz, int b, int n, int m, int k, int log2ss, int gsz, int* idx,
                         float* dist2, cudaStream_t st) {
    constexpr int NQ = NWARPS / S;
    const size_t smem = (size_t)3 * KS_TILE * 4 + (size_t)NQ * sizeof(KsWarp<G, S>);
    PDGN_CUDA(cudaFuncSetAttribute(knn_select_kernel<G, NWARPS, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((m + NQ * 32 - 1) / (NQ * 32), b);
    knn_select_kernel<G, NWARPS, S><<<grid, NWARPS * 32, smem, st>>>(xyz, new_xyz, n, m, k, log2ss, gsz, idx, dist2);
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
    for (int g = 1; g <= 4; g <<= 1) {  // most groups first: the k-th smallest of more group minima is a tighter bound
        const int ng = (nsub + g - 1) / g;
        if (ng >= k && ng <= G) {
            *log2ss = l;
            *gsz = g;
            return true;
        }
    }
    return false;
}

// knn_gram.cu: the full-size-cloud kernel (register-blocked Gram-form filter + lane-private exact selection)
bool knn_gram_eligible(int n, int k);
int knn_gram_launch(const float* xyz, const float* new_xyz, int b, int n, int m, int k, int* idx, float* dist2, cudaStream_t st);

static int knn_dispatch(const float* xyz, const float* new_xyz, int b, int n, int m, int k, int* idx, float* dist2, cudaStream_t st) {
    int log2ss = 0, gsz = 0;
    // Full-size clouds with enough queries to give every SM a 512-query CTA: the round-2 kernel.  Fewer queries (the training
    // shapes) keep the sliced select kernel below, whose CTAs shrink until the grid covers the chip.
    {
        static const char* impl = tune_env("PDGN_KNN_IMPL");   // tuning hook: "gram" / "select" force one kernel family
        // One 512-query CTA per SM and ~85 us per CTA whatever the query count: worth it when the last wave is >= 70 % full
        // (measured cross-over against the select kernel: ~100 CTAs; profiles/r02_knn_gram_vs_select.txt)
        const long long ctas = (long long)b * ((m + 511) / 512);
        const long long waves = (ctas + 147) / 148;
        const bool want = impl ? impl[0] == 'g' : (ctas * 10 >= waves * 148 * 7);
        if (want && knn_gram_eligible(n, k)) return knn_gram_launch(xyz, new_xyz, b, n, m, k, idx, dist2, st);
    }
    // tiny k with few candidates: register-resident lists (every lane inserts ~k ln(n) times, which keeps the whole warp in
    // the divergent insertion path once n is in the hundreds: 22 instr/pair at n = 1024, so larger n goes to the select kernel)
    static const bool force_smallk = tune_env("PDGN_KNN_SMALLK") != nullptr;  // tuning hook
    if (k <= 4 && (n < 256 || force_smallk || !select_plan(n, k, 32, &log2ss, &gsz))) {
        switch (k) {
            case 1: return launch_smallk<1>(xyz, new_xyz, b, n, m, idx, dist2, st);
            case 2: return launch_smallk<2>(xyz, new_xyz, b, n, m, idx, dist2, st);
            case 3: return launch_smallk<3>(xyz, new_xyz, b, n, m, idx, dist2, st);
            default: return launch_smallk<4>(xyz, new_xyz, b, n, m, idx, dist2, st);
        }
    }
    // Expected survivors of the bound "k-th smallest of G group minima": -ln(1 - k/G) * G, i.e. 31 for k = 20 of 32 groups
    // but 24 of 64 groups (one ranking round instead of two, fewer subgroups to rescan); below k ~ 12 the two are equal
    // and 32 groups cost less to sort.  k <= 24 of 32 / k <= 48 of 64 keeps the count well under KS_SCAP.
    // Warps per CTA: as many as possible (they share the candidate tile) while the grid still covers the SMs -- the
    // training shapes have only 256..1024 queries per batch element.
    const long long want = 132;  // ~0.9 x SMs: one full wave of the widest CTA beats two waves of narrower ones
    const int width = (long long)((m + 511) / 512) * b >= want ? 16 : (long long)((m + 255) / 256) * b >= want ? 8
                    : (long long)((m + 127) / 128) * b >= want ? 4 : 2;
    // fewer than ~132 CTAs even at 8 / 4 / 2 query blocks per CTA: S = 2 / 4 / 4 slices per block instead of narrower CTAs
    static const bool no_slices = tune_env("PDGN_KNN_NOSLICES") != nullptr;  // tuning hook
#define PDGN_KS_LAUNCH(G_)                                                                                        \
    do {                                                                                                          \
        const bool sliced = !no_slices && n <= KS_TILE && log2ss <= 4;                                            \
        if (width == 16) return launch_select<G_, 16>(xyz, new_xyz, b, n, m, k, log2ss, gsz, idx, dist2, st);     \
        if (sliced) {                                                                                             \
            if (width == 8) return launch_select<G_, 16, 2>(xyz, new_xyz, b, n, m, k, log2ss, gsz, idx, dist2, st); \
            if (width == 4) return launch_select<G_, 16, 4>(xyz, new_xyz, b, n, m, k, log2ss, gsz, idx, dist2, st); \
            return launch_select<G_, 8, 4>(xyz, new_xyz, b, n, m, k, log2ss, gsz, idx, dist2, st);                \
        }                                                                                                         \
        if (width == 8) return launch_select<G_, 8>(xyz, new_xyz, b, n, m, k, log2ss, gsz, idx, dist2, st);       \
        if (width == 4) return launch_select<G_, 4>(xyz, new_xyz, b, n, m, k, log2ss, gsz, idx, dist2, st);       \
        return launch_select<G_, 2>(xyz, new_xyz, b, n, m, k, log2ss, gsz, idx, dist2, st);                       \
    } while (0)
    static const bool wide_groups = tune_env("PDGN_KNN_G32") == nullptr;  // tuning hook: PDGN_KNN_G32=1 keeps 32 groups
    if (k > 12 && k <= 24 && wide_groups && select_plan(n, k, 64, &log2ss, &gsz) && (n + (1 << log2ss) - 1) >> log2ss > 32 * gsz)
        PDGN_KS_LAUNCH(64);
    if (k <= 24 && select_plan(n, k, 32, &log2ss, &gsz)) PDGN_KS_LAUNCH(32);
#undef PDGN_KS_LAUNCH
    if (k <= 48 && select_plan(n, k, 64, &log2ss, &gsz)) return launch_select<64, 8>(xyz, new_xyz, b, n, m, k, log2ss, gsz, idx, dist2, st);
    const size_t smem = (size_t)3 * KG_TILE * 4 + (size_t)k * KG_T * 8;
    PDGN_CUDA(cudaFuncSetAttribute(knn_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((m + KG_T - 1) / KG_T, b);
    knn_generic_kernel<<<grid, KG_T, smem, st>>>(xyz, new_xyz, n, m, k, idx, dist2);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}

template <int S>
static int knn_multi_launch_s(KnnTable& tb, int b, int k, cudaStream_t st) {
    // one instantiation for every problem: 64 groups, 16 warps = 16/S query blocks x S candidate slices
    constexpr int G = 64, NWARPS = 16, NQ = NWARPS / S;
    int ctas = 0;
    for (int i = 0; i < tb.count; ++i) {
        KnnProb& pr = tb.p[i];
        if (pr.n < 256 || pr.n > KS_TILE || pr.m < 1) return PDGN_ERR_UNSUPPORTED;
        if (!select_plan(pr.n, k, G, &pr.log2ss, &pr.gsz) || pr.log2ss > 4) return PDGN_ERR_UNSUPPORTED;
        pr.cta0 = ctas;
        ctas += (pr.m + NQ * 32 - 1) / (NQ * 32);
    }
    const size_t smem = (size_t)3 * KS_TILE * 4 + (size_t)NQ * sizeof(KsWarp<G, S>);
    PDGN_CUDA(cudaFuncSetAttribute(knn_select_multi_kernel<G, NWARPS, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_select_multi_kernel<G, NWARPS, S><<<dim3(ctas, b), NWARPS * 32, smem, st>>>(tb, k);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}

int knn_multi_launch(KnnTable& tb, int b, int k, cudaStream_t st) {
    if (k < 1 || k > 24 || tb.count < 1 || tb.count > 12) return PDGN_ERR_UNSUPPORTED;
    // slices per query block: with all problems of a step in one grid the chip is full without slicing small query sets
    // thinly; fewer slices = fewer copies of each candidate tile and fewer partial minima to merge
    long long queries = 0;
    for (int i = 0; i < tb.count; ++i) queries += tb.p[i].m;
    static const char* force = tune_env("PDGN_KNN_MULTI_S");
    const int s = force ? atoi(force) : (queries * b >= 148LL * 256 ? 2 : 4);   // measured at B=35: S=1 273 us, S=2 237 us, S=4 257 us
    if (s == 1) return knn_multi_launch_s<1>(tb, b, k, st);
    if (s == 2) return knn_multi_launch_s<2>(tb, b, k, st);
    return knn_multi_launch_s<4>(tb, b, k, st);
}

}  // namespace pdgn

using namespace pdgn;

extern "C" int pdgn_knn_xyz(const float* xyz, const float* new_xyz, int b, int n, int m, int k, int* idx, float* dist2, void* stream) {
    PDGN_RANGE("pdgn_knn_xyz");
    if (b < 0 || n < 0 || m < 0 || k < 1) return PDGN_ERR_BAD_ARG;
    if (k > 200 || b > 65535) return PDGN_ERR_UNSUPPORTED;   // the reference keeps best_dist[200] (knnquery_cuda_kernel.cu:21-22)
    if (b == 0 || m == 0) return PDGN_OK;  // empty query set: nothing to write (pointers may be null)
    if (!new_xyz || !idx || (!xyz && n > 0)) return PDGN_ERR_BAD_ARG;
    // n == 0 falls through to the generic kernel, which leaves the reference's initial values (idx 0, dist +inf)
    return knn_dispatch(xyz, new_xyz, b, n, m, k, idx, dist2, (cudaStream_t)stream);
}

extern "C" int pdgn_nn3(const float* unknown, const float* known, int b, int n, int m, float* dist2, int* idx, void* stream) {
    PDGN_RANGE("pdgn_nn3");
    if (b < 0 || n < 0 || m < 0) return PDGN_ERR_BAD_ARG;
    if (b > 65535) return PDGN_ERR_UNSUPPORTED;
    if (b == 0 || n == 0) return PDGN_OK;
    if (!unknown || !dist2 || !idx || (!known && m > 0)) return PDGN_ERR_BAD_ARG;
    return knn_dispatch(known, unknown, b, m, n, 3, idx, dist2, (cudaStream_t)stream);
}
