// knn_feat_tc.cu -- feature-space kNN graph of the generator on the 5th-generation tensor cores (tcgen05 + TMEM), with an exact
// FP32 re-rank, so that the indices are still the oracle's (north_star: "feature-space kNN (C >= 64) may use tcgen05 for the
// Gram tile only if the top-k candidates are re-ranked in exact FP32").
//
// Replaces, in get_edge_features / get_edge_features_xyz (models/PDGNet_v2.py:449-459, :492-502), the [B,N,N] cuBLAS Gram
// matrix + full torch.sort of every row + slice, and csrc/knn_feat.cu's FP32 SIMT kernel (2 FMA-pipe instructions per pair
// and channel) for the shapes below.  Same contract as knn_feat.cu: d2(i,j) = ONE fma chain over c = 0..C-1 of
// (x[c,i] - x[c,j])^2, total order (d2, index), ranks skip..skip+k-1; bit-exact against oracle_knn_feat.
//
//   prep    (2 small kernels)  channel means; centred copy xc = tf32_rn(x - mean) [B,C,N] (distances are shift invariant, the
//           norms shrink to the spread of the features, and values that ARE tf32 make the tensor core's operand truncation a
//           no-op); squared norms of the centred rows; xT = x transposed [B,N,C] for the re-rank.
//   filter  (knn_feat_tc_kernel)  CTA = 128 queries (= the 128 TMEM lanes) x all candidates, 128 at a time.  Operands are
//           staged by 16-byte cp.async into the canonical MN-major, no-swizzle UMMA layout (core matrix = 8 channels x 4
//           points); one thread issues tcgen05.mma.kind::tf32 (M = 128, N = 128, K = 8) for every 8 channels, the Gram tile
//           accumulates in TMEM and the 128 threads read their own row back with tcgen05.ld.  Thread = query:
//             pass A  h = |xi|^2 + |xj|^2 - 2 G as running minima of 128 STRIDED subgroups (candidate j -> subgroup j mod 128)
//                     in registers; bound = k'-th smallest of the 64 group minima (selnet.cuh, the xyz kernel's network);
//             pass B  the Gram tiles are computed AGAIN (the tensor-core time is small next to the operand traffic) and every
//                     candidate with h <= F is appended to the query's list (<= 32 entries).
//           The margins are rigorous (kf_bounds): U bounds the k'-th smallest REFERENCE distance from above and every
//           candidate with d_ref <= U has h <= F.  Nothing depends on the tensor core's internal summation order beyond a
//           generous absolute error term.
//   rank    (knn_feat_rerank_kernel)  warp = query, lane = candidate: the exact reference chain from xT (16-byte loads), rank
//           by counting over the (d2, index) keys, ranks skip..skip+k-1 stored.
//   Queries whose list overflowed, or that saw fewer than k' finite group minima (duplicates, clusters, NaN), are flagged and
//   recomputed by knn_feat.cu's exact kernel, which skips every CTA without a flagged query.
#include "common.cuh"
#include "selnet.cuh"

namespace pdgn {

constexpr int TF_M = 128;    // queries per CTA = TMEM lanes
constexpr int TF_N = 128;    // candidates per accumulator tile = TMEM columns = subgroups
constexpr int TF_KB = 32;    // channels per staged K-block (4 MMAs of K = 8)
constexpr int TF_CAP = 32;   // list entries per query
constexpr int TF_KMAX = 20;  // k + skip the bound network is tuned for (expected list: -ln(1 - k'/64) * 64 + margin)

// ---------------------------------------------------------------- prep
__global__ void __launch_bounds__(256) kf_mean_kernel(const float* __restrict__ x, int rows, int n, float* __restrict__ mean) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* p = x + (size_t)row * n;
    float s = 0.f;
    for (int i = lane; i < n; i += 32) s += p[i];
    s = warp_sum(s);
    if (lane == 0) mean[row] = s / (float)n;
}

__device__ __forceinline__ float tf32_rn(float v) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

// one CTA = 32 points of one batch element, all channels
__global__ void __launch_bounds__(256) kf_prep_kernel(const float* __restrict__ x, const float* __restrict__ mean, int c, int n,
                                                     float* __restrict__ xc, float* __restrict__ xT, float* __restrict__ nrm,
                                                     unsigned* __restrict__ maxn) {
    __shared__ float tile[32][33], tilec[32][33];
    __shared__ float part[8][32];
    const int bz = blockIdx.y, n0 = blockIdx.x * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float* xb = x + (size_t)bz * c * n;
    float* xcb = xc + (size_t)bz * n * c;   // centred rows, [n][c] like xT
    float* xtb = xT + (size_t)bz * n * c;
    const float* mb = mean + (size_t)bz * c;
    float acc = 0.f;
    for (int c0 = 0; c0 < c; c0 += 32) {
        for (int ch = ty; ch < 32; ch += 8) {
            float v = 0.f, w = 0.f;
            if (c0 + ch < c && n0 + tx < n) {
                v = xb[(size_t)(c0 + ch) * n + n0 + tx];
                w = tf32_rn(v - mb[c0 + ch]);
                acc = fmaf(w, w, acc);
            }
            tile[ch][tx] = v;
            tilec[ch][tx] = w;
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8)
            if (n0 + r < n && c0 + tx < c) {
                xtb[(size_t)(n0 + r) * c + c0 + tx] = tile[tx][r];
                xcb[(size_t)(n0 + r) * c + c0 + tx] = tilec[tx][r];
            }
        __syncthreads();
    }
    part[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && n0 + tx < n) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) s += part[r][tx];
        nrm[(size_t)bz * n + n0 + tx] = s;
        if (s <= 3.402823466e+38f) atomicMax(&maxn[bz], __float_as_uint(s));   // s >= 0: unsigned order == float order
    }
}

// ---------------------------------------------------------------- tcgen05 helpers
__device__ __forceinline__ uint64_t kf_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // UMMA shared-memory descriptor, K-major, no swizzle (canonical layout ((8,m),(4,2)):((4,SBO),(1,LBO)) in tf32 elements):
    // a core matrix is 8 points x 16 bytes (4 channels) = 128 contiguous bytes; LBO = bytes between the two core matrices that
    // make up K = 8, SBO = bytes between 8-point groups.  Offsets in 16-byte units, descriptor version 1 (sm_100).
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128
constexpr uint32_t KF_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((TF_N >> 3) << 17) | ((TF_M >> 4) << 24);

__device__ __forceinline__ void kf_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
        "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void kf_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void kf_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void kf_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void kf_tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void kf_cp16(uint32_t dst_s, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_s), "l"(src) : "memory");
}
__device__ __forceinline__ void kf_cp_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Bounds of one query.  ni = |xi|^2 and mx = max_j |xj|^2 of the centred, tf32-rounded rows; a16 = bf16 bits (rounded down) of
// the k'-th smallest group minimum of h = ni + nj - 2 G.  With x~ the rounded rows and x^ = x - mean: |x~ - x^| <= 2^-11 |x^|
// per component, so | |x~i - x~j| - |xi - xj| | <= eps = 2^-11 (|x^i| + |x^j|) <= 1.01 * 2^-11 (sqrt(ni) + sqrt(mx)), and the
// computed h differs from |x~i - x~j|^2 by at most acc (rounding of the norms, of ni + nj - 2G, and whatever order the tensor
// core sums its 22-bit-exact products in: 2^-17 (sqrt(ni) + sqrt(mx))^2 is > 100 FP32 ulps of the largest term).  The
// reference chain's own rounding is dc = (c + 4) 2^-23 relative.  Hence
//   U  = (sqrt(A + acc) + eps)^2 (1 + dc)        >= the k'-th smallest REFERENCE distance   (A = the bf16 bound, one ulp up)
//   F  = (sqrt(U (1 + 2 dc)) + eps)^2 + acc      >= h of every candidate with d_ref <= U.
__device__ __forceinline__ float kf_flag_threshold(unsigned a16, float ni, float mx, int c) {
    const float s = __fadd_ru(__fsqrt_ru(ni), __fsqrt_ru(mx));
    const float eps = __fmul_ru(s, 1.01f * 4.8828125e-4f);                 // 2^-11
    const float acc = __fmul_ru(__fmul_ru(s, s), 7.62939453125e-6f);       // 2^-17
    const float dc = (float)(c + 4) * 1.1920929e-7f;                       // 2^-23
    const float a = __uint_as_float((a16 + 1u) << 16);
    float r = __fadd_ru(__fsqrt_ru(__fadd_ru(a, acc)), eps);
    const float u = __fmul_ru(__fmul_ru(r, r), 1.f + dc);
    r = __fadd_ru(__fsqrt_ru(__fmul_ru(u, 1.f + 2.f * dc)), eps);
    return __fadd_ru(__fmul_ru(r, r), acc);
}

// ---------------------------------------------------------------- filter
// grid (n / 128, B), 128 threads.  Dynamic shared memory: A [c/8][32][128 B] | B stage 0, 1 [32][4][128 B] | norms [n] | lists.
__global__ void __launch_bounds__(TF_M, 1) knn_feat_tc_kernel(const float* __restrict__ xc, const float* __restrict__ nrm,
                                                            const unsigned* __restrict__ maxn, int c, int n, int kk,
                                                            int* __restrict__ cand, int* __restrict__ cnt, float* __restrict__ dbg, uint32_t idesc) {
    extern __shared__ __align__(128) unsigned char kf_smem[];
    __shared__ uint64_t bars[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bz = blockIdx.y, m0 = blockIdx.x * TF_M;
    const int chunks = c >> 2;                                       // 16-byte channel chunks of a point
    const uint32_t a_bytes = (uint32_t)TF_M * (uint32_t)c * 4u;
    const uint32_t sbo_a = (uint32_t)chunks * 128u;                  // bytes between the 8-point groups of A
    unsigned char* a_sm = kf_smem;
    unsigned char* b_sm = a_sm + a_bytes;                            // 2 stages of 16 KB
    float* nrm_s = reinterpret_cast<float*>(b_sm + 2 * 16384);       // [n]
    int* lst = reinterpret_cast<int*>(nrm_s + n);                    // [TF_CAP][128]
    const float* xcb = xc + (size_t)bz * n * c;                      // centred, tf32-rounded rows [n][c]

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // A: every (point, 4 channels) piece of 16 bytes.  Piece p: r8 = p % 8, chunk = (p / 8) % chunks, 8-point group = rest --
    // eight consecutive lanes fill one 128-byte core matrix (conflict-free) and a warp reads 8 points x 64 bytes of global memory.
    {
        const uint32_t a_s = smem_u32(a_sm);
        for (int p = tid; p < TF_M * chunks; p += TF_M) {
            const int r8 = p & 7, q = (p >> 3) % chunks, rg = (p >> 3) / chunks;
            kf_cp16(a_s + (uint32_t)rg * sbo_a + (uint32_t)q * 128u + (uint32_t)r8 * 16u, xcb + (size_t)(m0 + rg * 8 + r8) * c + q * 4);
        }
    }
    for (int j = tid; j < n; j += TF_M) nrm_s[j] = nrm[(size_t)bz * n + j];
    kf_cp_wait_all();
    fence_proxy_async();
    kf_fence_before();
    __syncthreads();
    kf_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);      // this warp's 32 lanes of the accumulator

    const float ni = nrm_s[m0 + tid];
    const float mx = __uint_as_float(maxn[bz]);
    const int ntile = n / TF_N, nkb = (c + TF_KB - 1) / TF_KB;
    float mn[TF_N];
#pragma unroll
    for (int i = 0; i < TF_N; ++i) mn[i] = kInf;
    float fv = 0.f;
    int nl = 0;
    bool over = false;
    int it = 0;                                                      // staged K-blocks so far (ring position + phases)

    for (int pass = 0; pass < 2; ++pass) {
        for (int nt = 0; nt < ntile; ++nt) {
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const int buf = it & 1;
                if (it >= 2) mbar_wait(&bars[buf], (unsigned)(((it >> 1) - 1) & 1));   // the MMAs that read this stage are done
                const int qs = min(8, chunks - kb * 8);             // 16-byte chunks of this K-block (even: c % 8 == 0)
                const uint32_t b_s = smem_u32(b_sm + buf * 16384);
                for (int p = tid; p < TF_N * qs; p += TF_M) {
                    const int r8 = p & 7, q = (p >> 3) % qs, rg = (p >> 3) / qs;
                    kf_cp16(b_s + (uint32_t)rg * 1024u + (uint32_t)q * 128u + (uint32_t)r8 * 16u,
                            xcb + (size_t)(nt * TF_N + rg * 8 + r8) * c + kb * TF_KB + q * 4);
                }
                kf_cp_wait_all();
                fence_proxy_async();
                __syncthreads();
                if (tid == 0) {
                    kf_fence_after();
                    for (int kc = 0; kc < (qs >> 1); ++kc)
                        kf_mma(tmem, kf_desc(smem_u32(a_sm) + (uint32_t)(kb * 8 + 2 * kc) * 128u, 128u, sbo_a),
                               kf_desc(b_s + (uint32_t)(2 * kc) * 128u, 128u, 1024u), (kb | kc) ? 1u : 0u, idesc);
                    kf_commit(&bars[buf]);
                }
            }
            // the tile's last commit covers every MMA issued before it
            mbar_wait(&bars[(it - 1) & 1], (unsigned)(((it - 1) >> 1) & 1));
            kf_fence_after();
#pragma unroll
            for (int ch = 0; ch < TF_N / 32; ++ch) {
                float g[32];
                kf_tmem_ld32(trow + (uint32_t)(ch * 32), g);
                if (dbg && pass == 0 && nt == 0 && blockIdx.x == 0 && blockIdx.y == 0) {
#pragma unroll
                    for (int u = 0; u < 32; ++u) dbg[tid * TF_N + ch * 32 + u] = g[u] + 1000.f;
                }
                const float4* nj4 = reinterpret_cast<const float4*>(nrm_s + nt * TF_N + ch * 32);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 nj = nj4[q];
                    const float njv[4] = {nj.x, nj.y, nj.z, nj.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float t = __fmaf_rn(-2.f, g[4 * q + u], njv[u]);
                        if (pass == 0) {
                            mn[ch * 32 + 4 * q + u] = fminf(mn[ch * 32 + 4 * q + u], t);   // NaN never wins
                        } else if (__fadd_rn(t, ni) <= fv) {
                            if (nl < TF_CAP) lst[nl * TF_M + tid] = nt * TF_N + ch * 32 + 4 * q + u;
                            else over = true;
                            nl += nl < TF_CAP ? 1 : 0;
                        }
                    }
                }
            }
            kf_fence_before();
            __syncthreads();                                         // TMEM is read: the next tile may overwrite it
        }
        if (pass == 0) {
            // bound: k'-th smallest of the 64 group minima (group = subgroups g and g + 64), bf16 rounded down, clamped at 0
            auto grp = [&](int i) -> unsigned {
                const float h = fmaxf(__fadd_rn(fminf(mn[i], mn[i + 64]), ni), 0.f);
                return __float_as_uint(h) >> 16;
            };
            const unsigned a16 = kq_kth_of_64(grp, kk);
            if (a16 >= 0x7f80u || !(ni <= 3.402823466e+38f)) {
                over = true;                                         // fewer than k' finite groups, or a non-finite query
                fv = -1.f;
            } else {
                fv = kf_flag_threshold(a16, ni, mx, c);
                if (!(fv <= 3.402823466e+38f)) {
                    over = true;
                    fv = -1.f;
                }
            }
        }
    }
    // lists out (flag = -1: the exact kernel recomputes this query)
    const size_t q = (size_t)bz * n + m0 + tid;
    if (dbg) {
        cnt[q] = -1;                                                 // bring-up dump: the Gram tile sits in the list buffer
    } else if (over || nl < kk) {
        cnt[q] = -1;
    } else {
        cnt[q] = nl;
        for (int e = 0; e < nl; ++e) cand[q * TF_CAP + e] = lst[e * TF_M + tid];
    }
    kf_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

// ---------------------------------------------------------------- exact re-rank: warp = query, lane = candidate
__global__ void __launch_bounds__(256) knn_feat_rerank_kernel(const float* __restrict__ xT, const int* __restrict__ cand,
                                                             const int* __restrict__ cnt, int c, int n, int k, int skip, int total,
                                                             long long* __restrict__ idx, float* __restrict__ dist2) {
    const int q = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= total) return;
    const int m = cnt[q];
    if (m < 0) return;                                               // flagged: the exact kernel writes this query
    const int bz = q / n;
    const int j = lane < m ? cand[(size_t)q * TF_CAP + lane] : -1;
    const float* xi = xT + (size_t)q * c;
    const float* xj = xT + ((size_t)bz * n + max(j, 0)) * c;
    float acc = 0.f;
    if ((c & 3) == 0) {
        for (int ch = 0; ch < c; ch += 4) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(xi + ch)), b = __ldg(reinterpret_cast<const float4*>(xj + ch));
            float d = __fsub_rn(a.x, b.x);
            acc = __fmaf_rn(d, d, acc);
            d = __fsub_rn(a.y, b.y);
            acc = __fmaf_rn(d, d, acc);
            d = __fsub_rn(a.z, b.z);
            acc = __fmaf_rn(d, d, acc);
            d = __fsub_rn(a.w, b.w);
            acc = __fmaf_rn(d, d, acc);
        }
    } else {
        for (int ch = 0; ch < c; ++ch) {
            const float d = __fsub_rn(__ldg(xi + ch), __ldg(xj + ch));
            acc = __fmaf_rn(d, d, acc);
        }
    }
    const unsigned long long key = j >= 0 ? (((unsigned long long)__float_as_uint(acc) << 32) | (unsigned)j) : ~0ull;
    int rank = 0;
#pragma unroll
    for (int o = 0; o < 32; ++o) rank += __shfl_sync(kFull, key, o) < key ? 1 : 0;
    if (j >= 0 && rank >= skip && rank < skip + k) {
        idx[(size_t)q * k + rank - skip] = j;
        if (dist2) dist2[(size_t)q * k + rank - skip] = acc;
    }
}

// ---------------------------------------------------------------- host side
static size_t kf_align(size_t v) { return (v + 255) & ~(size_t)255; }

bool knn_feat_tc_eligible(int c, int n, int k, int skip) {
    return c >= 8 && c <= 256 && (c & 7) == 0 && n >= 128 && n <= 4096 && (n & 127) == 0 && k + skip <= TF_KMAX && k >= 1;
}
size_t knn_feat_tc_workspace(int b, int c, int n) {
    const size_t bc = (size_t)b * c, bn = (size_t)b * n;
    return kf_align(bc * 4) + 2 * kf_align(bc * n * 4) + kf_align(bn * 4) + kf_align((size_t)b * 4) + kf_align(bn * TF_CAP * 4) + kf_align(bn * 4) + 256;
}

// Runs prep + filter + re-rank; *flags receives the per-query count array (negative = recompute with the exact kernel).
int knn_feat_tc_launch(const float* x, int b, int c, int n, int k, int skip, long long* idx, float* dist2, void* ws, const int** flags,
                       cudaStream_t st) {
    const size_t bc = (size_t)b * c, bn = (size_t)b * n;
    unsigned char* p = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
    float* mean = reinterpret_cast<float*>(p);
    p += kf_align(bc * 4);
    float* xc = reinterpret_cast<float*>(p);
    p += kf_align(bc * n * 4);
    float* xT = reinterpret_cast<float*>(p);
    p += kf_align(bc * n * 4);
    float* nrm = reinterpret_cast<float*>(p);
    p += kf_align(bn * 4);
    unsigned* maxn = reinterpret_cast<unsigned*>(p);
    p += kf_align((size_t)b * 4);
    int* cand = reinterpret_cast<int*>(p);
    p += kf_align(bn * TF_CAP * 4);
    int* cnt = reinterpret_cast<int*>(p);
    PDGN_CUDA(cudaMemsetAsync(maxn, 0, (size_t)b * 4, st));
    kf_mean_kernel<<<(unsigned)((bc + 7) / 8), 256, 0, st>>>(x, (int)bc, n, mean);
    PDGN_CHECK_LAUNCH();
    kf_prep_kernel<<<dim3((n + 31) / 32, b), 256, 0, st>>>(x, mean, c, n, xc, xT, nrm, maxn);
    PDGN_CHECK_LAUNCH();
    const size_t smem = (size_t)TF_M * c * 4 + 2 * 16384 + (size_t)n * 4 + (size_t)TF_CAP * TF_M * 4;
    PDGN_CUDA(cudaFuncSetAttribute(knn_feat_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_feat_tc_kernel<<<dim3(n / TF_M, b), TF_M, smem, st>>>(xc, nrm, maxn, c, n, k + skip, cand, cnt,
                                                                  tune_env("PDGN_KNN_FEAT_DBG") ? reinterpret_cast<float*>(cand) : nullptr,
                                                                  tune_env("PDGN_KNN_FEAT_IDESC") ? (uint32_t)strtoul(tune_env("PDGN_KNN_FEAT_IDESC"), nullptr, 16) : KF_IDESC);
    PDGN_CHECK_LAUNCH();
    knn_feat_rerank_kernel<<<(unsigned)((bn + 7) / 8), 256, 0, st>>>(xT, cand, cnt, c, n, k, skip, (int)bn, idx, dist2);
    PDGN_CHECK_LAUNCH();
    *flags = cnt;
    return PDGN_OK;
}

}  // namespace pdgn
