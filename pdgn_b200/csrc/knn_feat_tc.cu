// knn_feat_tc.cu -- feature-space kNN graph of the generator on the 5th-generation tensor cores (tcgen05 + TMEM), with an exact
// FP32 re-rank, so that the indices are still the oracle's (north_star: "feature-space kNN (C >= 64) may use tcgen05 for the
// Gram tile only if the top-k candidates are re-ranked in exact FP32").
//
// Replaces, in get_edge_features / get_edge_features_xyz (models/PDGNet_v2.py:449-459, :492-502), the [B,N,N] cuBLAS Gram
// matrix + full torch.sort of every row + slice, and csrc/knn_feat.cu's FP32 SIMT kernel (2 FMA-pipe instructions per pair
// and channel) for the shapes below.  Same contract as knn_feat.cu: d2(i,j) = ONE fma chain over c = 0..C-1 of
// (x[c,i] - x[c,j])^2, total order (d2, index), ranks skip..skip+k-1; bit-exact against oracle_knn_feat.
//
//   prep    (kf_mean_kernel, kf_prep_kernel)  channel means; the centred copy xc = tf32_rn(x - mean), PRE-TILED in the tensor
//           core's operand order (16 KB blocks of 128 points x 32 channels, K-major rows of 128 bytes, SWIZZLE_128B) so that an
//           operand stage is one contiguous bulk copy; squared norms of the rounded rows; xT = x transposed [B,N,C] for the
//           re-rank.  Distances are shift invariant: centring shrinks the norms to the spread of the features, and values that
//           ARE tf32 make the tensor core's operand truncation a no-op.
//   filter  (knn_feat_tc_kernel<128|256>)  CTA = 128 queries (= the 128 TMEM lanes) x all candidates; 9 warps.  Thread 256
//           copies (cp.async.bulk into a ring of stages, "full" mbarriers) and issues tcgen05.mma.kind::tf32 (M = 128,
//           N = 128 or 256, K = 8) for every 8 channels; tcgen05.commit releases the stage and hands a finished accumulator to
//           the epilogue; two accumulators in TMEM, so the K loop of the next tile runs under the epilogue of this one.  Warps
//           0-7 read accumulators with tcgen05.ld, two threads per query (every other 32-column chunk = disjoint subgroups):
//             pass A  h = |xi|^2 + |xj|^2 - 2 G as running minima of 128 STRIDED subgroups (candidate j -> subgroup j mod 128)
//                     in registers; bound = k'-th smallest of the 64 group minima (selnet.cuh, the xyz kernel's network);
//             pass B  the Gram tiles are computed AGAIN (the tensor-core time is small next to the operand traffic) and every
//                     candidate with h <= F is appended to the query's list in global memory (<= 64 entries).
//           The margins are rigorous (kf_flag_threshold): U bounds the k'-th smallest REFERENCE distance from above and every
//           candidate with d_ref <= U has h <= F.  Nothing depends on the tensor core's internal summation order beyond a
//           generous absolute error term.
//   rank    (knn_feat_rerank_kernel)  warp = query, lane = candidate: the exact reference chain over rows fetched coalesced into
//           shared memory, rank by counting over the (d2, index) keys, ranks skip..skip+k-1 stored; entries 32..63 in a second pass.
//   Queries whose list overflowed, or that saw fewer than k' finite group minima (duplicates, clusters, NaN), are flagged and
//   counted: up to KF_BRUTE_MAX of them are recomputed inside the re-rank kernel (kf_brute_query: warp = query, all n exact
//   distances, repeated minimum search), more than that (collapsed clouds) by knn_feat.cu's exact 64-query CTAs.
// Bring-up history and measurements: profiles/r02_knn_feat_tc.txt.
#include "common.cuh"
#include "selnet.cuh"

namespace pdgn {

constexpr int TF_M = 128;    // queries per CTA = TMEM lanes
constexpr int TF_SG = 128;   // strided subgroups of the bound (candidate j -> subgroup j % 128); accumulator tiles are 128 or 256 wide
constexpr int TF_KB = 32;    // channels per staged K-block (4 MMAs of K = 8)
constexpr int TF_NST = 4;    // B ring stages (3 when the norms of a large cloud need the room)
constexpr int TF_CAP = 64;   // list entries per query (global memory; the re-rank takes 32 in its fast pass, the rest one by one)
constexpr int TF_KMAX = 20;  // k + skip the bound network is tuned for (expected list: -ln(1 - k'/64) * 64 + margin)

// ---------------------------------------------------------------- prep
__global__ void __launch_bounds__(256) kf_mean_kernel(const float* __restrict__ x, int rows, int n, float* __restrict__ mean,
                                                     int* __restrict__ nflag) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *nflag = 0;            // the filter (two launches later) counts its flagged queries here
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

constexpr int KF_BRUTE_MAX = 512;   // flagged queries the re-rank kernel recomputes itself (one warp each, ~0.3 ms of latency); beyond
                                    // that (degenerate clouds: every query flagged) knn_feat.cu's 64-query CTAs are the faster fallback
constexpr int KF_SLABS = 4;   // channel slabs of the prep kernel (partial norms per slab, summed in a fixed order by the filter)

// one CTA = 32 points of one batch element x one slab of channels (blockIdx.z; slab = cps channels, a multiple of 32)
__global__ void __launch_bounds__(256) kf_prep_kernel(const float* __restrict__ x, const float* __restrict__ mean, int c, int n, int cps,
                                                     float* __restrict__ xc, float* __restrict__ xT, float* __restrict__ nrm) {
    __shared__ float tile[32][33], tilec[32][33];
    __shared__ float part[8][32];
    const int bz = blockIdx.y, n0 = blockIdx.x * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float* xb = x + (size_t)bz * c * n;
    // centred rows, PRE-TILED in the tensor core's operand order: one 16 KB block per (128 points, 32 channels) = 128 rows of
    // 128 bytes (K-major), the 16-byte chunks of a row XOR-swizzled with (row % 8) = the canonical SWIZZLE_128B UMMA layout
    // (8-row atoms of 1024 bytes), so a whole operand stage is ONE contiguous bulk copy.  (The no-swizzle core-matrix order
    // [8-point group][4-channel chunk][point][4 floats] was the first working version and runs at the same speed.)  Channels
    // are padded with zeros to a multiple of 32.
    const int nkb = (c + 31) >> 5;
    float* xcb = xc + (size_t)bz * n * nkb * 32;
    float* xtb = xT + (size_t)bz * n * c;
    const float* mb = mean + (size_t)bz * c;
    float acc = 0.f;
    const int c_lo = blockIdx.z * cps, c_hi = min(c, c_lo + cps);
    for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
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
            if (n0 + r < n) {
                if (c0 + tx < c) xtb[(size_t)(n0 + r) * c + c0 + tx] = tile[tx][r];
                const int pnt = n0 + r;
                xcb[((size_t)(pnt >> 7) * nkb + (c0 >> 5)) * 4096 + (pnt & 127) * 32 + (((tx >> 2) ^ (pnt & 7)) << 2) + (tx & 3)] = tilec[tx][r];
            }
        __syncthreads();
    }
    part[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && n0 + tx < n) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) s += part[r][tx];
        nrm[((size_t)bz * KF_SLABS + blockIdx.z) * n + n0 + tx] = s;          // partial squared norm of this slab (0 for an empty slab)
    }
}

// ---------------------------------------------------------------- tcgen05 helpers
__device__ __forceinline__ uint64_t kf_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // UMMA shared-memory descriptor, K-major, SWIZZLE_128B (canonical layout ((8,m),(4,2)):((32,SBO),(1,4)) in tf32 elements
    // under Swizzle<3,4,3>): a row is 128 contiguous bytes (32 channels), 8 rows form a 1024-byte atom whose 16-byte chunks are
    // XORed with the row number, SBO = bytes between atoms; the leading offset is unused for a swizzled K-major operand.
    // A K = 8 step inside the atom is addressed by advancing the start address by 32 bytes.  Offsets in 16-byte units,
    // descriptor version 1 (sm_100).  The tile base must be 1024-byte aligned (base offset 0).
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) |
           (2ull << 61);   // layout type 2 = SWIZZLE_128B
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, N = tn, M = 128
constexpr uint32_t kf_idesc(int tn) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(tn >> 3) << 17) | ((uint32_t)(TF_M >> 4) << 24); }

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
__device__ __forceinline__ void mbar_arrive_plain(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// Bounds of one query.  ni = |xi|^2 and mx = max_j |xj|^2 of the centred, tf32-rounded rows; a16 = bf16 bits (rounded down) of
// the k'-th smallest group minimum of h = ni + nj - 2 G.  With x~ the rounded rows and x^ = x - mean: |x~ - x^| <= 2^-11 |x^|
// per component, so | |x~i - x~j| - |xi - xj| | <= eps = 2^-11 (|x^i| + |x^j|) <= 1.01 * 2^-11 (sqrt(ni) + sqrt(mx)), and the
// computed h differs from |x~i - x~j|^2 by at most acc.  The tensor core multiplies the 11-bit operands exactly (22-bit products)
// and sums c of them in FP32 in an order and rounding mode we do not rely on: even truncating after every addition the sum is
// within c 2^-23 sum|x~ic x~jc| <= c 2^-23 |x~i||x~j| of the exact one; h doubles that, and the norms (FP32 sums of c squares)
// and the two final additions add a few ulps of ni + nj.  With s = |x~i| + max|x~j| (s^2 >= 4 |x~i||x~j|, s^2 >= ni + nj / 2):
// acc = (c + 16) 2^-23 s^2 >= 2 (2 c 2^-23 |x~i||x~j|) + 16 2^-23 (ni + nj) / 2 -- twice the worst case.  The reference chain's
// own rounding is dc = (c + 4) 2^-23 relative.  Hence
//   U  = (sqrt(A + acc) + eps)^2 (1 + dc)        >= the k'-th smallest REFERENCE distance   (A = the bf16 bound, one ulp up)
//   F  = (sqrt(U (1 + 2 dc)) + eps)^2 + acc      >= h of every candidate with d_ref <= U.
__device__ __forceinline__ float kf_flag_threshold(unsigned a16, float ni, float mx, int c) {
    const float s = __fadd_ru(__fsqrt_ru(ni), __fsqrt_ru(mx));
    const float eps = __fmul_ru(s, 1.01f * 4.8828125e-4f);                 // 2^-11
    const float acc = __fmul_ru(__fmul_ru(s, s), (float)(c + 16) * 1.1920929e-7f);   // (c + 16) 2^-23 s^2
    const float dc = (float)(c + 4) * 1.1920929e-7f;                       // 2^-23
    const float a = __uint_as_float((a16 + 1u) << 16);
    float r = __fadd_ru(__fsqrt_ru(__fadd_ru(a, acc)), eps);
    const float u = __fmul_ru(__fmul_ru(r, r), 1.f + dc);
    r = __fadd_ru(__fsqrt_ru(__fmul_ru(u, 1.f + 2.f * dc)), eps);
    return __fadd_ru(__fmul_ru(r, r), acc);
}

// ---------------------------------------------------------------- filter
// grid (n / 128, B), 128 threads.  Dynamic shared memory: A [c/8][32][128 B] | B stage 0, 1 [32][4][128 B] | norms [n] | lists.
constexpr int TF_E = 2 * TF_M;     // 8 epilogue warps: two threads per query (= TMEM lane), each takes every other 32-column chunk
constexpr int TF_T = TF_E + 32;    // + 1 warp whose lane 0 copies and issues the MMAs

template <int TN>
__global__ void __launch_bounds__(TF_T, 1) knn_feat_tc_kernel(const float* __restrict__ xc, const float* __restrict__ nrm,
                                                            int c, int n, int kk,
                                                            int* __restrict__ cand, int* __restrict__ cnt, int* __restrict__ nflag, int nst, float* __restrict__ dbg,
                                                            uint32_t idesc) {
    extern __shared__ __align__(1024) unsigned char kf_smem[];
    __shared__ uint64_t bars[TF_NST], fullb[TF_NST], abar, tfull[2], tempty[2];   // bars: stage consumed by the tensor core; fullb: stage filled
    __shared__ uint32_t tmem_slot;
    __shared__ float wmax[TF_T / 32];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int bz = blockIdx.y, m0 = blockIdx.x * TF_M;
    const int nkb = (c + TF_KB - 1) / TF_KB;                          // K-blocks (channels zero-padded to a multiple of 32)
    unsigned char* a_sm = kf_smem;                                   // nkb blocks of 16 KB: the CTA's 128 queries, all channels
    constexpr int NB = TN / 128;                                     // 16 KB blocks (128 candidates x 32 channels) per stage
    constexpr uint32_t SB = NB * 16384u;                             // bytes per ring stage
    unsigned char* b_sm = a_sm + (size_t)nkb * 16384;                // nst ring stages
    float* nrm_s = reinterpret_cast<float*>(b_sm + (size_t)nst * SB);   // [n]
    int* lst = reinterpret_cast<int*>(nrm_s + n);                    // [TF_CAP][128]
    const float* xcb = xc + (size_t)bz * n * nkb * 32;               // this batch element's pre-tiled blocks (kf_prep_kernel)

    if (tid == 0) {
        for (int i = 0; i < nst; ++i) {
            mbar_init(&bars[i], 1);
            mbar_init(&fullb[i], 1);
        }
        mbar_init(&abar, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);      // accumulator i holds a finished tile (tcgen05.commit of the tile's last K-block)
            mbar_init(&tempty[i], TF_E);  // accumulator i has been read by all 256 epilogue threads
        }
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"((uint32_t)(2 * TN)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    const bool issuer = tid == TF_E;                                 // lane 0 of warp 8
    if (issuer) {                                                    // A: the query tile's blocks, one bulk copy each
        mbar_expect_tx(&abar, (unsigned)nkb * 16384u);
        for (int kb = 0; kb < nkb; ++kb)
            bulk_g2s(a_sm + (size_t)kb * 16384, xcb + ((size_t)blockIdx.x * nkb + kb) * 4096, 16384u, &abar);
    }
    // squared norms: the prep kernel's per-slab partial sums added in a fixed order; their maximum over the cloud on the way
    float lmax = 0.f;
    for (int j = tid; j < n; j += TF_T) {
        float v = 0.f;
#pragma unroll
        for (int sl = 0; sl < KF_SLABS; ++sl) v += nrm[((size_t)bz * KF_SLABS + sl) * n + j];
        nrm_s[j] = v;
        if (v <= 3.402823466e+38f) lmax = fmaxf(lmax, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(kFull, lmax, o));
    if ((tid & 31) == 0) wmax[warp] = lmax;
    kf_fence_before();
    __syncthreads();
    kf_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);   // this warp's 32 lanes of the accumulators

    const float ni = nrm_s[m0 + (tid & (TF_M - 1))];
    unsigned short* exch = reinterpret_cast<unsigned short*>(lst);   // [2][32][128] group minima between the passes (the lists are idle)
    int* cnt_s = lst + 32 * TF_M;                                    // [128] list lengths, [128] overflow flags (after the 16 KB exchange area)
    float mx = 0.f;
#pragma unroll
    for (int w = 0; w < TF_T / 32; ++w) mx = fmaxf(mx, wmax[w]);
    const int ntile = n / TN;
    // Epilogue thread (row, half): half h takes the 32-column chunks ch with ch % 2 == h, i.e. the subgroups [32h, 32h + 32) and
    // [64 + 32h, 96 + 32h) -- the two threads of a query own DISJOINT groups (i, i + 64), 32 each.
    const int row = tid & (TF_M - 1), half = (tid >> 7) & 1;
    float mn[TF_SG / 2];
#pragma unroll
    for (int i = 0; i < TF_SG / 2; ++i) mn[i] = kInf;
    float fv = 0.f, ft = -kInf;
    bool over = false;
    // Warp-specialised: thread 128 (lane 0 of warp 4) is the producer AND the MMA issuer -- one 16 KB bulk copy per 128 candidates
    // and K-block (TMA engine, completes on the stage's "full" mbarrier), up to nst - 1 stages ahead of the tensor core; four
    // tcgen05.mma (K = 8 each) per K-block; tcgen05.commit releases the stage, and at a tile's last K-block hands the accumulator
    // to the epilogue.  Two accumulators in TMEM: the K loop of tile g + 1 runs under the epilogue of tile g.  Warps 0-3
    // (thread = query = TMEM lane) only read accumulators.  Both passes walk the same tile sequence; the threshold of pass B is
    // ready long before its first tile is (the issuer does not wait for it).
    const int T = 2 * ntile * nkb;
    if (issuer) {
        int p_kb = 0, p_nt = 0, p_buf = 0, p_round = 0, p_left = T;     // producer cursor
        auto issue_next = [&]() {
            mbar_expect_tx(&fullb[p_buf], SB);
#pragma unroll
            for (int h = 0; h < NB; ++h)       // consecutive 128-candidate blocks are consecutive 8-point groups of one operand
                bulk_g2s(b_sm + (size_t)p_buf * SB + (size_t)h * 16384, xcb + ((size_t)(p_nt * NB + h) * nkb + p_kb) * 4096, 16384u, &fullb[p_buf]);
            if (++p_kb == nkb) {
                p_kb = 0;
                if (++p_nt == ntile) p_nt = 0;
            }
            if (++p_buf == nst) {
                p_buf = 0;
                ++p_round;
            }
            --p_left;
        };
        for (int t = 0; t < nst - 1; ++t)                            // all but one stage
            if (p_left > 0) issue_next();
        mbar_wait(&abar, 0u);
        int buf = 0, round = 0;                                      // consumer side of the ring
        const uint32_t a_s0 = smem_u32(a_sm), b_s0 = smem_u32(b_sm);
        for (int g = 0; g < 2 * ntile; ++g) {
            const int acc = g & 1;
            if (g >= 2) mbar_wait(&tempty[acc], (unsigned)(((g >> 1) - 1) & 1));   // the epilogue has drained this accumulator
            kf_fence_after();
            const uint32_t tacc = tmem + (uint32_t)(acc * TN);
            for (int kb = 0; kb < nkb; ++kb) {
                const uint32_t b_s = b_s0 + (uint32_t)buf * SB, a_s = a_s0 + (uint32_t)kb * 16384u;
                mbar_wait(&fullb[buf], (unsigned)(round & 1));       // the block has landed (async proxy: no proxy fence needed)
                kf_fence_after();
#pragma unroll
                for (int kc = 0; kc < 4; ++kc)
                    if (!(idesc & 1u))                               // idesc bit 0 (sparse id, unused): ablation without MMAs
                        kf_mma(tacc, kf_desc(a_s + (uint32_t)kc * 32u, 16u, 1024u), kf_desc(b_s + (uint32_t)kc * 32u, 16u, 1024u),
                               (kb | kc) ? 1u : 0u, idesc);
                kf_commit(&bars[buf]);
                if (kb == nkb - 1) kf_commit(&tfull[acc]);           // the tile's last commit covers every MMA issued before it
                // refill the stage the PREVIOUS iteration used, once its MMAs are done: this iteration's MMAs are already queued
                // behind them, so the tensor core does not idle while the issuer waits here
                if (p_left > 0) {
                    if (p_round > 0) mbar_wait(&bars[p_buf], (unsigned)((p_round - 1) & 1));
                    issue_next();
                }
                if (++buf == nst) {
                    buf = 0;
                    ++round;
                }
            }
        }
    } else if (tid < TF_E) {
        if (half == 0) {
            cnt_s[row] = 0;
            cnt_s[TF_M + row] = 0;
        }
        int g = 0;
        for (int pass = 0; pass < 2; ++pass) {
            for (int nt = 0; nt < ntile; ++nt, ++g) {
                const int acc = g & 1;
                mbar_wait(&tfull[acc], (unsigned)((g >> 1) & 1));
                kf_fence_after();
#pragma unroll
                for (int cq = 0; cq < ((idesc & 2u) ? 0 : TN / 64); ++cq) {   // idesc bit 1 (unused sparse id): ablation without epilogue
                    const int ch = 2 * cq + half;                    // this thread's chunk of the pair
                    float gv[32];
                    kf_tmem_ld32(trow + (uint32_t)(acc * TN + ch * 32), gv);
                    if (dbg && pass == 0 && nt == 0 && blockIdx.x == 0 && blockIdx.y == 0) {
#pragma unroll
                        for (int u = 0; u < 32; ++u) dbg[row * TN + ch * 32 + u] = gv[u] + 1000.f;
                    }
                    const float4* nj4 = reinterpret_cast<const float4*>(nrm_s + nt * TN + ch * 32);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 nj = nj4[q];
                        const float njv[4] = {nj.x, nj.y, nj.z, nj.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float t = __fmaf_rn(-2.f, gv[4 * q + u], njv[u]);
                            if (pass == 0) {
                                // chunk ch covers subgroups (ch % 4) * 32 ..: local slot = lower / upper 32 of this thread's 64
                                mn[(cq & 1) * 32 + 4 * q + u] = fminf(mn[(cq & 1) * 32 + 4 * q + u], t);   // NaN never wins
                            } else if (t <= ft) {                    // one compare per candidate: ft >= every t with fl(t + ni) <= fv
                                const int slot = atomicAdd(&cnt_s[row], 1);          // the query's two threads share its list
                                if (slot < TF_CAP) cand[((size_t)bz * n + m0 + row) * TF_CAP + slot] = nt * TN + ch * 32 + 4 * q + u;
                                else cnt_s[TF_M + row] = 1;
                            }
                        }
                    }
                }
                kf_fence_before();
                mbar_arrive_plain(&tempty[acc]);                     // this thread has read the accumulator
            }
            if (pass == 0) {
                // this thread's 32 group minima (group = subgroups s and s + 64), bf16 rounded down, clamped at 0 -> shared memory;
                // both threads of the query then run the same network over all 64 and get the same threshold
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float h = fmaxf(__fadd_rn(fminf(mn[i], mn[i + 32]), ni), 0.f);
                    exch[(half * 32 + i) * TF_M + row] = (unsigned short)(__float_as_uint(h) >> 16);
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");        // the 8 epilogue warps only
                auto grp = [&](int i) -> unsigned { return exch[i * TF_M + row]; };
                const unsigned a16 = kq_kth_of_64(grp, kk);
                asm volatile("bar.sync 1, 256;" ::: "memory");        // exch is read: the lists may overwrite it
                if (a16 >= 0x7f80u || !(ni <= 3.402823466e+38f)) {
                    over = true;                                     // fewer than k' finite groups, or a non-finite query
                    fv = -1.f;
                } else {
                    fv = kf_flag_threshold(a16, ni, mx, c);
                    if (!(fv <= 3.402823466e+38f)) {
                        over = true;
                        fv = -1.f;
                    } else {
                        // pass B compares t = fl(nj - 2G) itself: fl(t + ni) <= fv implies t + ni <= fv (1 + 2^-23), hence
                        // t <= ft := ru(ru(fv (1 + 2^-22)) - ni) + one more ulp of slack
                        const float up = __fmul_ru(fv, 1.f + 2.3841858e-7f);
                        ft = __fadd_ru(__fsub_ru(up, ni), __fmul_ru(fabsf(up) + fabsf(ni), 1.1920929e-7f));
                    }
                }
            }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");                // both halves have appended
        // lists out (flag = -1: the exact brute force recomputes this query)
        if (half == 0) {
            const size_t q = (size_t)bz * n + m0 + row;
            const int nlist = cnt_s[row];
            if (dbg || over || cnt_s[TF_M + row] || nlist > TF_CAP || nlist < kk) {
                cnt[q] = -1;
                atomicAdd(nflag, 1);                                  // how many: decides who recomputes them (KF_BRUTE_MAX)
            } else {
                cnt[q] = nlist;                                       // the entries are already in cand[q][0..nlist)
            }
        }
    }
    kf_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)(2 * TN)) : "memory");
}

// ---------------------------------------------------------------- exact re-rank: warp = query, lane = candidate
// The exact reference chain needs every channel of every candidate row (~14 rows of C floats per query, from the L2-resident
// xT).  The rows are fetched COALESCED, 64 channels at a time (a half-warp per row), into a padded shared-memory tile, and each
// lane then walks its own row from there: the first version let every lane read its row straight from global memory (16 bytes
// per lane and request, 14+ lines per warp request) and ran at 3.5 TB/s of L2 traffic.  The chain stays strictly sequential
// in c (the accumulator is carried across the 64-channel rounds).
__device__ __noinline__ void kf_brute_query(const float* __restrict__ xT, int q, int c, int n, int k, int skip, int lane, float* __restrict__ dsm,
                               long long* __restrict__ idx, float* __restrict__ dist2);

constexpr int RR_CH = 64;            // channels per round
constexpr int RR_RS = RR_CH + 4;     // row stride in floats: 272 bytes, conflict-free LDS.128 for 32 lanes on 32 different rows
__global__ void __launch_bounds__(256) knn_feat_rerank_kernel(const float* __restrict__ xT, const int* __restrict__ cand,
                                                             const int* __restrict__ cnt, const int* __restrict__ nflag, int c, int n, int k,
                                                             int skip, int total, long long* __restrict__ idx, float* __restrict__ dist2) {
    extern __shared__ __align__(16) float rows_s[];                  // [8 warps][33 rows][RR_RS]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * 8 + warp;
    if (q >= total) return;
    const int mraw = cnt[q];
    const int m = min(mraw, 32);
    float* my = rows_s + (size_t)warp * 33 * RR_RS;
    if (mraw < 0) {                                                     // flagged by the filter: exact brute force (few queries)
        if (n <= 33 * RR_RS && *nflag <= KF_BRUTE_MAX) kf_brute_query(xT, q, c, n, k, skip, lane, my, idx, dist2);
        return;                                                      // otherwise knn_feat.cu's kernel recomputes it
    }
    const int bz = q / n;
    const int mfull = mraw;                                          // up to TF_CAP entries; the staged pass below takes the first 32
    const int j = lane < mfull ? cand[(size_t)q * TF_CAP + lane] : -1;
    const int half = lane >> 4, li = lane & 15;
    float acc = 0.f;
    // row 0 = the query, row 1 + l = lane l's candidate; two rows per step, 16 lanes x 16 bytes each.  The first RR_PF steps
    // (rows 0..15: the whole list in the common case) are software-pipelined through registers: the next round's loads are in
    // flight while this round's chains run.
    constexpr int RR_PF = 8;
    const float* srcp[RR_PF];
#pragma unroll
    for (int i = 0; i < RR_PF; ++i) {
        const int row = 2 * i + half;
        const int jj = __shfl_sync(kFull, j, max(row - 1, 0));
        srcp[i] = row > m ? nullptr : (row == 0 ? xT + (size_t)q * c : xT + ((size_t)bz * n + jj) * c);
    }
    float4 pf[RR_PF];
    auto prefetch = [&](int c0) {
        const int cw = min(RR_CH, c - c0);
#pragma unroll
        for (int i = 0; i < RR_PF; ++i)
            if (srcp[i] && li * 4 < cw) pf[i] = __ldg(reinterpret_cast<const float4*>(srcp[i] + c0) + li);
    };
    prefetch(0);
    for (int c0 = 0; c0 < c; c0 += RR_CH) {
        const int cw = min(RR_CH, c - c0);                           // multiple of 4 (c % 8 == 0 on this path)
#pragma unroll
        for (int i = 0; i < RR_PF; ++i)
            if (srcp[i] && li * 4 < cw) *reinterpret_cast<float4*>(my + (2 * i + half) * RR_RS + li * 4) = pf[i];
        for (int r0 = 2 * RR_PF; r0 <= m; r0 += 2) {                 // lists beyond 15 candidates: straight to shared memory
            const int row = r0 + half;
            const int jj = __shfl_sync(kFull, j, max(row - 1, 0));
            if (row <= m && li * 4 < cw)
                *reinterpret_cast<float4*>(my + row * RR_RS + li * 4) = __ldg(reinterpret_cast<const float4*>(xT + ((size_t)bz * n + jj) * c + c0) + li);
        }
        if (c0 + RR_CH < c) prefetch(c0 + RR_CH);
        __syncwarp();
        if (j >= 0) {
            const float4* a4 = reinterpret_cast<const float4*>(my);
            const float4* b4 = reinterpret_cast<const float4*>(my + (1 + lane) * RR_RS);
#pragma unroll 4
            for (int u = 0; u < (cw >> 2); ++u) {
                const float4 a = a4[u], b = b4[u];
                float d = __fsub_rn(a.x, b.x);
                acc = __fmaf_rn(d, d, acc);
                d = __fsub_rn(a.y, b.y);
                acc = __fmaf_rn(d, d, acc);
                d = __fsub_rn(a.z, b.z);
                acc = __fmaf_rn(d, d, acc);
                d = __fsub_rn(a.w, b.w);
                acc = __fmaf_rn(d, d, acc);
            }
        }
        __syncwarp();
    }
    // entries 32..63 (rare: the list is ~14 long): one more candidate per lane, its row read straight from global memory
    int j1 = -1;
    float acc1 = 0.f;
    if (mfull > 32) {                                                // warp-uniform
        j1 = lane + 32 < mfull ? cand[(size_t)q * TF_CAP + 32 + lane] : -1;
        if (j1 >= 0) {
            const float* xi = xT + (size_t)q * c;
            const float* xj = xT + ((size_t)bz * n + j1) * c;
            for (int ch = 0; ch < c; ch += 4) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(xi + ch)), b = __ldg(reinterpret_cast<const float4*>(xj + ch));
                float d = __fsub_rn(a.x, b.x);
                acc1 = __fmaf_rn(d, d, acc1);
                d = __fsub_rn(a.y, b.y);
                acc1 = __fmaf_rn(d, d, acc1);
                d = __fsub_rn(a.z, b.z);
                acc1 = __fmaf_rn(d, d, acc1);
                d = __fsub_rn(a.w, b.w);
                acc1 = __fmaf_rn(d, d, acc1);
            }
        }
    }
    const unsigned long long key = j >= 0 ? (((unsigned long long)__float_as_uint(acc) << 32) | (unsigned)j) : ~0ull;
    const unsigned long long key1 = j1 >= 0 ? (((unsigned long long)__float_as_uint(acc1) << 32) | (unsigned)j1) : ~0ull;
    int rank = 0, rank1 = 0;
#pragma unroll
    for (int o = 0; o < 32; ++o) {
        const unsigned long long a = __shfl_sync(kFull, key, o);
        rank += a < key ? 1 : 0;
        rank1 += a < key1 ? 1 : 0;
    }
    if (mfull > 32) {
#pragma unroll
        for (int o = 0; o < 32; ++o) {
            const unsigned long long b = __shfl_sync(kFull, key1, o);
            rank += b < key ? 1 : 0;
            rank1 += b < key1 ? 1 : 0;
        }
    }
    if (j >= 0 && rank >= skip && rank < skip + k) {
        idx[(size_t)q * k + rank - skip] = j;
        if (dist2) dist2[(size_t)q * k + rank - skip] = acc;
    }
    if (j1 >= 0 && rank1 >= skip && rank1 < skip + k) {
        idx[(size_t)q * k + rank1 - skip] = j1;
        if (dist2) dist2[(size_t)q * k + rank1 - skip] = acc1;
    }
}

// ---------------------------------------------------------------- flagged queries: exact brute force, warp = query
// Queries the filter could not bound (list overflow, duplicates / clusters, non-finite values) are few; a warp computes all n
// reference distances of its query once into shared memory and selects ranks skip..skip+k-1 by repeated minimum search over the
// (d2, index) keys.  Non-finite distances are never selected and missing ranks read index 0 / +inf, exactly like
// knn_feat_kernel's insertion list (knn_feat.cu), whose per-CTA cost (64 queries at a time) made one flagged query a 0.3 ms tail.
__device__ __noinline__ void kf_brute_query(const float* __restrict__ xT, int q, int c, int n, int k, int skip, int lane, float* __restrict__ dsm,
                               long long* __restrict__ idx, float* __restrict__ dist2) {
    const int bz = q / n;
    const float* xi = xT + (size_t)q * c;
    for (int t = lane; t < n; t += 32) {
        const float* xj = xT + ((size_t)bz * n + t) * c;
        float acc = 0.f;
        for (int ch = 0; ch < c; ch += 4) {                          // c % 8 == 0 on this path
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
        dsm[t] = acc;
    }
    __syncwarp();
    unsigned long long last = 0;
    for (int e = 0; e < k + skip; ++e) {
        unsigned long long best = ~0ull;
        for (int t = lane; t < n; t += 32) {
            const float d = dsm[t];
            const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)t;
            if (d < kInf && (e == 0 || key > last) && key < best) best = key;   // NaN / +inf are never selected
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(kFull, best, o);
            best = other < best ? other : best;
        }
        if (best == ~0ull) {                                         // fewer finite candidates than ranks
            for (int f = max(e, skip) + lane; f < k + skip; f += 32) {
                idx[(size_t)q * k + f - skip] = 0;
                if (dist2) dist2[(size_t)q * k + f - skip] = kInf;
            }
            break;
        }
        if (lane == 0 && e >= skip) {
            idx[(size_t)q * k + e - skip] = (long long)(unsigned)best;
            if (dist2) dist2[(size_t)q * k + e - skip] = __uint_as_float((unsigned)(best >> 32));
        }
        last = best;
    }
}

// ---------------------------------------------------------------- host side
static size_t kf_align(size_t v) { return (v + 255) & ~(size_t)255; }

int knn_feat_tc_brute_max() { return KF_BRUTE_MAX; }

bool knn_feat_tc_eligible(int c, int n, int k, int skip) {
    return c >= 8 && c <= 256 && (c & 7) == 0 && n >= 128 && n <= 4096 && (n & 127) == 0 && k + skip <= TF_KMAX && k >= 1;
}
size_t knn_feat_tc_workspace(int b, int c, int n) {
    const size_t bc = (size_t)b * c, bn = (size_t)b * n;
    const size_t cp = (size_t)((c + 31) / 32) * 32;                 // channels padded to whole K-blocks in the tiled copy
    return kf_align(bc * 4) + kf_align((size_t)b * cp * n * 4) + kf_align(bc * n * 4) + kf_align(bn * KF_SLABS * 4) + kf_align(bn * TF_CAP * 4) + kf_align(bn * 4) + 256 + 256;
}

// Runs prep + filter + re-rank; *flags receives the per-query count array (negative = recompute with the exact kernel).
int knn_feat_tc_launch(const float* x, int b, int c, int n, int k, int skip, long long* idx, float* dist2, void* ws, const int** flags,
                       const int** nflag_out, int* brute_n_max, cudaStream_t st) {
    const size_t bc = (size_t)b * c, bn = (size_t)b * n;
    unsigned char* p = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
    float* mean = reinterpret_cast<float*>(p);
    p += kf_align(bc * 4);
    float* xc = reinterpret_cast<float*>(p);
    p += kf_align((size_t)b * ((c + 31) / 32) * 32 * n * 4);
    float* xT = reinterpret_cast<float*>(p);
    p += kf_align(bc * n * 4);
    float* nrm = reinterpret_cast<float*>(p);                       // [b][KF_SLABS][n] partial squared norms
    p += kf_align(bn * KF_SLABS * 4);
    int* cand = reinterpret_cast<int*>(p);
    p += kf_align(bn * TF_CAP * 4);
    int* cnt = reinterpret_cast<int*>(p);
    p += kf_align(bn * 4);
    int* nflag = reinterpret_cast<int*>(p);
    kf_mean_kernel<<<(unsigned)((bc + 7) / 8), 256, 0, st>>>(x, (int)bc, n, mean, nflag);
    PDGN_CHECK_LAUNCH();
    const int cps = (((c + 31) / 32 + KF_SLABS - 1) / KF_SLABS) * 32;   // channels per slab
    kf_prep_kernel<<<dim3((n + 31) / 32, b, KF_SLABS), 256, 0, st>>>(x, mean, c, n, cps, xc, xT, nrm);
    PDGN_CHECK_LAUNCH();
    // accumulator tiles of 256 candidates when the cloud allows (half the tcgen05.mma count: the K = 8 instructions are issue /
    // operand-fetch bound, ~0.26 us each at N = 128), ring as deep as shared memory allows (>= 2 stages)
    const int tn = (n % 256 == 0 && n >= 1024) ? 256 : 128;   // small clouds: more, smaller tiles keep the ring deep
    static const char* tn_env = tune_env("PDGN_KNN_FEAT_TN");
    const int tnsel = (tn_env && atoi(tn_env) == 128) ? 128 : tn;
    const size_t fixed = (size_t)((c + 31) / 32) * 16384 + (size_t)n * 4 + (size_t)32 * TF_M * 4 + 2 * TF_M * 4, sb = (size_t)(tnsel / 128) * 16384;
    int nst = (int)((216 * 1024 - fixed) / sb);
    if (nst > TF_NST) nst = TF_NST;
    if (nst < 2) return PDGN_ERR_UNSUPPORTED;
    const size_t smem = fixed + (size_t)nst * sb;
    float* dbg = tune_env("PDGN_KNN_FEAT_DBG") ? reinterpret_cast<float*>(cand) : nullptr;
    uint32_t idesc = kf_idesc(tnsel);
    if (tune_env("PDGN_KNN_FEAT_NOMMA")) idesc |= 1u;
    if (tune_env("PDGN_KNN_FEAT_NOEPI")) idesc |= 2u;
    if (tnsel == 256) {
        PDGN_CUDA(cudaFuncSetAttribute(knn_feat_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn_feat_tc_kernel<256><<<dim3(n / TF_M, b), TF_T, smem, st>>>(xc, nrm, c, n, k + skip, cand, cnt, nflag, nst, dbg, idesc);
    } else {
        PDGN_CUDA(cudaFuncSetAttribute(knn_feat_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn_feat_tc_kernel<128><<<dim3(n / TF_M, b), TF_T, smem, st>>>(xc, nrm, c, n, k + skip, cand, cnt, nflag, nst, dbg, idesc);
    }
    PDGN_CHECK_LAUNCH();
    const size_t rr_smem = (size_t)8 * 33 * RR_RS * 4;
    PDGN_CUDA(cudaFuncSetAttribute(knn_feat_rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rr_smem));
    knn_feat_rerank_kernel<<<(unsigned)((bn + 7) / 8), 256, rr_smem, st>>>(xT, cand, cnt, nflag, c, n, k, skip, (int)bn, idx, dist2);
    PDGN_CHECK_LAUNCH();
    *flags = cnt;                                                    // the caller's exact kernel takes the flagged queries the
    *nflag_out = nflag;                                              //   re-rank kernel left (many of them, or a large cloud)
    *brute_n_max = 33 * RR_RS;
    return PDGN_OK;
}

}  // namespace pdgn
