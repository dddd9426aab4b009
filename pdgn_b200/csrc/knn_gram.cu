// knn_gram.cu -- xyz kNN for full-size clouds (256 < n <= 2048, k <= 20), bit-exact with the reference, round-2 design.
//
// Same contract as knn_select_kernel (knn_xyz.cu): idx / dist2 in ascending (d2, index) order, d2 from the reference's compiled
// FMUL/FFMA/FFMA chain (knnquery_cuda_kernel.cu:15-47), NaN / +inf distances never selected, missing neighbours idx 0 / +inf.
// What changed is where the instructions go (measured with tools/knn_probe.cu, profiles/r02_knn_probe.txt):
//
//   lane = 4 queries (register blocked), one CTA = 4 warps = 512 queries against one cloud staged in shared memory.
//   pass 1   FILTER ONLY, so it need not be the reference arithmetic: g = |p|^2 - 2 q.p as 3 FFMA on a precomputed |p|^2 plane,
//            minimum per subgroup of SS candidates (FMNMX3 trees), stored as bf16 rounded DOWN.  Subgroups are STRIDED
//            (candidate j -> subgroup j mod 128): an index-coherent cloud (generator output, scan order) then spreads a
//            query's neighbours over many subgroups exactly like a shuffled one, and the bound below stays tight.
//   bound    tau = k-th smallest of 64 group minima (group = subgroups g and g+64), by a lane-private bitonic network on
//            packed u16x2 values (VIMNMX.U16x2).  With the rigorous error margin E of the Gram form this yields U >= the k-th
//            nearest REFERENCE distance, and a flag threshold F such that every subgroup holding a candidate with d_ref <= U has
//            its stored minimum <= F.  (Derivation at kq_margin.)
//   pass 2   lane = query again: each lane walks its own flagged subgroups (~22 of 128), recomputes the EXACT reference
//            distance of their candidates and appends the survivors d_ref <= U (24 +- 2 for k = 20) as 64-bit (d2, index) keys
//            to a lane-private list; no ballots, no cross-lane traffic.
//   rank     lane-private bitonic sort of <= 32 keys.  Fast path: 32-bit keys (d2 bits with the low 5 bits replaced by the
//            list slot) -- exact unless two survivors share their upper 27 bits, which is detected after the sort and sends
//            the warp through the 64-bit network instead.  Survivor overflow (duplicates, clusters, fewer than k finite
//            candidates) goes to a warp-cooperative exact selection, so the result never depends on the heuristics.
#include "common.cuh"
#include "selnet.cuh"

namespace pdgn {

constexpr int KQ_NQ = 512;                      // queries per CTA = W warps x 32 lanes x R queries per lane
constexpr int KQ_NSUB = 128;                    // subgroups per query
constexpr int KQ_TILE = 2048;                   // candidate capacity
constexpr int KQ_PL = KQ_TILE + KQ_TILE / 32 * 4;   // plane length: 4 floats of padding per 32 (bank skew for the lane-private rescans)
constexpr int KQ_CAP = 46;                      // key slots per query
constexpr int KQ_SORT = 32;                     // keys the ranking networks take
// Bytes per (warp, r) block.  Pass 1 fills its first 8 KB with the sub-minimum words [64][32] u32.  The flag pass compacts the
// flagged-subgroup list IN PLACE over the words it has already read (byte i of a lane's list lives in word i/4 of its column:
// at most 2 entries per word read, so the list never overtakes the read position), and the 64-bit survivor keys then grow from
// the TOP of the block downwards (slot s = 256 (s+1) bytes below the end), so the two only meet for absurd survivor counts.
constexpr int KQ_BLK = KQ_CAP * 32 * 8;
constexpr int KQ_KMAX = 20;
static_assert(KQ_BLK >= 64 * 32 * 4, "sub-minimum words must fit the block");

__device__ __forceinline__ int kq_pad(int pos) { return pos + ((pos >> 5) << 2); }
// Error margin of the pass-1 value h = fl(g + |q|^2), g = fl(|p|^2 - 2 q.p) (3 FFMA on a 3-rounding |p|^2), against the
// reference's computed d_ref.  With u = 2^-24: |g - (|p|^2 - 2 q.p)| <= 3u|p|^2 + 3u(|p|^2 + 2|q||p|), |fl(|q|^2) - |q|^2| <= 3u|q|^2,
// the final add contributes u(|p|+|q|)^2, and |d_ref - |q-p|^2| <= 5.1u|q-p|^2: in total < 15.1u(|p|+|q|)^2.
// E = 1.2e-6 (|p|max + |q|)^2 > 16u(...)(1 + slack for the two square roots and the product).
__device__ __forceinline__ float kq_margin(float pmax2, float qq) {
    const float s = __fadd_ru(__fsqrt_ru(pmax2), __fsqrt_ru(qq));
    return __fmul_ru(1.2e-6f, __fmul_ru(s, s));
}

// k-th smallest (k <= 32) of the 64 group minima of one query.  words[e] (e < 64, stride 32 words) = (sub e) | (sub e+64) << 16.
__device__ __forceinline__ unsigned kq_kth_group_min(const unsigned* __restrict__ words, int k) {
    unsigned v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const unsigned a = words[i * 32], b = words[(i + 32) * 32];
        const unsigned ga = min(a & 0xffffu, a >> 16), gb = min(b & 0xffffu, b >> 16);
        v[i] = ga | ((gb ^ 0xffffu) << 16);        // upper half complemented: its ascending order is descending in gb
    }
    kq_bitonic_sort<32>([&](int i, int p, bool up) {
        const unsigned lo = kq_min2(v[i], v[p]), hi = kq_max2(v[i], v[p]);
        v[i] = up ? lo : hi;
        v[p] = up ? hi : lo;
    });
    // (ga ascending) ++ (gb descending) is bitonic: the element-wise minimum holds the 32 smallest of the 64, bitonic again
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = min(v[i] & 0xffffu, (v[i] >> 16) ^ 0xffffu);
    kq_bitonic_merge<32>([&](int i, int p, bool) {
        const unsigned lo = min(v[i], v[p]), hi = max(v[i], v[p]);
        v[i] = lo;
        v[p] = hi;
    });
    unsigned t = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) t = (i < k) ? max(t, v[i]) : t;   // v ascending: v[k-1] without a dynamically indexed array
    return t;
}

// Warp-cooperative exact selection for ONE query (slow path): every lane stages the (d2, index) keys of its 4*SS candidates in
// shared memory once (`stage`: the warp's own, by now idle, key blocks), then k rounds of "smallest key above the previous one".
template <int SS>
__device__ void kq_exact_query(const float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ Z, int n, int k,
                               float qx, float qy, float qz, int lane, unsigned long long* __restrict__ stage,
                               int* __restrict__ oi, float* __restrict__ od) {
    constexpr int PER = KQ_NSUB * SS / 32;
    // staged as 4-byte distances (+inf = not a candidate), the index is the position: 8 KB for a full tile, which also fits the
    // single 11.5 KB block a warp owns in the 16 x 1 shape
    float* dst = reinterpret_cast<float*>(stage);
#pragma unroll 4
    for (int i = 0; i < PER; ++i) {
        const int pos = lane + 32 * i;
        const int j = (pos % SS) * KQ_NSUB + pos / SS;
        const int pp = kq_pad(pos);
        const float d = d2_xyz(qx, qy, qz, X[pp], Y[pp], Z[pp]);
        const bool ok = j < n && d <= 3.402823466e+38f;                           // NaN / +inf are never selected
        dst[i * 32 + lane] = ok ? d : kInf;
    }
    unsigned long long last = 0;
    for (int e = 0; e < k; ++e) {
        unsigned long long best = ~0ull;
#pragma unroll 4
        for (int i = 0; i < PER; ++i) {
            const int pos = lane + 32 * i;
            const float d = dst[i * 32 + lane];
            const unsigned long long key = d < kInf ? (((unsigned long long)__float_as_uint(d) << 32) | (unsigned)((pos % SS) * KQ_NSUB + pos / SS)) : ~0ull;
            if ((e == 0 || key > last) && key < best) best = key;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(kFull, best, o);
            best = other < best ? other : best;
        }
        if (best == ~0ull) {                    // fewer than k finite candidates
            for (int f = e + lane; f < k; f += 32) {
                oi[f] = 0;
                if (od) od[f] = kInf;
            }
            return;
        }
        if (lane == 0) {
            oi[e] = (int)(unsigned)best;
            if (od) od[e] = __uint_as_float((unsigned)(best >> 32));
        }
        last = best;
    }
}

template <int SS, int KQ_W, int KQ_R>
__global__ void __launch_bounds__(KQ_W * 32, 1) knn_gram_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, int n,
                                                               int m, int k, int* __restrict__ idx, float* __restrict__ dist2) {
    constexpr int KQ_T = KQ_W * 32;
    static_assert(KQ_T * KQ_R == KQ_NQ, "512 queries per CTA");
    extern __shared__ __align__(16) unsigned char kq_smem[];
    float* X = reinterpret_cast<float*>(kq_smem);
    float* Y = X + KQ_PL;
    float* Z = Y + KQ_PL;
    float* P = Z + KQ_PL;
    unsigned char* blocks = reinterpret_cast<unsigned char*>(P + KQ_PL);
    __shared__ unsigned s_pmax2;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, bz = blockIdx.y;
    const float* pb = xyz + (size_t)bz * n * 3;
    const float nanv = __int_as_float(0x7fc00000);
    if (t == 0) s_pmax2 = 0u;

    // ---------------- tile.  Stage the cloud's 3n floats as they lie (16-byte loads, all in flight at once), then transpose:
    // candidate j -> position (j mod 128) * SS + j / 128 of four planes (x, y, z, |p|^2), NaN beyond n.
    {
        float* raw = reinterpret_cast<float*>(blocks);                 // the key blocks are idle until pass 1
        const int nfl = 3 * n;
        if ((reinterpret_cast<uintptr_t>(pb) & 15) == 0) {
            const int nv = nfl >> 2;
            const float4* src = reinterpret_cast<const float4*>(pb);
            float4* dst = reinterpret_cast<float4*>(raw);
#pragma unroll 4
            for (int i = t; i < nv; i += KQ_T) dst[i] = __ldg(src + i);
            if (t < nfl - 4 * nv) raw[4 * nv + t] = __ldg(pb + 4 * nv + t);
        } else {
#pragma unroll 8
            for (int i = t; i < nfl; i += KQ_T) raw[i] = __ldg(pb + i);
        }
        __syncthreads();
        float pmax = 0.f;
#pragma unroll 4
        for (int j = t; j < KQ_NSUB * SS; j += KQ_T) {
            float x = nanv, y = nanv, z = nanv, pp = nanv;
            if (j < n) {
                x = raw[3 * j]; y = raw[3 * j + 1]; z = raw[3 * j + 2];
                pp = __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x)));
                if (pp <= 3.402823466e+38f) pmax = fmaxf(pmax, pp);
                else pp = nanv;                 // non-finite / overflowing candidate: invisible to the filter; its exact distance is
            }                                   //   +inf or NaN, which the reference never selects either
            const int pos = kq_pad((j & (KQ_NSUB - 1)) * SS + (j >> 7));
            X[pos] = x; Y[pos] = y; Z[pos] = z; P[pos] = pp;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) pmax = fmaxf(pmax, __shfl_xor_sync(kFull, pmax, o));
        if (lane == 0) atomicMax(&s_pmax2, __float_as_uint(pmax));     // pmax >= 0: unsigned order == float order
    }

    // ---------------- queries: R per lane
    const int q0 = blockIdx.x * KQ_NQ + warp * (32 * KQ_R) + lane;     // query r of this lane = q0 + 32 r
    float qx[KQ_R], qy[KQ_R], qz[KQ_R], qq[KQ_R], ax[KQ_R], ay[KQ_R], az[KQ_R];
#pragma unroll
    for (int r = 0; r < KQ_R; ++r) {
        const int q = max(0, min(q0 + 32 * r, m - 1));
        const float* qp = new_xyz + ((size_t)bz * m + q) * 3;
        qx[r] = qp[0]; qy[r] = qp[1]; qz[r] = qp[2];
        qq[r] = __fmaf_rn(qz[r], qz[r], __fmaf_rn(qy[r], qy[r], __fmul_rn(qx[r], qx[r])));
        ax[r] = -2.f * qx[r]; ay[r] = -2.f * qy[r]; az[r] = -2.f * qz[r];
    }
    // sub-minimum words of absent subgroups (n < 128 * SS leaves none absent, but a ragged m leaves nothing to init either:
    // every subgroup index 0..127 is written by pass 1 below, absent candidates being NaN => minimum +inf => 0x7f80)
    __syncthreads();
    const float pmax2 = __uint_as_float(s_pmax2);

    // ---------------- pass 1: Gram-form subgroup minima (filter only)
    {
        unsigned short* sub = reinterpret_cast<unsigned short*>(blocks + (size_t)(warp * KQ_R) * KQ_BLK) + 2 * lane;
#pragma unroll 2
        for (int sg = 0; sg < KQ_NSUB; ++sg) {
            const int base = kq_pad(sg * SS);
            float mn[KQ_R];
#pragma unroll
            for (int r = 0; r < KQ_R; ++r) mn[r] = kInf;
#pragma unroll
            for (int qd = 0; qd < SS / 4; ++qd) {
                const float4 x4 = *reinterpret_cast<const float4*>(X + base + 4 * qd);
                const float4 y4 = *reinterpret_cast<const float4*>(Y + base + 4 * qd);
                const float4 z4 = *reinterpret_cast<const float4*>(Z + base + 4 * qd);
                const float4 p4 = *reinterpret_cast<const float4*>(P + base + 4 * qd);
                // coordinate-major order: consecutive FFMAs share the candidate operand (register reuse cache), which is what lets
                // a three-source FFMA issue near full rate (tools/knn_probe.cu: 90 % vs 80 % of the 6-instr roofline at R = 2)
                const float xs[4] = {x4.x, x4.y, x4.z, x4.w}, ys[4] = {y4.x, y4.y, y4.z, y4.w};
                const float zs[4] = {z4.x, z4.y, z4.z, z4.w}, ps[4] = {p4.x, p4.y, p4.z, p4.w};
                float g[KQ_R][4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int r = 0; r < KQ_R; ++r) g[r][u] = __fmaf_rn(az[r], zs[u], ps[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int r = 0; r < KQ_R; ++r) g[r][u] = __fmaf_rn(ay[r], ys[u], g[r][u]);
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int r = 0; r < KQ_R; ++r) g[r][u] = __fmaf_rn(ax[r], xs[u], g[r][u]);
#pragma unroll
                for (int r = 0; r < KQ_R; ++r) mn[r] = min3(min3(mn[r], g[r][0], g[r][1]), g[r][2], g[r][3]);   // fminf drops NaN: padding never wins
            }
#pragma unroll
            for (int r = 0; r < KQ_R; ++r) {
                const float h = fmaxf(__fadd_rn(mn[r], qq[r]), 0.f);   // >= 0 (and 0 for a NaN query): bf16 bits order like integers
                // word (sg & 63) of block r, half (sg >> 6); truncation = rounding down
                sub[(size_t)r * (KQ_BLK / 2) + (sg & 63) * 64 + (sg >> 6)] = (unsigned short)(__float_as_uint(h) >> 16);
            }
        }
    }
    __syncwarp();

    // ---------------- per query: bound, flags, exact rescan, rank
    unsigned slow_overflow = 0;                 // bit r: query r of this lane needs the cooperative exact selection
#pragma unroll 1
    for (int r = 0; r < KQ_R; ++r) {
        unsigned char* blk = blocks + (size_t)(warp * KQ_R + r) * KQ_BLK;
        const unsigned* words = reinterpret_cast<const unsigned*>(blk) + lane;
        const float qxr = qx[r], qyr = qy[r], qzr = qz[r];
        // -- bound
        const unsigned t16 = kq_kth_group_min(words, k);
        float uf = 3.402823466e+38f;
        unsigned f16 = 0x7f80u;
        if (t16 < 0x7f80u) {
            const float e = kq_margin(pmax2, qq[r]);
            const float u = __fadd_ru(__uint_as_float((t16 + 1u) << 16), e);      // k subgroups hold a candidate with d_ref <= u
            const float fv = __fadd_ru(u, e);
            if (fv < kInf) {                                                   // (false for NaN too)
                uf = fminf(u, 3.402823466e+38f);
                f16 = __float_as_uint(fv) >> 16;
            }
        }
        // -- flags: list (one byte each, lane-private) of the subgroups whose stored minimum is <= f16.  Absent / all-NaN
        //    subgroups hold 0x7f80: harmless when flagged, their candidates fail the exact test.
        unsigned char* plist = blk + 4 * lane;  // entry i of this lane: plist[(i >> 2) * 128 + (i & 3)]
        auto pl = [&](int i) -> unsigned char& { return plist[(i >> 2) * 128 + (i & 3)]; };
        int nf = 0;
#pragma unroll
        for (int e = 0; e < 64; ++e) {
            const unsigned w = words[e * 32];
            const bool a = (w & 0xffffu) <= f16, b = (w >> 16) <= f16;       // predicated stores, no branches
            if (a) pl(nf) = (unsigned char)e;
            nf += a ? 1 : 0;
            if (b) pl(nf) = (unsigned char)(64 + e);
            nf += b ? 1 : 0;
        }
        bool over = false;
        __syncwarp();                           // every lane has read its words: the rest of the block becomes the key list
        // -- exact rescan of the flagged subgroups.  Software pipelined: the candidates of the next subgroup are loaded before
        //    the survivors of the current one are stored (the compiler cannot move shared loads above shared stores itself).
        unsigned long long* const ktop = reinterpret_cast<unsigned long long*>(blk + KQ_BLK) + lane;
        auto key_at = [&](int slot) -> unsigned long long& { return ktop[-32 * (slot + 1)]; };
        unsigned long long* kp = ktop;          // one slot above the next free one
        const int nfmax = __reduce_max_sync(kFull, nf);
        // lowest slot address that stays clear of EVERY lane's flag list (a 256-byte key row spans all 32 lanes' columns of
        // two 128-byte list rows, so the floor is the warp's longest list), plus the 4 slots a quad may add before the check
        const unsigned long long* const kfloor = reinterpret_cast<const unsigned long long*>(blk + ((nfmax + 3) >> 2) * 128) + 4 * 32;
        float4 cx[SS / 4], cy[SS / 4], cz[SS / 4];
        int sg = nf > 0 ? (int)pl(0) : -1;
        {
            const int base = kq_pad(max(sg, 0) * SS);
#pragma unroll
            for (int qd = 0; qd < SS / 4; ++qd) {
                cx[qd] = *reinterpret_cast<const float4*>(X + base + 4 * qd);
                cy[qd] = *reinterpret_cast<const float4*>(Y + base + 4 * qd);
                cz[qd] = *reinterpret_cast<const float4*>(Z + base + 4 * qd);
            }
        }
        for (int it = 0; it < nfmax; ++it) {
            const int sgn = it + 1 < nf ? (int)pl(it + 1) : -1;
            float d[SS];
#pragma unroll
            for (int qd = 0; qd < SS / 4; ++qd) {
                d[4 * qd] = d2_xyz(qxr, qyr, qzr, cx[qd].x, cy[qd].x, cz[qd].x);
                d[4 * qd + 1] = d2_xyz(qxr, qyr, qzr, cx[qd].y, cy[qd].y, cz[qd].y);
                d[4 * qd + 2] = d2_xyz(qxr, qyr, qzr, cx[qd].z, cy[qd].z, cz[qd].z);
                d[4 * qd + 3] = d2_xyz(qxr, qyr, qzr, cx[qd].w, cy[qd].w, cz[qd].w);
            }
            {
                const int base = kq_pad(max(sgn, 0) * SS);
#pragma unroll
                for (int qd = 0; qd < SS / 4; ++qd) {
                    cx[qd] = *reinterpret_cast<const float4*>(X + base + 4 * qd);
                    cy[qd] = *reinterpret_cast<const float4*>(Y + base + 4 * qd);
                    cz[qd] = *reinterpret_cast<const float4*>(Z + base + 4 * qd);
                }
            }
            const unsigned sgc = (unsigned)max(sg, 0);
#pragma unroll
            for (int qd = 0; qd < SS / 4; ++qd) {
                over = over || kp < kfloor;                            // a quad may add 4 keys: never write into the flag list
                const float ul = (sg >= 0 && !over) ? uf : -1.0f;      // idle lanes / overflowed lists keep nothing
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (d[4 * qd + u] <= ul) {                         // candidate of slot (qd, u): sg + 128 (4 qd + u)
                        kp -= 32;
                        *kp = ((unsigned long long)__float_as_uint(d[4 * qd + u]) << 32) | (sgc + KQ_NSUB * (4 * qd + u));
                    }
                }
            }
            sg = sgn;
        }
        int cnt = (int)(ktop - kp) >> 5;
        // more survivors than the ranking network takes (rare: 24 +- 2 expected for k = 20): drop the largest keys, one at a time
        while (__any_sync(kFull, !over && cnt > KQ_SORT)) {
            if (!over && cnt > KQ_SORT) {
                unsigned long long big = key_at(0);
                int at = 0;
                for (int s = 1; s < cnt; ++s) {
                    const unsigned long long ks = key_at(s);
                    if (ks > big) { big = ks; at = s; }
                }
                --cnt;
                key_at(at) = key_at(cnt);
            }
        }
        if (over) slow_overflow |= 1u << r;
        // -- rank (lane-private).  Fast path: 32-bit keys = d2 bits with the low 5 bits replaced by the slot number
        const int q = q0 + 32 * r;
        const bool live = q < m && !over;
        int* oi = idx + ((size_t)bz * m + min(q, m - 1)) * k;
        float* od = dist2 ? dist2 + ((size_t)bz * m + min(q, m - 1)) * k : nullptr;
        unsigned mk[KQ_SORT];
#pragma unroll
        for (int s = 0; s < KQ_SORT; ++s) {
            const unsigned hi = (unsigned)(key_at(s) >> 32);
            // unused slots: distinct values above every finite d2, so they neither win nor look like ties
            mk[s] = (s < cnt) ? ((hi & ~31u) | (unsigned)s) : (0xfffffc00u | (unsigned)(s << 5) | (unsigned)s);
        }
        kq_bitonic_sort<KQ_SORT>([&](int i, int p, bool up) {
            const unsigned lo = min(mk[i], mk[p]), hi = max(mk[i], mk[p]);
            mk[i] = up ? lo : hi;
            mk[p] = up ? hi : lo;
        });
        bool tie = false;
#pragma unroll
        for (int s = 0; s + 1 < KQ_SORT; ++s) tie = tie || ((mk[s] ^ mk[s + 1]) < 32u);
        tie = tie && !over && cnt > 1;
        if (__any_sync(kFull, tie)) {
            // some lane has two survivors whose d2 agree in the upper 27 bits: order by the full 64-bit keys instead
            unsigned long long kk[KQ_SORT];
#pragma unroll
            for (int s = 0; s < KQ_SORT; ++s) kk[s] = (s < cnt) ? key_at(s) : (~0ull - (unsigned)(KQ_SORT - s));
            kq_bitonic_sort<KQ_SORT>([&](int i, int p, bool up) {
                const bool sw = (kk[p] < kk[i]) == up;
                const unsigned long long a = sw ? kk[p] : kk[i], b = sw ? kk[i] : kk[p];
                kk[i] = a;
                kk[p] = b;
            });
            if (live) {
#pragma unroll
                for (int e = 0; e < KQ_KMAX; ++e) {
                    if (e < k) {
                        const bool have = e < cnt;
                        oi[e] = have ? (int)(unsigned)kk[e] : 0;
                        if (od) od[e] = have ? __uint_as_float((unsigned)(kk[e] >> 32)) : kInf;
                    }
                }
            }
        } else if (live) {
#pragma unroll
            for (int e = 0; e < KQ_KMAX; ++e) {
                if (e < k) {
                    const bool have = e < cnt;
                    const unsigned long long key = key_at((int)(mk[e] & 31u));
                    oi[e] = have ? (int)(unsigned)key : 0;
                    if (od) od[e] = have ? __uint_as_float((unsigned)(key >> 32)) : kInf;
                }
            }
        }
        __syncwarp();
    }

    // ---------------- survivor overflow: exact cooperative selection, one query at a time
#pragma unroll 1
    for (int r = 0; r < KQ_R; ++r) {
        unsigned todo = __ballot_sync(kFull, (slow_overflow >> r) & 1u);
        while (todo) {
            const int src = __ffs((int)todo) - 1;
            todo &= todo - 1;
            const float sx = __shfl_sync(kFull, qx[r], src), sy = __shfl_sync(kFull, qy[r], src), sz = __shfl_sync(kFull, qz[r], src);
            const int q = q0 - lane + src + 32 * r;
            if (q < m)
                kq_exact_query<SS>(X, Y, Z, n, k, sx, sy, sz, lane,
                                   reinterpret_cast<unsigned long long*>(blocks + (size_t)(warp * KQ_R) * KQ_BLK),
                                   idx + ((size_t)bz * m + q) * k, dist2 ? dist2 + ((size_t)bz * m + q) * k : nullptr);
            __syncwarp();
        }
    }
}

template <int SS, int W, int R>
static int kq_launch(const float* xyz, const float* new_xyz, int b, int n, int m, int k, int* idx, float* dist2, cudaStream_t st) {
    const size_t smem = (size_t)4 * KQ_PL * sizeof(float) + (size_t)W * R * KQ_BLK;
    PDGN_CUDA(cudaFuncSetAttribute(knn_gram_kernel<SS, W, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((m + KQ_NQ - 1) / KQ_NQ, b);
    knn_gram_kernel<SS, W, R><<<grid, W * 32, smem, st>>>(xyz, new_xyz, n, m, k, idx, dist2);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}

// Shapes this kernel takes: one resident tile, at least 2 candidates per subgroup, k within the 32-key ranking network's
// comfortable range (expected survivors for k = 20: 23.6 +- 2.1).
bool knn_gram_eligible(int n, int k) { return n > 256 && n <= KQ_TILE && k >= 1 && k <= KQ_KMAX; }

template <int W, int R>
static int kq_launch_ss(const float* xyz, const float* new_xyz, int b, int n, int m, int k, int* idx, float* dist2, cudaStream_t st) {
    if (n > 1024) return kq_launch<16, W, R>(xyz, new_xyz, b, n, m, k, idx, dist2, st);
    if (n > 512) return kq_launch<8, W, R>(xyz, new_xyz, b, n, m, k, idx, dist2, st);
    return kq_launch<4, W, R>(xyz, new_xyz, b, n, m, k, idx, dist2, st);
}

int knn_gram_launch(const float* xyz, const float* new_xyz, int b, int n, int m, int k, int* idx, float* dist2, cudaStream_t st) {
    static const char* shape = tune_env("PDGN_KNN_GRAM_SHAPE");   // tuning hook: "4x4" = 4 warps x 4 queries per lane instead of 8 x 2
    if (shape && shape[0] == '4') return kq_launch_ss<4, 4>(xyz, new_xyz, b, n, m, k, idx, dist2, st);
    if (shape && shape[0] == '1') return kq_launch_ss<16, 1>(xyz, new_xyz, b, n, m, k, idx, dist2, st);   // 16 warps x 1 query per lane
    return kq_launch_ss<8, 2>(xyz, new_xyz, b, n, m, k, idx, dist2, st);
}

}  // namespace pdgn
