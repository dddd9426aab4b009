// emd.cu -- all-pairs approximate Earth Mover's Distance (the other half of _pairwise_EMD_CD_).
//
// Replaces, for evaluation, ApproxMatch + MatchCost of the reference
// (evaluation/pytorch_structural_losses/src/approxmatch.cu:3-182 and :184-224, driven per batch of `batch_size`
// expanded pairs from evaluation/evaluation_metrics.py:26-31,110): per cloud pair the reference runs 9 annealing
// levels x 3 sweeps of exp(level*d2) over all n*m point pairs, RMW-ing a dense n x m match matrix in HBM (16.8 MB at
// 2048^2) nine times and sweeping it once more for the cost.  Here one CTA keeps BOTH clouds and the four n/m-sized
// state vectors on chip for the whole pair (64 KB of float4 in shared memory, remainders in registers); the match
// matrix is never materialised because the cost is linear in the per-level weights: cost = sum_levels sum_kl w_kl*|p_k-q_l|.
// HBM traffic per cloud pair: 49 KB in, 4 bytes out.
//
// Bound: the MUFU pipe (one ex2 per point pair in the second sweep, one ex2 + one sqrt in the fused transport pass):
// 27 sweeps x 2048^2 pairs x ~10 issue slots.  The arithmetic follows the reference statement by statement
// (same d2 chain, level*d2 then __expf, same update formulas); sums are taken in a different order, so parity is by
// tolerance (tests: 2e-4 relative against the recompiled reference kernels and the CPU oracle).
#include "common.cuh"

namespace pdgn {

// exp(level*d2) as ONE multiply + ONE MUFU.EX2: ex2.approx.ftz(d2 * (level*log2(e))).  The reference's __expf(level*d2)
// is the same MUFU with two roundings in the argument and denormal fix-up code around it; flushing results below
// 1.2e-38 to zero changes nothing measurable (tests: 2e-4 relative against the reference kernels).
__device__ __forceinline__ float exp2_fast(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float mul_ftz(float a, float b) {
    float r;
    asm("mul.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float sqrt_fast(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

constexpr int EM_T = 512;             // threads per CTA
constexpr int EM_R = 4;               // points of each cloud owned per thread (8 warps/SMSP keep the MUFU queue fed)
constexpr int EM_MAX = EM_T * EM_R;   // 2048 points per cloud
constexpr int EM_CH = 32;             // points per bounding-sphere chunk
constexpr int EM_NCH = EM_MAX / EM_CH;
constexpr int EM_WPTS = EM_R * 32;    // points owned by one warp: 4 consecutive chunks of the kd-ordered cloud

// ---- pre-pass: kd-tree order ------------------------------------------------------------------------------------
// exp(level*d2) underflows to exactly +0 for most point pairs at the first annealing levels (level = -4^7 .. -4^4: only
// pairs closer than 0.07 .. 0.6 contribute).  With the points of every cloud in kd-tree leaf order (recursive median split
// along the widest axis down to 32-point leaves), 32 consecutive points form a compact chunk and the 128 points a warp owns
// a compact patch, so whole (patch, chunk) blocks can be skipped by a bounding-box distance test without changing a single
// bit of the result (the skipped terms are exact zeros).
__device__ __forceinline__ unsigned ord_u32(float f) {  // monotone float -> unsigned
    const unsigned b = __float_as_uint(f);
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float ord_f32(unsigned u) {
    return __uint_as_float(u ^ ((u >> 31) ? 0x80000000u : 0xffffffffu));
}

constexpr int EM_SORT_T = 1024;

__global__ void __launch_bounds__(EM_SORT_T) emd_sort_kernel(const float* __restrict__ src, int cloud0, int n, float* __restrict__ dst) {
    __shared__ unsigned long long key[EM_MAX];   // (ordered coordinate << 11) | point index; ~0 = padding (sorts last)
    __shared__ float pts[3][EM_MAX];
    __shared__ unsigned box[EM_MAX / 64][6];     // per segment: min x,y,z then max x,y,z (ordered-unsigned)
    const int t = threadIdx.x, lane = t & 31;
    const float* P = src + (size_t)(cloud0 + blockIdx.x) * n * 3;
    int np2 = 64;
    while (np2 < n) np2 <<= 1;
    for (int i = t; i < np2; i += EM_SORT_T) {
        key[i] = i < n ? (unsigned long long)i : ~0ull;
        if (i < n) { pts[0][i] = P[i * 3]; pts[1][i] = P[i * 3 + 1]; pts[2][i] = P[i * 3 + 2]; }
    }
    __syncthreads();
    for (int W = np2; W >= 64; W >>= 1) {
        const int nseg = np2 / W;
        if (t < nseg * 6) box[t / 6][t % 6] = (t % 6) < 3 ? 0xffffffffu : 0u;
        __syncthreads();
        for (int i = t; i < np2; i += EM_SORT_T) {  // a warp's 32 consecutive ranks lie in one segment (W >= 64)
            const unsigned long long k = key[i];
            const bool valid = k != ~0ull;
            const int id = (int)(k & 2047);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const unsigned o = valid ? ord_u32(pts[c][id]) : 0u;
                const unsigned lo = __reduce_min_sync(kFull, valid ? o : 0xffffffffu), hi = __reduce_max_sync(kFull, o);
                if (lane == 0) { atomicMin(&box[i / W][c], lo); atomicMax(&box[i / W][3 + c], hi); }
            }
        }
        __syncthreads();
        for (int i = t; i < np2; i += EM_SORT_T) {
            const unsigned long long k = key[i];
            if (k == ~0ull) continue;
            const unsigned* bx = box[i / W];
            const float ex = ord_f32(bx[3]) - ord_f32(bx[0]), ey = ord_f32(bx[4]) - ord_f32(bx[1]), ez = ord_f32(bx[5]) - ord_f32(bx[2]);
            const int axis = (ex >= ey && ex >= ez) ? 0 : (ey >= ez ? 1 : 2);
            const int id = (int)(k & 2047);
            key[i] = ((unsigned long long)ord_u32(pts[axis][id]) << 11) | (unsigned)id;
        }
        __syncthreads();
        // bitonic sort of every W-block, all ascending
        for (int size = 2; size <= W; size <<= 1)
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int i = t; i < np2; i += EM_SORT_T) {
                    const int j = i ^ stride;
                    if (j > i) {
                        const unsigned long long a = key[i], b = key[j];
                        const bool up = size == W || (i & size) == 0;
                        if ((a > b) == up) { key[i] = b; key[j] = a; }
                    }
                }
                __syncthreads();
            }
    }
    float* D = dst + (size_t)blockIdx.x * n * 3;
    for (int i = t; i < n; i += EM_SORT_T) {
        const int id = (int)(key[i] & 2047);
        D[i * 3] = pts[0][id];
        D[i * 3 + 1] = pts[1][id];
        D[i * 3 + 2] = pts[2][id];
    }
}

struct EmBox { float lx, ly, lz, hx, hy, hz; };

// Axis-aligned bounding box of the valid points held one per lane (empty: lo = +inf, hi = -inf, which tests as "apart").
__device__ __forceinline__ EmBox warp_box(float x, float y, float z, bool valid) {
    EmBox b{valid ? x : kInf, valid ? y : kInf, valid ? z : kInf, valid ? x : -kInf, valid ? y : -kInf, valid ? z : -kInf};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        b.lx = fminf(b.lx, __shfl_xor_sync(kFull, b.lx, o)); b.hx = fmaxf(b.hx, __shfl_xor_sync(kFull, b.hx, o));
        b.ly = fminf(b.ly, __shfl_xor_sync(kFull, b.ly, o)); b.hy = fmaxf(b.hy, __shfl_xor_sync(kFull, b.hy, o));
        b.lz = fminf(b.lz, __shfl_xor_sync(kFull, b.lz, o)); b.hz = fmaxf(b.hz, __shfl_xor_sync(kFull, b.hz, o));
    }
    return b;
}
__device__ __forceinline__ void box_join(EmBox& a, const EmBox& b) {
    a.lx = fminf(a.lx, b.lx); a.ly = fminf(a.ly, b.ly); a.lz = fminf(a.lz, b.lz);
    a.hx = fmaxf(a.hx, b.hx); a.hy = fmaxf(a.hy, b.hy); a.hz = fmaxf(a.hz, b.hz);
}
// squared distance between two boxes stored as (lo, hi) float4 pairs
__device__ __forceinline__ float box_gap2(const float4* a, const float4* b) {
    const float4 alo = a[0], ahi = a[1], blo = b[0], bhi = b[1];
    const float dx = fmaxf(0.f, fmaxf(alo.x - bhi.x, blo.x - ahi.x));
    const float dy = fmaxf(0.f, fmaxf(alo.y - bhi.y, blo.y - ahi.y));
    const float dz = fmaxf(0.f, fmaxf(alo.z - bhi.z, blo.z - ahi.z));
    return dx * dx + dy * dy + dz * dz;
}

// One pass of a warp's own left points (4 rows per lane) over every right point: the transport sweep of level `l2a`
//   w = exp(l2a*d2) * ratioL[k] * ratioR[l];  acc3[k] += w;  cost += w * |p_k - q_l|
// fused (NEXT) with the first sweep of the NEXT level `l2b`, which walks the same point pairs
//   acc1[k] += exp(l2b*d2) * remainR[l]
// so that d2 is computed once and one sweep (and two CTA barriers) per level disappears.  Blocks of 32 right points whose
// bounding box is farther from the warp's box than `reach2` (that of the coarser of the two levels) hold exact zeros only.
template <bool NEXT>
__device__ __forceinline__ void emd_transport_pass(const float4* __restrict__ Rr, const float* __restrict__ RemR, const float4* Wb,
                                                   const float4* Rc, int nchR, int m, float reach2, float l2a, float l2b,
                                                   const float (&ox)[EM_R], const float (&oy)[EM_R], const float (&oz)[EM_R],
                                                   const float (&ratL)[EM_R], float (&acc3)[EM_R], float (&acc1)[EM_R],
                                                   float& cost) {
    for (int c = 0; c < nchR; ++c) {
        if (box_gap2(Wb, Rc + 2 * c) > reach2) continue;
        const int le = min(m, (c + 1) * EM_CH);
#pragma unroll 2
        for (int l = c * EM_CH; l < le; ++l) {
            const float4 q = Rr[l];
            const float rr = NEXT ? RemR[l] : 0.f;
#pragma unroll
            for (int i = 0; i < EM_R; ++i) {
                const float d2 = d2_xyz(q.x, q.y, q.z, ox[i], oy[i], oz[i]);
                float e3;
                if (NEXT) {
                    // the levels are a factor 4 apart (l2a == 4*l2b exactly), so exp(l2a*d2) = exp(l2b*d2)^4: two multiplies
                    // instead of a second MUFU.EX2 (the kernel is MUFU bound); mul.ftz keeps the underflow-to-zero of ex2.ftz
                    const float e1 = exp2_fast(l2b * d2);
                    acc1[i] = __fmaf_rn(e1, rr, acc1[i]);
                    const float e2 = mul_ftz(e1, e1);
                    e3 = mul_ftz(e2, e2);
                } else {
                    e3 = exp2_fast(l2a * d2);
                }
                const float w = e3 * ratL[i] * q.w;
                acc3[i] += w;
                cost = __fmaf_rn(w, sqrt_fast(d2), cost);
            }
        }
    }
}

__global__ void __launch_bounds__(EM_T, 2)
emd_allpairs_kernel(const float* __restrict__ A, const float* __restrict__ B, int ncols, int n, int m, int rstrip,
                    float* __restrict__ out, long long ld_out, int paired) {
    extern __shared__ __align__(16) float4 em_sm[];
    float4* L = em_sm;                 // [EM_MAX] left cloud:  x, y, z, ratioL
    float4* Rr = em_sm + EM_MAX;       // [EM_MAX] right cloud: x, y, z, ratioR
    float4* Lc = em_sm + 2 * EM_MAX;   // [2*EM_NCH] chunk boxes of the left cloud: (lo, hi) pairs
    float4* Rc = Lc + 2 * EM_NCH;      // [2*EM_NCH] chunk boxes of the right cloud
    float4* Wb = Rc + 2 * EM_NCH + 4 * (threadIdx.x >> 5);  // this warp's own boxes: left (lo, hi), right (lo, hi)
    float* RemR = reinterpret_cast<float*>(Rc + 2 * EM_NCH + 4 * (EM_T / 32));  // [EM_MAX] remainR of the right cloud
    __shared__ float red[EM_T / 32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int s = blockIdx.y;
    // paired mode (match_cost.py:6-44 on a batch): row s meets column s only, out[s] (ld_out == 0)
    const int r_begin = paired ? s : blockIdx.x * rstrip, r_end = paired ? s + 1 : min(ncols, r_begin + rstrip);
    const float multiL = n >= m ? 1.f : (float)(m / n), multiR = n >= m ? (float)(n / m) : 1.f;
    const int nchL = (n + EM_CH - 1) / EM_CH, nchR = (m + EM_CH - 1) / EM_CH;
    const int own0 = warp * EM_WPTS + lane;   // this thread owns points own0 + 32*i of each cloud: chunks 4*warp .. 4*warp+3

    const float* Ap = A + (size_t)s * n * 3;
    EmBox ownL{kInf, kInf, kInf, -kInf, -kInf, -kInf};  // box of the warp's own left points
    float ox[EM_R], oy[EM_R], oz[EM_R];                 // own left points (registers for the whole strip)
#pragma unroll
    for (int i = 0; i < EM_R; ++i) {
        const int p = own0 + 32 * i;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p < n) { v.x = Ap[p * 3]; v.y = Ap[p * 3 + 1]; v.z = Ap[p * 3 + 2]; }
        L[p] = v;
        ox[i] = v.x; oy[i] = v.y; oz[i] = v.z;
        const EmBox bx = warp_box(v.x, v.y, v.z, p < n);
        if (lane == 0) {
            Lc[2 * (warp * EM_R + i)] = make_float4(bx.lx, bx.ly, bx.lz, 0.f);
            Lc[2 * (warp * EM_R + i) + 1] = make_float4(bx.hx, bx.hy, bx.hz, 0.f);
        }
        box_join(ownL, bx);
    }
    if (lane == 0) { Wb[0] = make_float4(ownL.lx, ownL.ly, ownL.lz, 0.f); Wb[1] = make_float4(ownL.hx, ownL.hy, ownL.hz, 0.f); }
    // ex2.approx.ftz(l2*d2) is exactly +0 once l2*d2 <= -127; blocks whose boxes are farther apart than sqrt(reach2) (one
    // more unit of margin) contribute exact zeros and are skipped
    auto reach2_of = [](float l2) { return (128.0f / -l2) * 1.0002f; };
    constexpr float kLog2e = 1.4426950408889634f;
    for (int r = r_begin; r < r_end; ++r) {
        const float* Bp = B + (size_t)r * m * 3;
        __syncthreads();  // previous pair fully consumed
        float remL[EM_R], remR[EM_R], ratL[EM_R];
        EmBox ownR{kInf, kInf, kInf, -kInf, -kInf, -kInf};
#pragma unroll
        for (int i = 0; i < EM_R; ++i) {
            const int p = own0 + 32 * i;
            remL[i] = p < n ? multiL : 0.f;
            remR[i] = p < m ? multiR : 0.f;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p < m) { v.x = Bp[p * 3]; v.y = Bp[p * 3 + 1]; v.z = Bp[p * 3 + 2]; }
            Rr[p] = v;
            RemR[p] = remR[i];
            const EmBox bx = warp_box(v.x, v.y, v.z, p < m);
            if (lane == 0) {
                Rc[2 * (warp * EM_R + i)] = make_float4(bx.lx, bx.ly, bx.lz, 0.f);
                Rc[2 * (warp * EM_R + i) + 1] = make_float4(bx.hx, bx.hy, bx.hz, 0.f);
            }
            box_join(ownR, bx);
        }
        if (lane == 0) { Wb[2] = make_float4(ownR.lx, ownR.ly, ownR.lz, 0.f); Wb[3] = make_float4(ownR.hx, ownR.hy, ownR.hz, 0.f); }
        __syncthreads();
        float cost = 0.f;
        float level = -16384.f;  // -4^7, then /4 per level down to -4^-1
        // ---- first sweep of the first level: ratioL[k] = remainL[k] / (1e-9 + sum_l exp(level*d2) * remainR[l])
        {
            const float l2 = level * kLog2e, reach2 = reach2_of(l2);
            float acc[EM_R];
#pragma unroll
            for (int i = 0; i < EM_R; ++i) acc[i] = 1e-9f;
            for (int c = 0; c < nchR; ++c) {
                if (box_gap2(Wb, Rc + 2 * c) > reach2) continue;
                const int le = min(m, (c + 1) * EM_CH);
#pragma unroll 2
                for (int l = c * EM_CH; l < le; ++l) {
                    const float4 q = Rr[l];
                    const float rr = RemR[l];
#pragma unroll
                    for (int i = 0; i < EM_R; ++i)
                        acc[i] = __fmaf_rn(exp2_fast(l2 * d2_xyz(q.x, q.y, q.z, ox[i], oy[i], oz[i])), rr, acc[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < EM_R; ++i) {
                ratL[i] = remL[i] / acc[i];
                L[own0 + 32 * i].w = ratL[i];
            }
        }
        for (int j = 7; j > -2; --j, level *= 0.25f) {
            const float l2 = level * kLog2e, reach2 = reach2_of(l2);
            __syncthreads();  // ratioL of this level is in L[].w
            // ---- second sweep: sumr[l] = remainR[l] * sum_k exp(level*d2) * ratioL[k]; consumption; ratioR; remainR
            {
                float px[EM_R], py[EM_R], pz[EM_R], acc[EM_R];
#pragma unroll
                for (int i = 0; i < EM_R; ++i) {
                    const float4 v = Rr[own0 + 32 * i];
                    px[i] = v.x; py[i] = v.y; pz[i] = v.z;
                    acc[i] = 0.f;
                }
                for (int c = 0; c < nchL; ++c) {
                    if (box_gap2(Wb + 2, Lc + 2 * c) > reach2) continue;
                    const int ke = min(n, (c + 1) * EM_CH);
#pragma unroll 2
                    for (int k = c * EM_CH; k < ke; ++k) {
                        const float4 p = L[k];
#pragma unroll
                        for (int i = 0; i < EM_R; ++i)
                            acc[i] = __fmaf_rn(exp2_fast(l2 * d2_xyz(px[i], py[i], pz[i], p.x, p.y, p.z)), p.w, acc[i]);
                    }
                }
                // (own entries only: nobody else reads ratioR / remainR before the barrier below)
#pragma unroll
                for (int i = 0; i < EM_R; ++i) {
                    const float sumr = acc[i] * remR[i];
                    const float consumption = fminf(remR[i] / (sumr + 1e-9f), 1.0f);
                    Rr[own0 + 32 * i].w = consumption * remR[i];  // ratioR
                    remR[i] = fmaxf(0.0f, remR[i] - sumr);
                    RemR[own0 + 32 * i] = remR[i];
                }
            }
            __syncthreads();  // ratioR and the new remainR are in shared memory
            // ---- transport sweep of this level fused with the first sweep of the next one
            float acc3[EM_R], acc1[EM_R];
#pragma unroll
            for (int i = 0; i < EM_R; ++i) { acc3[i] = 0.f; acc1[i] = 1e-9f; }
            if (j > -1) {
                const float l2n = level * 0.25f * kLog2e;
                emd_transport_pass<true>(Rr, RemR, Wb, Rc, nchR, m, reach2_of(l2n), l2, l2n, ox, oy, oz, ratL, acc3, acc1, cost);
            } else {
                emd_transport_pass<false>(Rr, RemR, Wb, Rc, nchR, m, reach2, l2, 0.f, ox, oy, oz, ratL, acc3, acc1, cost);
            }
#pragma unroll
            for (int i = 0; i < EM_R; ++i) {
                remL[i] = fmaxf(0.0f, remL[i] - acc3[i]);
                ratL[i] = remL[i] / acc1[i];
                L[own0 + 32 * i].w = ratL[i];  // read by the next level's second sweep (after its barrier)
            }
        }
        cost = warp_sum(cost);
        if (lane == 0) red[warp] = cost;
        __syncthreads();
        if (t == 0) {
            float tot = 0.f;
#pragma unroll
            for (int w = 0; w < EM_T / 32; ++w) tot += red[w];
            out[(size_t)s * ld_out + r] = tot / (float)n;  // emd_approx: match_cost / N (evaluation_metrics.py:29-30)
        }
    }
}

}  // namespace pdgn

using namespace pdgn;

extern "C" size_t pdgn_emd_allpairs_workspace(int na, int nb, int n, int m) {
    if (na < 0 || nb < 0 || n < 0 || m < 0) return 0;
    return ((size_t)na * n + (size_t)nb * m) * 3 * sizeof(float) + 256;
}

extern "C" int pdgn_emd_allpairs(const float* A, const float* B, int na, int nb, int n, int m, int row0, int row1, int col0,
                                 int col1, float* out, long long ld_out, void* workspace, size_t workspace_bytes, void* stream) {
    PDGN_RANGE("pdgn_emd_allpairs");
    if (na < 0 || nb < 0 || n <= 0 || m <= 0) return PDGN_ERR_BAD_ARG;
    if (row0 < 0 || row1 > na || row0 > row1 || col0 < 0 || col1 > nb || col0 > col1) return PDGN_ERR_BAD_ARG;
    if (n > EM_MAX || m > EM_MAX) return PDGN_ERR_UNSUPPORTED;
    const int nrows = row1 - row0, ncols = col1 - col0;
    if (nrows == 0 || ncols == 0) return PDGN_OK;  // empty tile (pointers may be null)
    if (!A || !B || !out) return PDGN_ERR_BAD_ARG;
    if (ld_out < ncols || nrows > 65535) return ld_out < ncols ? PDGN_ERR_BAD_ARG : PDGN_ERR_UNSUPPORTED;
    const size_t need = ((size_t)nrows * n + (size_t)ncols * m) * 3 * sizeof(float);
    if (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 15)) return PDGN_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    float* SA = reinterpret_cast<float*>(workspace);       // kd-ordered copies of the rows / columns of this tile
    float* SB = SA + (size_t)nrows * n * 3;
    emd_sort_kernel<<<nrows, EM_SORT_T, 0, st>>>(A, row0, n, SA);
    PDGN_CHECK_LAUNCH();
    emd_sort_kernel<<<ncols, EM_SORT_T, 0, st>>>(B, col0, m, SB);
    PDGN_CHECK_LAUNCH();
    const size_t smem = (size_t)(2 * EM_MAX + 4 * EM_NCH + 4 * (EM_T / 32)) * sizeof(float4) + (size_t)EM_MAX * sizeof(float);
    PDGN_CUDA(cudaFuncSetAttribute(emd_allpairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // a CTA takes ~ms per cloud pair: keep strips short so the tail is small, but amortise the left-cloud load
    int strips = (8 * 2 * sms + nrows - 1) / nrows;
    if (strips > ncols) strips = ncols;
    if (strips < 1) strips = 1;
    const int rstrip = (ncols + strips - 1) / strips;
    strips = (ncols + rstrip - 1) / rstrip;
    emd_allpairs_kernel<<<dim3(strips, nrows), EM_T, smem, st>>>(SA, SB, ncols, n, m, rstrip, out, ld_out, 0);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}

extern "C" size_t pdgn_emd_paired_workspace(int b, int n, int m) { return pdgn_emd_allpairs_workspace(b, b, n, m); }

extern "C" int pdgn_emd_paired(const float* A, const float* B, int b, int n, int m, float* out, void* workspace,
                               size_t workspace_bytes, void* stream) {
    PDGN_RANGE("pdgn_emd_paired");
    if (b < 0 || n <= 0 || m <= 0) return PDGN_ERR_BAD_ARG;
    if (n > EM_MAX || m > EM_MAX) return PDGN_ERR_UNSUPPORTED;
    if (b == 0) return PDGN_OK;
    if (!A || !B || !out) return PDGN_ERR_BAD_ARG;
    const size_t need = ((size_t)b * n + (size_t)b * m) * 3 * sizeof(float);
    if (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 15)) return PDGN_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    float* SA = reinterpret_cast<float*>(workspace);
    float* SB = SA + (size_t)b * n * 3;
    const size_t smem = (size_t)(2 * EM_MAX + 4 * EM_NCH + 4 * (EM_T / 32)) * sizeof(float4) + (size_t)EM_MAX * sizeof(float);
    PDGN_CUDA(cudaFuncSetAttribute(emd_allpairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int at = 0; at < b; at += 65535) {  // gridDim.y limit
        const int cnt = b - at < 65535 ? b - at : 65535;
        emd_sort_kernel<<<cnt, EM_SORT_T, 0, st>>>(A, at, n, SA + (size_t)at * n * 3);
        PDGN_CHECK_LAUNCH();
        emd_sort_kernel<<<cnt, EM_SORT_T, 0, st>>>(B, at, m, SB + (size_t)at * m * 3);
        PDGN_CHECK_LAUNCH();
        emd_allpairs_kernel<<<dim3(1, cnt), EM_T, smem, st>>>(SA + (size_t)at * n * 3, SB + (size_t)at * m * 3, cnt, n, m, 1, out + at, 0, 1);
        PDGN_CHECK_LAUNCH();
    }
    return PDGN_OK;
}
