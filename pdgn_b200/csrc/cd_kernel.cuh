// cd_kernel.cuh -- device code of the all-pairs Chamfer kernel (see cd_allpairs.cu for the design notes).
// Templated so that tools/cd_tune.cu can instantiate variants; the library instantiates CdConfig (cd_allpairs.cu).
#pragma once
#include "common.cuh"

namespace pdgn {

constexpr int CD_HALF = 128;                // threads per half (one A cloud per half)
constexpr int CD_TILE = 2048;               // candidates per shared-memory stage
constexpr unsigned CD_INF_BITS = 0x7f800000u;

// variant bits (tuning switches; the shipped combination is CD_VARIANT in cd_allpairs.cu)
constexpr int CDV_PRED_RED = 1;    // predicated red.shared.min (inline PTX) instead of a divergent branch around atomicMin
constexpr int CDV_PREFETCH = 2;    // software-pipelined LDS.128 of the next 4 candidates
constexpr int CDV_RED4 = 4;        // one shared atomic per 4 candidates (lanes 0..3) instead of one per 2
constexpr int CDV_AOS = 8;         // candidates staged as float4 (x,y,z,0): x/z always land in even registers, y in odd
constexpr int CDV_WARPCOL = 16;    // every warp owns a column-minimum array: one STS.128 per 4 candidates, no shared atomics,
                                   // no initialisation (needs a single row block: npts <= R*128, and 4x the column storage)

// AoS [cloud][npts][3] -> SoA planes [cloud][3][npad]; pad entries replicate point 0 (harmless for minima).
__global__ void cd_pack_kernel(const float* __restrict__ src, int cloud0, int npts, int npad, float* __restrict__ dst) {
    const int cl = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= npad) return;
    const float* p = src + ((size_t)(cloud0 + cl) * npts + (j < npts ? j : 0)) * 3;
    float* d = dst + (size_t)cl * 3 * npad + j;
    d[0] = p[0];
    d[npad] = p[1];
    d[2 * (size_t)npad] = p[2];
}

// AoS [cloud][npts][3] -> padded AoS [cloud][npad] float4 (x,y,z,0) for the CDV_AOS candidate layout.
__global__ void cd_pack4_kernel(const float* __restrict__ src, int cloud0, int npts, int npad, float4* __restrict__ dst) {
    const int cl = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= npad) return;
    const float* p = src + ((size_t)(cloud0 + cl) * npts + (j < npts ? j : 0)) * 3;
    dst[(size_t)cl * npad + j] = make_float4(p[0], p[1], p[2], 0.f);
}

__device__ __forceinline__ void red_min_shared_pred(unsigned* addr, unsigned v, bool pred) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.u32 p, %2, 0;\n\t"
        "@p red.shared.min.u32 [%0], %1;\n\t}" ::"r"(smem_u32(addr)), "r"(v), "r"((unsigned)pred) : "memory");
}

// ablation switches for tools/cd_tune.cu only (results are WRONG with them; they price the two minimum streams)
constexpr int CDV_ABL_NOCOL = 32;  // no column minima
constexpr int CDV_ABL_NOROW = 64;  // no row minima

constexpr int CDV_IMIN = 128;      // minima as unsigned-integer min3 on the bit patterns (d2 >= +0: same order)
constexpr int CDV_IMIN_ROW = 256;  // ... row minima only
constexpr int CDV_IMIN_COL = 512;  // ... column minima only
constexpr int CDV_UNROLL2 = 1024;  // candidate loop unrolled twice (8 candidates per iteration)

template <int VAR, int WHICH>
__device__ __forceinline__ float cd_min3(float a, float b, float c) {
    if (VAR & (CDV_IMIN | WHICH)) return __uint_as_float(min(min(__float_as_uint(a), __float_as_uint(b)), __float_as_uint(c)));
    return min3(a, b, c);
}

// Two candidates against the thread's R rows: updates the row minima, returns the two column partial minima.
template <int R, int VAR = 0>
__device__ __forceinline__ void cd_two_candidates(const float (&qx)[R], const float (&qy)[R], const float (&qz)[R], float (&rowmin)[R],
                                                  float x0, float y0, float z0, float x1, float y1, float z1, float& c0, float& c1) {
    {
        const float a0 = d2_xyz(qx[0], qy[0], qz[0], x0, y0, z0), a1 = d2_xyz(qx[0], qy[0], qz[0], x1, y1, z1);
        const float b0 = d2_xyz(qx[1], qy[1], qz[1], x0, y0, z0), b1 = d2_xyz(qx[1], qy[1], qz[1], x1, y1, z1);
        if (!(VAR & CDV_ABL_NOROW)) {
            rowmin[0] = cd_min3<VAR, CDV_IMIN_ROW>(rowmin[0], a0, a1);
            rowmin[1] = cd_min3<VAR, CDV_IMIN_ROW>(rowmin[1], b0, b1);
        }
        c0 = fminf(a0, b0);
        c1 = fminf(a1, b1);
    }
#pragma unroll
    for (int k = 2; k < R; k += 2) {
        const float a0 = d2_xyz(qx[k], qy[k], qz[k], x0, y0, z0), a1 = d2_xyz(qx[k], qy[k], qz[k], x1, y1, z1);
        const float b0 = d2_xyz(qx[k + 1], qy[k + 1], qz[k + 1], x0, y0, z0);
        const float b1 = d2_xyz(qx[k + 1], qy[k + 1], qz[k + 1], x1, y1, z1);
        if (!(VAR & CDV_ABL_NOROW)) {
            rowmin[k] = cd_min3<VAR, CDV_IMIN_ROW>(rowmin[k], a0, a1);
            rowmin[k + 1] = cd_min3<VAR, CDV_IMIN_ROW>(rowmin[k + 1], b0, b1);
        }
        if (!(VAR & CDV_ABL_NOCOL)) {
            c0 = cd_min3<VAR, CDV_IMIN_COL>(c0, a0, b0);
            c1 = cd_min3<VAR, CDV_IMIN_COL>(c1, a1, b1);
        }
    }
}

// Four candidates (one LDS.128 per plane): row minima in registers, column minima -> warp CREDUX -> shared atomics.
template <int R, int VAR>
__device__ __forceinline__ void cd_four_candidates(const float (&qx)[R], const float (&qy)[R], const float (&qz)[R], float (&rowmin)[R],
                                                   const float4& X, const float4& Y, const float4& Z, unsigned* col, int lane) {
    float c0, c1, c2, c3;
    cd_two_candidates<R, VAR>(qx, qy, qz, rowmin, X.x, Y.x, Z.x, X.y, Y.y, Z.y, c0, c1);
    const unsigned r0 = __reduce_min_sync(kFull, __float_as_uint(c0));
    const unsigned r1 = __reduce_min_sync(kFull, __float_as_uint(c1));
    if (!(VAR & (CDV_RED4 | CDV_WARPCOL))) {
        if (VAR & CDV_PRED_RED) red_min_shared_pred(col + lane, lane ? r1 : r0, lane < 2);
        else if (lane < 2) atomicMin(col + lane, lane ? r1 : r0);
    }
    cd_two_candidates<R, VAR>(qx, qy, qz, rowmin, X.z, Y.z, Z.z, X.w, Y.w, Z.w, c2, c3);
    const unsigned r2 = __reduce_min_sync(kFull, __float_as_uint(c2));
    const unsigned r3 = __reduce_min_sync(kFull, __float_as_uint(c3));
    if (VAR & CDV_WARPCOL) {
        // this warp visits each candidate exactly once per cloud pair: its column minimum is final, plain store
        if (lane == 0) *reinterpret_cast<uint4*>(col) = make_uint4(r0, r1, r2, r3);
        return;
    }
    if (VAR & CDV_RED4) {
        const unsigned v = lane == 0 ? r0 : lane == 1 ? r1 : lane == 2 ? r2 : r3;
        if (VAR & CDV_PRED_RED) red_min_shared_pred(col + lane, v, lane < 4);
        else if (lane < 4) atomicMin(col + lane, v);
    } else {
        if (VAR & CDV_PRED_RED) red_min_shared_pred(col + 2 + lane, lane ? r3 : r2, lane < 2);
        else if (lane < 2) atomicMin(col + 2 + lane, lane ? r3 : r2);
    }
}

// Two candidates given as float4 (x,y,z,-): CDV_AOS path.
template <int R, int VAR>
__device__ __forceinline__ void cd_two_candidates_aos(const float (&qx)[R], const float (&qy)[R], const float (&qz)[R], float (&rowmin)[R],
                                                      const float4& P0, const float4& P1, unsigned* col, int lane) {
    float c0, c1;
    cd_two_candidates<R, VAR>(qx, qy, qz, rowmin, P0.x, P0.y, P0.z, P1.x, P1.y, P1.z, c0, c1);
    const unsigned r0 = __reduce_min_sync(kFull, __float_as_uint(c0));
    const unsigned r1 = __reduce_min_sync(kFull, __float_as_uint(c1));
    if (VAR & CDV_PRED_RED) red_min_shared_pred(col + lane, lane ? r1 : r0, lane < 2);
    else if (lane < 2) atomicMin(col + lane, lane ? r1 : r0);
}

// NH halves of 128 threads per CTA; each half owns one A cloud (R rows per thread, CD_ROWS = R*128 per row block).
template <int R, int NH, int MINB, int VAR, bool SYM = false>
__global__ void __launch_bounds__(NH * CD_HALF, MINB)
cd_allpairs_kernel(const float* __restrict__ PA, const float* __restrict__ PB, int nrows, int ncols, int npts, int npad,
                   int rstrip, float* __restrict__ out, long long ld_out, const unsigned* __restrict__ gate = nullptr) {
    constexpr int ROWS = R * CD_HALF;
    if (gate && *gate != 0u) return;  // the Gram-form kernel takes this tile (cd_gate_kernel)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int STAGE = ((VAR & CDV_AOS) ? 4 : 3) * CD_TILE;                // floats per stage
    float* tile = reinterpret_cast<float*>(smem_raw);                         // [2 stages][3 planes][CD_TILE] or [2][CD_TILE] float4
    constexpr int NCOL = (VAR & CDV_WARPCOL) ? 4 : 1;                         // column arrays per half
    unsigned* colmin = reinterpret_cast<unsigned*>(tile + 2 * STAGE);         // [NH][npad], WARPCOL: [NH][4 warps][npad]
    float* red = reinterpret_cast<float*>(colmin + NH * NCOL * (size_t)npad); // [NH][4 warps]
    uint64_t* bars = reinterpret_cast<uint64_t*>(red + NH * 4);              // [2 stages]

    const int tid = threadIdx.x, half = tid >> 7, ht = tid & (CD_HALF - 1), lane = tid & 31, hw = ht >> 5;
    int s = blockIdx.y * NH + half;
    const bool s_valid = s < nrows;
    if (!s_valid) s = nrows - 1;  // ragged row count: spare halves recompute the last cloud and discard it
    // SYM: A and B are the same cloud set => CD(s,r) == CD(r,s); this CTA only walks r >= its first row, the lower
    // triangle is filled by cd_mirror_kernel afterwards (saves ~half the work of the rr / ss matrices)
    const int r_begin = SYM ? max((int)(blockIdx.x * rstrip), (int)(blockIdx.y * NH)) : blockIdx.x * rstrip;
    const int r_end = min(ncols, (int)(blockIdx.x * rstrip) + rstrip);
    if (SYM && r_begin >= r_end) return;
    const int nrb = (npts + ROWS - 1) / ROWS;
    const int ncb = (npad + CD_TILE - 1) / CD_TILE;
    const int ntiles = (r_end - r_begin) * nrb * ncb;
    unsigned* halfcol = colmin + (size_t)half * NCOL * npad;
    unsigned* mycol = (VAR & CDV_WARPCOL) ? halfcol + (size_t)hw * npad : halfcol;

    if (!(VAR & CDV_WARPCOL))
        for (int j = ht; j < npad; j += CD_HALF) mycol[j] = CD_INF_BITS;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // tile t (flattened over r, row block, candidate block) -> stage t&1
    auto issue = [&](int t) {
        const int cb = t % ncb;
        const int r = r_begin + t / (ncb * nrb);
        const int c0 = cb * CD_TILE;
        const unsigned bytes = (unsigned)min(CD_TILE, npad - c0) * 4u;
        uint64_t* bar = &bars[t & 1];
        float* dst = tile + (t & 1) * STAGE;
        if (VAR & CDV_AOS) {
            mbar_expect_tx(bar, 4u * bytes);
            bulk_g2s(dst, PB + ((size_t)r * npad + c0) * 4, 4u * bytes, bar);
        } else {
            const float* src = PB + (size_t)r * 3 * npad + c0;
            mbar_expect_tx(bar, 3u * bytes);
            bulk_g2s(dst, src, bytes, bar);
            bulk_g2s(dst + CD_TILE, src + npad, bytes, bar);
            bulk_g2s(dst + 2 * CD_TILE, src + 2 * (size_t)npad, bytes, bar);
        }
    };
    if (tid == 0) {
        issue(0);
        if (ntiles > 1) issue(1);
    }

    float qx[R], qy[R], qz[R], rowmin[R];
    const float* arow = PA + (size_t)s * 3 * npad;
    const float inv_n = 1.0f / (float)npts;
    int t = 0;
    for (int r = r_begin; r < r_end; ++r) {
        float total = 0.f;
        for (int rb = 0; rb < nrb; ++rb) {
            const int i0 = rb * ROWS + ht * R;
            if (nrb > 1 || r == r_begin) {
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    const int i = (i0 + k < npts) ? i0 + k : 0;
                    qx[k] = arow[i];
                    qy[k] = arow[npad + i];
                    qz[k] = arow[2 * (size_t)npad + i];
                }
            }
#pragma unroll
            for (int k = 0; k < R; ++k) rowmin[k] = __int_as_float(CD_INF_BITS);

            for (int cb = 0; cb < ncb; ++cb, ++t) {
                const float* st = tile + (t & 1) * STAGE;
                const int cnt = min(CD_TILE, npad - cb * CD_TILE);
                unsigned* col = mycol + cb * CD_TILE;
                mbar_wait(&bars[t & 1], (unsigned)((t >> 1) & 1));
                if (VAR & CDV_AOS) {
                    const float4* st4 = reinterpret_cast<const float4*>(st);
                    if (VAR & CDV_PREFETCH) {
                        float4 P0 = st4[0], P1 = st4[1];
#pragma unroll 1
                        for (int j = 0; j < cnt; j += 2) {
                            const int jn = (j + 2 < cnt) ? j + 2 : j;
                            const float4 N0 = st4[jn], N1 = st4[jn + 1];
                            cd_two_candidates_aos<R, VAR>(qx, qy, qz, rowmin, P0, P1, col + j, lane);
                            P0 = N0; P1 = N1;
                        }
                    } else {
#pragma unroll 2
                        for (int j = 0; j < cnt; j += 2)
                            cd_two_candidates_aos<R, VAR>(qx, qy, qz, rowmin, st4[j], st4[j + 1], col + j, lane);
                    }
                } else if (VAR & CDV_PREFETCH) {
                    float4 X = *reinterpret_cast<const float4*>(st);
                    float4 Y = *reinterpret_cast<const float4*>(st + CD_TILE);
                    float4 Z = *reinterpret_cast<const float4*>(st + 2 * CD_TILE);
#pragma unroll((VAR & CDV_UNROLL2) ? 2 : 1)
                    for (int j = 0; j < cnt; j += 4) {
                        const int jn = (j + 4 < cnt) ? j + 4 : j;  // last iteration re-reads its own quad
                        const float4 Xn = *reinterpret_cast<const float4*>(st + jn);
                        const float4 Yn = *reinterpret_cast<const float4*>(st + CD_TILE + jn);
                        const float4 Zn = *reinterpret_cast<const float4*>(st + 2 * CD_TILE + jn);
                        cd_four_candidates<R, VAR>(qx, qy, qz, rowmin, X, Y, Z, col + j, lane);
                        X = Xn; Y = Yn; Z = Zn;
                    }
                } else {
#pragma unroll 1
                    for (int j = 0; j < cnt; j += 4) {
                        const float4 X = *reinterpret_cast<const float4*>(st + j);
                        const float4 Y = *reinterpret_cast<const float4*>(st + CD_TILE + j);
                        const float4 Z = *reinterpret_cast<const float4*>(st + 2 * CD_TILE + j);
                        cd_four_candidates<R, VAR>(qx, qy, qz, rowmin, X, Y, Z, col + j, lane);
                    }
                }
                __syncthreads();  // stage drained by every warp; this tile's column atomics are done
                if (tid == 0 && t + 2 < ntiles) {
                    fence_proxy_async();
                    issue(t + 2);
                }
            }
            const int nvalid = npts - i0;
#pragma unroll
            for (int k = 0; k < R; ++k)
                if (k < nvalid) total += rowmin[k];
        }
        // cloud pair (s, r) complete: fold this half's column minima, reset them for the next r
        if (VAR & CDV_WARPCOL) {
            for (int j = ht; j < npts; j += CD_HALF)
                total += __uint_as_float(min(min(halfcol[j], halfcol[npad + j]), min(halfcol[2 * npad + j], halfcol[3 * npad + j])));
        } else {
            for (int j = ht; j < npad; j += CD_HALF) {
                if (j < npts) total += __uint_as_float(mycol[j]);
                mycol[j] = CD_INF_BITS;
            }
        }
        total = warp_sum(total);
        if (lane == 0) red[half * 4 + hw] = total;
        __syncthreads();
        if (ht == 0 && s_valid) {
            const float* rr = red + half * 4;
            out[(size_t)s * ld_out + r] = (rr[0] + rr[1] + rr[2] + rr[3]) * inv_n;
        }
    }
}

// =====================================================================================================
// Gram-form variant (round 2): e_ij = |a_i|^2 - 2 a_i.b_j as THREE FFMAs on precomputed coefficients instead of the six
// instructions of the reference chain; d_ij = e_ij + |b_j|^2.  The column minimum is taken on e (|b_j|^2 is added once per
// column when the pair is folded), the row minimum on e + |b_j|^2 (one FADD per pair).  Measured with tools/cd_probe.cu: 84 %
// of the 6-instruction roofline against 68 % for the direct form in the same loop.  This is the arithmetic of the reference's
// own default path (torch bmm Gram form, evaluation_metrics.py:35-45) and meets its 1e-5 contract, but it is not the
// bit pattern of NmDistanceKernel: the direct-form kernel above stays (PDGN_B200_CD_EXACT=1, clouds of more than 2048
// points, and every tile whose clouds are not centred -- see cd_gate_kernel).
// =====================================================================================================
constexpr int CDG_TILE = 1024;   // candidates per stage: 2 stages x 4 planes x 4 KB = 32 KB (+ 64 KB of column arrays: 2 CTAs per SM)

// AoS [cloud][npts][3] -> planes [cloud][4][npad]: x, y, z, |p|^2 (fma chain); pad entries replicate point 0.  Also the
// largest |p|^2 of everything packed (stats[0], float bits, atomicMax: all values >= 0).
__global__ void cd_pack_gram_kernel(const float* __restrict__ src, int cloud0, int npts, int npad, float* __restrict__ dst,
                                    unsigned* __restrict__ stats) {
    const int cl = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    float pp = 0.f;
    if (j < npad) {
        const float* p = src + ((size_t)(cloud0 + cl) * npts + (j < npts ? j : 0)) * 3;
        const float x = p[0], y = p[1], z = p[2];
        pp = __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x)));
        float* d = dst + (size_t)cl * 4 * npad + j;
        d[0] = x;
        d[npad] = y;
        d[2 * (size_t)npad] = z;
        d[3 * (size_t)npad] = pp;
    }
    // NaN / inf coordinates: report +inf so that the gate sends the tile to the direct-form kernel
    unsigned bits = (pp <= 3.402823466e+38f) ? __float_as_uint(pp) : CD_INF_BITS;
    bits = __reduce_max_sync(kFull, bits);
    if ((threadIdx.x & 31) == 0) atomicMax(stats, bits);
}

// stats[1] = mean nearest-neighbour squared distance INSIDE cloud `cloud` of `src`, estimated on 64 evenly spaced sample points
// (one CTA of 8 warps; warp w takes samples w, w+8, ...; the point itself is skipped, a duplicate gives 0).  It sets the scale of
// the pair scalars the Gram form's rounding error has to be measured against.
__global__ void __launch_bounds__(256) cd_scale_kernel(const float* __restrict__ src, int cloud, int npts, unsigned* __restrict__ stats) {
    __shared__ float part[8];
    const float* p = src + (size_t)cloud * npts * 3;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nsamp = npts < 64 ? npts : 64;
    float acc = 0.f;
    for (int sidx = warp; sidx < nsamp; sidx += 8) {
        const int i = (int)(((long long)sidx * npts) / nsamp);
        const float qx = p[3 * i], qy = p[3 * i + 1], qz = p[3 * i + 2];
        float best = kInf;
        for (int j = lane; j < npts; j += 32)
            if (j != i) best = fminf(best, d2_xyz(qx, qy, qz, p[3 * j], p[3 * j + 1], p[3 * j + 2]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = fminf(best, __shfl_xor_sync(kFull, best, o));
        acc += best <= 3.402823466e+38f ? best : 0.f;   // a one-point or non-finite cloud counts as scale 0 => direct form
    }
    if (lane == 0) part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += part[w];
        stats[1] = __float_as_uint(t / (float)nsamp);
    }
}

// gate = 1: the Gram-form kernel computes this tile; 0: the direct-form kernel.  The Gram form's rounding error per distance
// is ~3u(|a|^2 + |b|^2) <= 6u R^2 (u = 2^-24, R^2 = largest |p|^2 of the tile); it averages as a random walk over the 2 npts
// minima of a pair, whose scalar is ~2 dmean (dmean = mean nearest-neighbour d^2, estimated by cd_scale_kernel):
//     predicted relative error = 3u R^2 / (sqrt(2 npts) dmean)
// (measured / predicted on tools/cd_check.py: sphere 1.2e-6 / 1.2e-6, cube 2.2e-7 / 6e-7, thin plane 6.7e-6 / 9e-6).  The Gram
// form runs when the prediction is <= 3e-6, a third of the 1e-5 contract: every normalised shape set; clouds far from the
// origin, scattered positions, near-planar / near-linear sets with tiny neighbour distances and non-finite coordinates take
// the direct form.
__global__ void cd_gate_kernel(const unsigned* __restrict__ stats, int npts, unsigned* __restrict__ gate) {
    const float rmax2 = __uint_as_float(stats[0]), dmean = __uint_as_float(stats[1]);
    const float pred = 3.f * 5.9604645e-8f * rmax2 / (sqrtf(2.f * (float)npts) * dmean);   // inf / NaN for dmean == 0 or non-finite data
    *gate = (pred <= 3e-6f) ? 1u : 0u;
}

__device__ __forceinline__ int cdg_key(float v) {  // order-preserving float -> signed int (e can be negative); an involution
    const int b = __float_as_int(v);
    return b ^ ((b >> 31) & 0x7fffffff);
}

// Two candidates against the thread's R rows (Gram form): row minima on e + bb, column partial minima on e.  SPLIT: two
// independent column chains per candidate (rows 0..R/2-1 and R/2..R-1) merged at the end -- half the dependent-FMNMX3 depth.
template <int R, bool SPLIT = false>
__device__ __forceinline__ void cdg_two_candidates(const float (&ax)[R], const float (&ay)[R], const float (&az)[R], const float (&aa)[R],
                                                   float (&rowmin)[R], float x0, float y0, float z0, float b0, float x1, float y1,
                                                   float z1, float b1, float& c0, float& c1) {
    c0 = kInf;
    c1 = kInf;
    float d0 = kInf, d1 = kInf;
#pragma unroll
    for (int k = 0; k < R; k += 2) {
        const float e00 = __fmaf_rn(ax[k], x0, __fmaf_rn(ay[k], y0, __fmaf_rn(az[k], z0, aa[k])));
        const float e01 = __fmaf_rn(ax[k], x1, __fmaf_rn(ay[k], y1, __fmaf_rn(az[k], z1, aa[k])));
        const float e10 = __fmaf_rn(ax[k + 1], x0, __fmaf_rn(ay[k + 1], y0, __fmaf_rn(az[k + 1], z0, aa[k + 1])));
        const float e11 = __fmaf_rn(ax[k + 1], x1, __fmaf_rn(ay[k + 1], y1, __fmaf_rn(az[k + 1], z1, aa[k + 1])));
        rowmin[k] = min3(rowmin[k], __fadd_rn(e00, b0), __fadd_rn(e01, b1));
        rowmin[k + 1] = min3(rowmin[k + 1], __fadd_rn(e10, b0), __fadd_rn(e11, b1));
        if (SPLIT && k >= R / 2) {
            d0 = min3(d0, e00, e10);
            d1 = min3(d1, e01, e11);
        } else {
            c0 = min3(c0, e00, e10);
            c1 = min3(c1, e01, e11);
        }
    }
    if (SPLIT) {
        c0 = fminf(c0, d0);
        c1 = fminf(c1, d1);
    }
}

// Same CTA shape as cd_allpairs_kernel (NH halves of 128 threads, R rows per thread, one row block: npts <= R*128), per-warp
// column arrays (plain STS.128 of four CREDUX results, no atomics).  PA / PB: [cloud][4][npad] from cd_pack_gram_kernel.
template <int R, int NH, int MINB, bool SYM, int OPT = 0>
__global__ void __launch_bounds__(NH * CD_HALF, MINB)
cd_gram_kernel(const float* __restrict__ PA, const float* __restrict__ PB, int nrows, int ncols, int npts, int npad, int rstrip,
               float* __restrict__ out, long long ld_out, const unsigned* __restrict__ gate) {
    if (gate && *gate == 0u) return;  // the direct-form kernel takes this tile
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int STAGE = 4 * CDG_TILE;
    float* tile = reinterpret_cast<float*>(smem_raw);                       // [2 stages][4 planes][CDG_TILE]
    int* colmin = reinterpret_cast<int*>(tile + 2 * STAGE);                 // [NH][4 warps][npad] signed keys of e
    float* red = reinterpret_cast<float*>(colmin + NH * 4 * (size_t)npad);  // [NH][4 warps]
    uint64_t* bars = reinterpret_cast<uint64_t*>(red + NH * 4);             // [2 stages]

    const int tid = threadIdx.x, half = tid >> 7, ht = tid & (CD_HALF - 1), lane = tid & 31, hw = ht >> 5;
    int s = blockIdx.y * NH + half;
    const bool s_valid = s < nrows;
    if (!s_valid) s = nrows - 1;
    const int r_begin = SYM ? max((int)(blockIdx.x * rstrip), (int)(blockIdx.y * NH)) : blockIdx.x * rstrip;
    const int r_end = min(ncols, (int)(blockIdx.x * rstrip) + rstrip);
    if (SYM && r_begin >= r_end) return;
    const int ncb = (npad + CDG_TILE - 1) / CDG_TILE;
    const int ntiles = (r_end - r_begin) * ncb;
    int* halfcol = colmin + (size_t)half * 4 * npad;
    int* mycol = halfcol + (size_t)hw * npad;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int t) {
        const int cb = t % ncb;
        const int r = r_begin + t / ncb;
        const int c0 = cb * CDG_TILE;
        const unsigned bytes = (unsigned)min(CDG_TILE, npad - c0) * 4u;
        uint64_t* bar = &bars[t & 1];
        float* dst = tile + (t & 1) * STAGE;
        const float* src = PB + (size_t)r * 4 * npad + c0;
        mbar_expect_tx(bar, 4u * bytes);
#pragma unroll
        for (int pl = 0; pl < 4; ++pl) bulk_g2s(dst + pl * CDG_TILE, src + pl * (size_t)npad, bytes, bar);
    };
    if (tid == 0) {
        issue(0);
        if (ntiles > 1) issue(1);
    }

    // this thread's R rows of cloud s: coefficients -2a and |a|^2 (rows beyond npts replicate point 0 and are not summed)
    float ax[R], ay[R], az[R], aa[R], rowmin[R];
    {
        const float* arow = PA + (size_t)s * 4 * npad;
        const int i0 = ht * R;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int i = (i0 + k < npts) ? i0 + k : 0;
            ax[k] = -2.f * arow[i];
            ay[k] = -2.f * arow[npad + i];
            az[k] = -2.f * arow[2 * (size_t)npad + i];
            aa[k] = arow[3 * (size_t)npad + i];
        }
    }
    const float inv_n = 1.0f / (float)npts;
    const int nvalid = npts - ht * R;
    int t = 0;
    for (int r = r_begin; r < r_end; ++r) {
#pragma unroll
        for (int k = 0; k < R; ++k) rowmin[k] = kInf;
        for (int cb = 0; cb < ncb; ++cb, ++t) {
            const float* st = tile + (t & 1) * STAGE;
            const int cnt = min(CDG_TILE, npad - cb * CDG_TILE);
            int* col = mycol + cb * CDG_TILE;
            mbar_wait(&bars[t & 1], (unsigned)((t >> 1) & 1));
            auto quad = [&](const float4& X, const float4& Y, const float4& Z, const float4& Q, int j) {
                float c0, c1, c2, c3;
                cdg_two_candidates<R, (OPT & 4) != 0>(ax, ay, az, aa, rowmin, X.x, Y.x, Z.x, Q.x, X.y, Y.y, Z.y, Q.y, c0, c1);
                const int k0 = __reduce_min_sync(kFull, cdg_key(c0));
                const int k1 = __reduce_min_sync(kFull, cdg_key(c1));
                cdg_two_candidates<R, (OPT & 4) != 0>(ax, ay, az, aa, rowmin, X.z, Y.z, Z.z, Q.z, X.w, Y.w, Z.w, Q.w, c2, c3);
                const int k2 = __reduce_min_sync(kFull, cdg_key(c2));
                const int k3 = __reduce_min_sync(kFull, cdg_key(c3));
                // this warp meets each candidate exactly once per cloud pair: its column minimum is final, plain store
                if (lane == 0) *reinterpret_cast<int4*>(col + j) = make_int4(k0, k1, k2, k3);
            };
            if (OPT & 8) {
                // deferred epilogue: the column partial minima of quad j are reduced across the warp (CREDUX) and stored while
                // the FFMAs of quad j+1 issue, instead of in an ALU-only burst at the end of the body
                float p0 = kInf, p1 = kInf, p2 = kInf, p3 = kInf;
#pragma unroll 1
                for (int j = 0; j < cnt; j += 4) {
                    const float4 X = *reinterpret_cast<const float4*>(st + j);
                    const float4 Y = *reinterpret_cast<const float4*>(st + CDG_TILE + j);
                    const float4 Z = *reinterpret_cast<const float4*>(st + 2 * CDG_TILE + j);
                    const float4 Q = *reinterpret_cast<const float4*>(st + 3 * CDG_TILE + j);
                    float c0, c1, c2, c3;
                    cdg_two_candidates<R, (OPT & 4) != 0>(ax, ay, az, aa, rowmin, X.x, Y.x, Z.x, Q.x, X.y, Y.y, Z.y, Q.y, c0, c1);
                    if (j > 0) {
                        const int k0 = __reduce_min_sync(kFull, cdg_key(p0)), k1 = __reduce_min_sync(kFull, cdg_key(p1));
                        const int k2 = __reduce_min_sync(kFull, cdg_key(p2)), k3 = __reduce_min_sync(kFull, cdg_key(p3));
                        if (lane == 0) *reinterpret_cast<int4*>(col + j - 4) = make_int4(k0, k1, k2, k3);
                    }
                    cdg_two_candidates<R, (OPT & 4) != 0>(ax, ay, az, aa, rowmin, X.z, Y.z, Z.z, Q.z, X.w, Y.w, Z.w, Q.w, c2, c3);
                    p0 = c0; p1 = c1; p2 = c2; p3 = c3;
                }
                const int k0 = __reduce_min_sync(kFull, cdg_key(p0)), k1 = __reduce_min_sync(kFull, cdg_key(p1));
                const int k2 = __reduce_min_sync(kFull, cdg_key(p2)), k3 = __reduce_min_sync(kFull, cdg_key(p3));
                if (lane == 0) *reinterpret_cast<int4*>(col + cnt - 4) = make_int4(k0, k1, k2, k3);
            } else
            if (OPT & 2) {  // software-pipelined loads: the next quad is in flight while this one is consumed
                float4 X = *reinterpret_cast<const float4*>(st), Y = *reinterpret_cast<const float4*>(st + CDG_TILE);
                float4 Z = *reinterpret_cast<const float4*>(st + 2 * CDG_TILE), Q = *reinterpret_cast<const float4*>(st + 3 * CDG_TILE);
#pragma unroll((OPT & 1) ? 2 : 1)
                for (int j = 0; j < cnt; j += 4) {
                    const int jn = (j + 4 < cnt) ? j + 4 : j;
                    const float4 Xn = *reinterpret_cast<const float4*>(st + jn), Yn = *reinterpret_cast<const float4*>(st + CDG_TILE + jn);
                    const float4 Zn = *reinterpret_cast<const float4*>(st + 2 * CDG_TILE + jn), Qn = *reinterpret_cast<const float4*>(st + 3 * CDG_TILE + jn);
                    quad(X, Y, Z, Q, j);
                    X = Xn; Y = Yn; Z = Zn; Q = Qn;
                }
            } else {
#pragma unroll((OPT & 1) ? 2 : 1)
                for (int j = 0; j < cnt; j += 4) {
                    const float4 X = *reinterpret_cast<const float4*>(st + j);
                    const float4 Y = *reinterpret_cast<const float4*>(st + CDG_TILE + j);
                    const float4 Z = *reinterpret_cast<const float4*>(st + 2 * CDG_TILE + j);
                    const float4 Q = *reinterpret_cast<const float4*>(st + 3 * CDG_TILE + j);
                    quad(X, Y, Z, Q, j);
                }
            }
            __syncthreads();  // stage drained by every warp
            if (tid == 0 && t + 2 < ntiles) {
                fence_proxy_async();
                issue(t + 2);
            }
        }
        float total = 0.f;
#pragma unroll
        for (int k = 0; k < R; ++k)
            if (k < nvalid) total += rowmin[k];
        // fold this half's column minima: min over its 4 warps of e, plus |b_j|^2
        const float* bb = PB + ((size_t)r * 4 + 3) * npad;
        for (int j = ht; j < npts; j += CD_HALF) {
            const int key = min(min(halfcol[j], halfcol[npad + j]), min(halfcol[2 * npad + j], halfcol[3 * npad + j]));
            total += __fadd_rn(__int_as_float(key ^ ((key >> 31) & 0x7fffffff)), __ldg(bb + j));
        }
        total = warp_sum(total);
        if (lane == 0) red[half * 4 + hw] = total;
        __syncthreads();
        if (ht == 0 && s_valid) {
            const float* rr = red + half * 4;
            // a cloud against itself is exactly 0 (the Gram form would leave rounding noise on the diagonal)
            out[(size_t)s * ld_out + r] = (SYM && s == r) ? 0.f : (rr[0] + rr[1] + rr[2] + rr[3]) * inv_n;
        }
    }
}

template <int NH>
constexpr size_t cdg_smem_bytes(int npad) {
    return (size_t)(2 * 4 * CDG_TILE + NH * 4 * (size_t)npad + NH * 4) * 4 + 2 * sizeof(uint64_t);
}

// out[r][s] = out[s][r] for r > s (square matrix)
__global__ void cd_mirror_kernel(float* __restrict__ out, int n, long long ld) {
    const int r = blockIdx.y * blockDim.y + threadIdx.y, s = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n && s < r) out[(size_t)r * ld + s] = out[(size_t)s * ld + r];
}

template <int NH, int VAR = 0>
constexpr size_t cd_smem_bytes(int npad) {
    return (size_t)(2 * ((VAR & CDV_AOS) ? 4 : 3) * CD_TILE + NH * ((VAR & CDV_WARPCOL) ? 4 : 1) * (size_t)npad + NH * 4) * 4 +
           2 * sizeof(uint64_t);
}

}  // namespace pdgn
