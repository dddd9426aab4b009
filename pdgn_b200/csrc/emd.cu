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
// Bound: FP32 issue + MUFU (one ex2 per point pair per sweep, one more rsqrt/sqrt in the cost sweep):
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
__device__ __forceinline__ float sqrt_fast(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

constexpr int EM_T = 512;             // threads per CTA
constexpr int EM_R = 4;               // points of each cloud owned per thread (8 warps/SMSP keep the MUFU queue fed)
constexpr int EM_MAX = EM_T * EM_R;   // 2048 points per cloud

__global__ void __launch_bounds__(EM_T, 2)
emd_allpairs_kernel(const float* __restrict__ A, const float* __restrict__ B, int ncols, int n, int m, int rstrip,
                    float* __restrict__ out, long long ld_out) {
    extern __shared__ __align__(16) float4 em_sm[];
    float4* L = em_sm;            // [EM_MAX] left cloud:  x, y, z, ratioL
    float4* Rr = em_sm + EM_MAX;  // [EM_MAX] right cloud: x, y, z, remainR (sweep 1) / ratioR (sweep 3)
    __shared__ float red[EM_T / 32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int s = blockIdx.y;
    const int r_begin = blockIdx.x * rstrip, r_end = min(ncols, r_begin + rstrip);
    const float multiL = n >= m ? 1.f : (float)(m / n), multiR = n >= m ? (float)(n / m) : 1.f;

    const float* Ap = A + (size_t)s * n * 3;
    for (int k = t; k < EM_MAX; k += EM_T) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < n) { v.x = Ap[k * 3]; v.y = Ap[k * 3 + 1]; v.z = Ap[k * 3 + 2]; }
        L[k] = v;
    }
    for (int r = r_begin; r < r_end; ++r) {
        const float* Bp = B + (size_t)r * m * 3;
        __syncthreads();  // previous pair fully consumed
        float remL[EM_R], remR[EM_R], ratL[EM_R];
#pragma unroll
        for (int i = 0; i < EM_R; ++i) {
            const int p = t + i * EM_T;
            remL[i] = p < n ? multiL : 0.f;
            remR[i] = p < m ? multiR : 0.f;
            float4 v = make_float4(0.f, 0.f, 0.f, remR[i]);
            if (p < m) { v.x = Bp[p * 3]; v.y = Bp[p * 3 + 1]; v.z = Bp[p * 3 + 2]; }
            Rr[p] = v;
        }
        __syncthreads();
        float cost = 0.f;
        float level = -16384.f;  // -4^7, then /4 per level down to -4^-1
        for (int j = 7; j > -2; --j, level *= 0.25f) {
            const float l2 = level * 1.4426950408889634f;  // level * log2(e)
            float ox[EM_R], oy[EM_R], oz[EM_R], acc[EM_R];
            // ---- sweep 1: ratioL[k] = remainL[k] / (1e-9 + sum_l exp(level*d2) * remainR[l])
#pragma unroll
            for (int i = 0; i < EM_R; ++i) {
                const float4 v = L[t + i * EM_T];
                ox[i] = v.x; oy[i] = v.y; oz[i] = v.z;
                acc[i] = 1e-9f;
            }
#pragma unroll 2
            for (int l = 0; l < m; ++l) {
                const float4 q = Rr[l];
#pragma unroll
                for (int i = 0; i < EM_R; ++i)
                    acc[i] = __fmaf_rn(exp2_fast(l2 * d2_xyz(q.x, q.y, q.z, ox[i], oy[i], oz[i])), q.w, acc[i]);
            }
#pragma unroll
            for (int i = 0; i < EM_R; ++i) {
                ratL[i] = remL[i] / acc[i];
                L[t + i * EM_T].w = ratL[i];
            }
            __syncthreads();
            // ---- sweep 2: sumr[l] = remainR[l] * sum_k exp(level*d2) * ratioL[k]; consumption; ratioR; remainR
#pragma unroll
            for (int i = 0; i < EM_R; ++i) {
                const float4 v = Rr[t + i * EM_T];
                ox[i] = v.x; oy[i] = v.y; oz[i] = v.z;
                acc[i] = 0.f;
            }
#pragma unroll 2
            for (int k = 0; k < n; ++k) {
                const float4 p = L[k];
#pragma unroll
                for (int i = 0; i < EM_R; ++i)
                    acc[i] = __fmaf_rn(exp2_fast(l2 * d2_xyz(ox[i], oy[i], oz[i], p.x, p.y, p.z)), p.w, acc[i]);
            }
#pragma unroll
            for (int i = 0; i < EM_R; ++i) {
                const float sumr = acc[i] * remR[i];
                const float consumption = fminf(remR[i] / (sumr + 1e-9f), 1.0f);
                Rr[t + i * EM_T].w = consumption * remR[i];  // ratioR
                remR[i] = fmaxf(0.0f, remR[i] - sumr);
            }
            __syncthreads();
            // ---- sweep 3: w = exp(level*d2)*ratioL[k]*ratioR[l]; remainL[k] -= sum_l w; cost += w*|p_k - q_l|
#pragma unroll
            for (int i = 0; i < EM_R; ++i) {
                const float4 v = L[t + i * EM_T];
                ox[i] = v.x; oy[i] = v.y; oz[i] = v.z;
                acc[i] = 0.f;
            }
#pragma unroll 2
            for (int l = 0; l < m; ++l) {
                const float4 q = Rr[l];
#pragma unroll
                for (int i = 0; i < EM_R; ++i) {
                    const float d2 = d2_xyz(q.x, q.y, q.z, ox[i], oy[i], oz[i]);
                    const float w = exp2_fast(l2 * d2) * ratL[i] * q.w;
                    acc[i] += w;
                    cost = __fmaf_rn(w, sqrt_fast(d2), cost);
                }
            }
            __syncthreads();  // everyone is done reading ratioR before remainR goes back into the .w slots
#pragma unroll
            for (int i = 0; i < EM_R; ++i) {
                remL[i] = fmaxf(0.0f, remL[i] - acc[i]);
                Rr[t + i * EM_T].w = remR[i];
            }
            __syncthreads();
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

extern "C" int pdgn_emd_allpairs(const float* A, const float* B, int na, int nb, int n, int m, int row0, int row1, int col0,
                                 int col1, float* out, long long ld_out, void* stream) {
    if (na < 0 || nb < 0 || n <= 0 || m <= 0) return PDGN_ERR_BAD_ARG;
    if (row0 < 0 || row1 > na || row0 > row1 || col0 < 0 || col1 > nb || col0 > col1) return PDGN_ERR_BAD_ARG;
    if (n > EM_MAX || m > EM_MAX) return PDGN_ERR_UNSUPPORTED;
    const int nrows = row1 - row0, ncols = col1 - col0;
    if (nrows == 0 || ncols == 0) return PDGN_OK;  // empty tile (pointers may be null)
    if (!A || !B || !out) return PDGN_ERR_BAD_ARG;
    if (ld_out < ncols || nrows > 65535) return ld_out < ncols ? PDGN_ERR_BAD_ARG : PDGN_ERR_UNSUPPORTED;
    const size_t smem = (size_t)2 * EM_MAX * sizeof(float4);
    PDGN_CUDA(cudaFuncSetAttribute(emd_allpairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // a CTA takes ~ms per cloud pair: keep strips short so the tail is small, but amortise the left-cloud load
    int strips = (8 * 2 * sms + nrows - 1) / nrows;
    if (strips > ncols) strips = ncols;
    if (strips < 1) strips = 1;
    const int rstrip = (ncols + strips - 1) / strips;
    strips = (ncols + rstrip - 1) / rstrip;
    emd_allpairs_kernel<<<dim3(strips, nrows), EM_T, smem, (cudaStream_t)stream>>>(
        A + (size_t)row0 * n * 3, B + (size_t)col0 * m * 3, ncols, n, m, rstrip, out, ld_out);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}
