// cd_probe.cu -- would a Gram-form inner loop pay for the all-pairs Chamfer kernel?  (tools/ only; run on the GPU box)
// Both variants keep 16 A points per thread in registers, stream 2048 B points from shared memory (LDS.128 broadcast) and take
// BOTH directional minima from one distance, like cd_allpairs_kernel:
//   direct : d = fma(dz,dz, fma(dx,dx, dy*dy)) on three FADD differences (6 FMA-pipe instructions, the shipped arithmetic)
//   gram   : e = fma(ax',bx, fma(ay',by, fma(az',bz, aa))) (3 FFMA, three sources each); column minimum on e (+ |b|^2 once per
//            column), row minimum on e + |b|^2 (one FADD per pair)
// Output: cycles per point pair per SM-lane slot and the fraction of the 6-instr/pair roofline.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
constexpr int N = 2048, R = 16;
__device__ __forceinline__ float min3(float a, float b, float c) { return fminf(fminf(a, b), c); }

template <bool GRAM>
__global__ void __launch_bounds__(256, 2) cd_probe(const float* __restrict__ pts, float* __restrict__ out, long long* cyc, int reps) {
    extern __shared__ __align__(16) float sm[];
    float* X = sm; float* Y = sm + N; float* Z = sm + 2 * N; float* P = sm + 3 * N;
    unsigned* col = reinterpret_cast<unsigned*>(sm + 4 * N) + (threadIdx.x >> 5) * N;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        const float x = pts[3 * j], y = pts[3 * j + 1], z = pts[3 * j + 2];
        X[j] = x; Y[j] = y; Z[j] = z; P[j] = fmaf(z, z, fmaf(y, y, x * x));
    }
    __syncthreads();
    float qx[R], qy[R], qz[R], qa[R], rowmin[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const int q = (threadIdx.x * R + k) % N;
        qx[k] = pts[3 * q]; qy[k] = pts[3 * q + 1]; qz[k] = pts[3 * q + 2];
        qa[k] = qx[k] * qx[k] + qy[k] * qy[k] + qz[k] * qz[k];
        if (GRAM) { qx[k] *= -2.f; qy[k] *= -2.f; qz[k] *= -2.f; }
        rowmin[k] = __builtin_huge_valf();
    }
    const int lane = threadIdx.x & 31;
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 1
        for (int j = 0; j < N; j += 4) {
            const float4 x4 = *reinterpret_cast<const float4*>(X + j), y4 = *reinterpret_cast<const float4*>(Y + j);
            const float4 z4 = *reinterpret_cast<const float4*>(Z + j);
            float4 p4 = make_float4(0, 0, 0, 0);
            if (GRAM) p4 = *reinterpret_cast<const float4*>(P + j);
            const float xs[4] = {x4.x, x4.y, x4.z, x4.w}, ys[4] = {y4.x, y4.y, y4.z, y4.w}, zs[4] = {z4.x, z4.y, z4.z, z4.w}, ps[4] = {p4.x, p4.y, p4.z, p4.w};
            unsigned r4[4];
#pragma unroll
            for (int u = 0; u < 4; u += 2) {
                float c0 = __builtin_huge_valf(), c1 = __builtin_huge_valf();
#pragma unroll
                for (int k = 0; k < R; k += 2) {
                    float a0, a1, b0, b1;      // (row k, cand u), (row k, cand u+1), (row k+1, cand u), (row k+1, cand u+1)
                    if (GRAM) {
                        a0 = fmaf(qx[k], xs[u], fmaf(qy[k], ys[u], fmaf(qz[k], zs[u], qa[k])));
                        a1 = fmaf(qx[k], xs[u + 1], fmaf(qy[k], ys[u + 1], fmaf(qz[k], zs[u + 1], qa[k])));
                        b0 = fmaf(qx[k + 1], xs[u], fmaf(qy[k + 1], ys[u], fmaf(qz[k + 1], zs[u], qa[k + 1])));
                        b1 = fmaf(qx[k + 1], xs[u + 1], fmaf(qy[k + 1], ys[u + 1], fmaf(qz[k + 1], zs[u + 1], qa[k + 1])));
                        rowmin[k] = min3(rowmin[k], a0 + ps[u], a1 + ps[u + 1]);
                        rowmin[k + 1] = min3(rowmin[k + 1], b0 + ps[u], b1 + ps[u + 1]);
                    } else {
#define D2(qi, c) ({ const float dx = qx[qi] - xs[c], dy = qy[qi] - ys[c], dz = qz[qi] - zs[c]; fmaf(dz, dz, fmaf(dx, dx, dy * dy)); })
                        a0 = D2(k, u); a1 = D2(k, u + 1); b0 = D2(k + 1, u); b1 = D2(k + 1, u + 1);
                        rowmin[k] = min3(rowmin[k], a0, a1);
                        rowmin[k + 1] = min3(rowmin[k + 1], b0, b1);
                    }
                    c0 = min3(c0, a0, b0);
                    c1 = min3(c1, a1, b1);
                }
                unsigned k0 = __float_as_uint(c0), k1 = __float_as_uint(c1);
                if (GRAM) {  // e can be negative: order-preserving signed key
                    k0 ^= (unsigned)((int)k0 >> 31) & 0x7fffffffu;
                    k1 ^= (unsigned)((int)k1 >> 31) & 0x7fffffffu;
                    r4[u] = (unsigned)__reduce_min_sync(0xffffffffu, (int)k0);
                    r4[u + 1] = (unsigned)__reduce_min_sync(0xffffffffu, (int)k1);
                } else {
                    r4[u] = __reduce_min_sync(0xffffffffu, k0);
                    r4[u + 1] = __reduce_min_sync(0xffffffffu, k1);
                }
            }
            if (lane == 0) *reinterpret_cast<uint4*>(col + j) = make_uint4(r4[0], r4[1], r4[2], r4[3]);
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int k = 0; k < R; ++k) s += rowmin[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(col[threadIdx.x]);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <bool GRAM>
int run(int nsm) {
    float* out; long long* cyc; float* pts;
    float* h = new float[3 * N];
    unsigned s = 12345;
    for (int i = 0; i < 3 * N; ++i) { s = s * 1664525u + 1013904223u; h[i] = (float)(s >> 8) / 8388608.0f - 1.0f; }
    CK(cudaMalloc(&pts, sizeof(float) * 3 * N)); CK(cudaMemcpy(pts, h, sizeof(float) * 3 * N, cudaMemcpyHostToDevice));
    const int ctas = 2 * nsm, reps = 8;
    CK(cudaMalloc(&out, sizeof(float) * ctas * 256)); CK(cudaMalloc(&cyc, sizeof(long long) * ctas));
    const size_t smem = (size_t)4 * N * 4 + (size_t)8 * N * 4;
    CK(cudaFuncSetAttribute(cd_probe<GRAM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cd_probe<GRAM><<<ctas, 256, smem>>>(pts, out, cyc, reps); CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int big = 64;
    CK(cudaEventRecord(e0));
    cd_probe<GRAM><<<ctas, 256, smem>>>(pts, out, cyc, big);
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    {
        const double pairs = (double)ctas * big * N * 256 * R;
        printf("cd inner loop %-6s by CUDA events: %.3f ms for %.3e pairs = %5.1f %% of the 6-instr/pair roofline at 1965 MHz\n", GRAM ? "gram" : "direct", ms,
               pairs, 100.0 * pairs * 6.0 / (ms * 1e-3) / (nsm * 128.0 * 1.965e9));
    }
    cd_probe<GRAM><<<ctas, 256, smem>>>(pts, out, cyc, reps); CK(cudaDeviceSynchronize());
    long long* hc = new long long[ctas]; CK(cudaMemcpy(hc, cyc, sizeof(long long) * ctas, cudaMemcpyDeviceToHost));
    double mean = 0; for (int i = 0; i < ctas; ++i) mean += hc[i]; mean /= ctas;
    const double pairs_per_sm = 2.0 * reps * N * 256 * R;   // two CTAs per SM
    printf("cd inner loop %-6s 2 CTAs x 8 warps/SM, 16 rows per thread: cycles %9.0f  pairs/cycle/SM %6.2f = %5.1f %% of the 6-instr/pair roofline\n",
           GRAM ? "gram" : "direct", mean, pairs_per_sm / mean, 100.0 * pairs_per_sm / mean / (128.0 / 6.0));
    return 0;
}
int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    run<false>(pr.multiProcessorCount);
    run<true>(pr.multiProcessorCount);
    return 0;
}
