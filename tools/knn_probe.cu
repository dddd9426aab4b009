// knn_probe.cu -- design probes for the round-2 kNN kernel (run on the B200 box):
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/knn_probe tools/knn_probe.cu && tools/bin/knn_probe
// (1) pass-1 inner loop, lane = R queries, 2048 candidates broadcast from shared memory:
//       direct form (3 FADD + FMUL + 2 FFMA per pair, the exact reference chain) vs Gram form (3 FFMA per pair on a
//       precomputed |p|^2 plane), R = 1, 2, 4 queries per lane, 4..16 warps per SM -> cycles per (query, candidate) pair
// (2) ALU-pipe building blocks of the selection phase: VIMNMX.U16x2 / HMNMX2.BF16 / VIMNMX.U32 compare-exchange networks
//     and a 64-bit key compare-exchange, as warp-instructions per cycle per SM.
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int N = 2048;
constexpr int PL = N + N / 32 * 4;  // padded plane

__device__ __forceinline__ float min3(float a, float b, float c) { return fminf(fminf(a, b), c); }

template <int R, bool GRAM, int ORDER = 0>
__global__ void __launch_bounds__(512) pass1_probe(const float* __restrict__ pts, float* __restrict__ out, long long* cyc, int reps) {
    extern __shared__ __align__(16) float sm[];
    float* X = sm; float* Y = sm + PL; float* Z = sm + 2 * PL; float* P = sm + 3 * PL;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        const int pj = j + ((j >> 5) << 2);
        const float x = pts[3 * j], y = pts[3 * j + 1], z = pts[3 * j + 2];
        X[pj] = x; Y[pj] = y; Z[pj] = z; P[pj] = fmaf(z, z, fmaf(y, y, x * x));
    }
    __syncthreads();
    float qx[R], qy[R], qz[R], qq[R], acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int q = (threadIdx.x * R + r) % N;
        qx[r] = pts[3 * q]; qy[r] = pts[3 * q + 1]; qz[r] = pts[3 * q + 2];
        qq[r] = qx[r] * qx[r] + qy[r] * qy[r] + qz[r] * qz[r];
        if (GRAM) { qx[r] *= -2.f; qy[r] *= -2.f; qz[r] *= -2.f; }
        acc[r] = 0.f;
    }
    unsigned short* subs = reinterpret_cast<unsigned short*>(sm + 4 * PL) + (threadIdx.x >> 5) * (128 * 32 * R) + (threadIdx.x & 31);
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 2
        for (int sg = 0; sg < N / 16; ++sg) {
            const int base = sg * 16 + ((sg >> 1) << 2);
            float m[R];
#pragma unroll
            for (int r = 0; r < R; ++r) m[r] = __builtin_huge_valf();
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                const float4 x4 = *reinterpret_cast<const float4*>(X + base + 4 * qd);
                const float4 y4 = *reinterpret_cast<const float4*>(Y + base + 4 * qd);
                const float4 z4 = *reinterpret_cast<const float4*>(Z + base + 4 * qd);
                float4 p4 = make_float4(0, 0, 0, 0);
                if (GRAM) p4 = *reinterpret_cast<const float4*>(P + base + 4 * qd);
                if (GRAM && ORDER == 1) {
                    const float xs[4] = {x4.x, x4.y, x4.z, x4.w}, ys[4] = {y4.x, y4.y, y4.z, y4.w}, zs[4] = {z4.x, z4.y, z4.z, z4.w}, ps[4] = {p4.x, p4.y, p4.z, p4.w};
                    float g[R][4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int r = 0; r < R; ++r) g[r][u] = fmaf(qz[r], zs[u], ps[u]);
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int r = 0; r < R; ++r) g[r][u] = fmaf(qy[r], ys[u], g[r][u]);
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int r = 0; r < R; ++r) g[r][u] = fmaf(qx[r], xs[u], g[r][u]);
#pragma unroll
                    for (int r = 0; r < R; ++r) m[r] = min3(min3(m[r], g[r][0], g[r][1]), g[r][2], g[r][3]);
                } else
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    float g0, g1, g2, g3;
                    if (GRAM) {
                        g0 = fmaf(qx[r], x4.x, fmaf(qy[r], y4.x, fmaf(qz[r], z4.x, p4.x)));
                        g1 = fmaf(qx[r], x4.y, fmaf(qy[r], y4.y, fmaf(qz[r], z4.y, p4.y)));
                        g2 = fmaf(qx[r], x4.z, fmaf(qy[r], y4.z, fmaf(qz[r], z4.z, p4.z)));
                        g3 = fmaf(qx[r], x4.w, fmaf(qy[r], y4.w, fmaf(qz[r], z4.w, p4.w)));
                    } else {
#define D2(px, py, pz) ({ const float dx = qx[r] - (px), dy = qy[r] - (py), dz = qz[r] - (pz); fmaf(dz, dz, fmaf(dx, dx, dy * dy)); })
                        g0 = D2(x4.x, y4.x, z4.x); g1 = D2(x4.y, y4.y, z4.y); g2 = D2(x4.z, y4.z, z4.z); g3 = D2(x4.w, y4.w, z4.w);
                    }
                    m[r] = min3(min3(m[r], g0, g1), g2, g3);
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float v = fmaxf(GRAM ? m[r] + qq[r] : m[r], 0.f);
                if (R <= 4) subs[(r * 128 + sg) * 32] = (unsigned short)(__float_as_uint(v) >> 16);
                acc[r] += v;
            }
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) s += acc[r];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int R, bool GRAM, int ORDER = 0>
int run_pass1(int warps, int nsm, const float* pts) {
    float* out; long long* cyc;
    const int reps = 8;
    CK(cudaMalloc(&out, sizeof(float) * nsm * 512)); CK(cudaMalloc(&cyc, sizeof(long long) * nsm));
    const size_t smem = (size_t)4 * PL * 4 + (size_t)warps * 128 * 32 * (R <= 4 ? R : 1) * 2;
    CK(cudaFuncSetAttribute(pass1_probe<R, GRAM, ORDER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pass1_probe<R, GRAM, ORDER><<<nsm, warps * 32, smem>>>(pts, out, cyc, reps);
    CK(cudaDeviceSynchronize());
    pass1_probe<R, GRAM, ORDER><<<nsm, warps * 32, smem>>>(pts, out, cyc, reps);
    CK(cudaDeviceSynchronize());
    long long* h = new long long[nsm]; CK(cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
    double mean = 0; for (int i = 0; i < nsm; ++i) mean += h[i]; mean /= nsm;
    const double pairs_per_sm = (double)reps * N * warps * 32 * R;     // (query, candidate) pairs
    // lane-pairs per cycle per SM; the FP32 peak is 128 lane-instr/cycle/SM => pairs/cycle at the 6-instr roofline = 21.3
    printf("pass1 %-6s order=%d R=%2d warps/SM=%2d  smem %6zu B  cycles %9.0f  pairs/cycle/SM %6.2f  = %5.1f %% of the 6-instr/pair roofline (21.33)\n",
           GRAM ? "gram" : "direct", ORDER, R, warps, smem, mean, pairs_per_sm / mean, 100.0 * pairs_per_sm / mean / (128.0 / 6.0));
    cudaFree(out); cudaFree(cyc); delete[] h; return 0;
}

// ------------------------------------------------------------------------------------------- ALU-pipe probes
__device__ __forceinline__ unsigned minu2(unsigned a, unsigned b) { unsigned r; asm volatile("min.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ unsigned maxu2(unsigned a, unsigned b) { unsigned r; asm volatile("max.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ unsigned minbf2(unsigned a, unsigned b) { unsigned r; asm volatile("min.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ unsigned maxbf2(unsigned a, unsigned b) { unsigned r; asm volatile("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ unsigned minu(unsigned a, unsigned b) { unsigned r; asm volatile("min.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ unsigned maxu(unsigned a, unsigned b) { unsigned r; asm volatile("max.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

// MODE 0: u16x2 CE   1: bf16x2 CE   2: u32 CE   3: 64-bit key CE (setp + 4 selp)   4: u16x2 CE interleaved 1:2 with FFMA
template <int MODE>
__global__ void __launch_bounds__(1024) alu_probe(unsigned* out, unsigned seed, long long* cyc) {
    constexpr int NV = 16;
    unsigned v[NV];
    unsigned long long k[NV / 2];
    float f[NV];
    for (int i = 0; i < NV; ++i) { v[i] = seed * (i + 3) + threadIdx.x * 2654435761u; f[i] = (float)i + seed; }
    for (int i = 0; i < NV / 2; ++i) k[i] = ((unsigned long long)v[2 * i] << 32) | v[2 * i + 1];
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 2048; ++it) {
#pragma unroll
        for (int s = 1; s < NV; s <<= 1) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int p = i ^ s;
                if (p > i) {
                    if (MODE == 0 || MODE == 4) { const unsigned lo = minu2(v[i], v[p]), hi = maxu2(v[i], v[p]); v[i] = lo; v[p] = hi; }
                    if (MODE == 1) { const unsigned lo = minbf2(v[i], v[p]), hi = maxbf2(v[i], v[p]); v[i] = lo; v[p] = hi; }
                    if (MODE == 2) { const unsigned lo = minu(v[i], v[p]), hi = maxu(v[i], v[p]); v[i] = lo; v[p] = hi; }
                    if (MODE == 3 && p < NV / 2) { const bool sw = k[p] < k[i]; const unsigned long long a = sw ? k[p] : k[i], b = sw ? k[i] : k[p]; k[i] = a; k[p] = b; }
                    if (MODE == 4) { f[i] = fmaf(f[i], f[p], 1.0f); f[p] = fmaf(f[p], f[i], 0.5f); f[i] = fmaf(f[i], 0.999f, f[p]); f[p] = fmaf(f[p], 1.001f, f[i]); }
                }
            }
        }
    }
    const long long t1 = clock64();
    unsigned acc = 0;
    for (int i = 0; i < NV; ++i) acc += v[i] + __float_as_uint(f[i]);
    for (int i = 0; i < NV / 2; ++i) acc += (unsigned)k[i] + (unsigned)(k[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
int run_alu(const char* name, double ce_per_iter, int warps, int nsm) {
    unsigned* out; long long* cyc;
    CK(cudaMalloc(&out, sizeof(unsigned) * nsm * 1024)); CK(cudaMalloc(&cyc, sizeof(long long) * nsm));
    alu_probe<MODE><<<nsm, warps * 32>>>(out, 7u, cyc); CK(cudaDeviceSynchronize());
    alu_probe<MODE><<<nsm, warps * 32>>>(out, 7u, cyc); CK(cudaDeviceSynchronize());
    long long* h = new long long[nsm]; CK(cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
    double mean = 0; for (int i = 0; i < nsm; ++i) mean += h[i]; mean /= nsm;
    printf("%-44s warps/SM=%2d  compare-exchanges/cycle/SM %.3f (warp-wide)\n", name, warps, 2048.0 * ce_per_iter * warps / mean);
    cudaFree(out); cudaFree(cyc); delete[] h; return 0;
}

int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    const int nsm = pr.multiProcessorCount;
    printf("device %s, %d SMs\n", pr.name, nsm);
    float* h = new float[3 * N];
    unsigned s = 12345;
    for (int i = 0; i < 3 * N; ++i) { s = s * 1664525u + 1013904223u; h[i] = (float)(s >> 8) / 8388608.0f - 1.0f; }
    float* pts; CK(cudaMalloc(&pts, sizeof(float) * 3 * N)); CK(cudaMemcpy(pts, h, sizeof(float) * 3 * N, cudaMemcpyHostToDevice));
    for (int w = 4; w <= 8; w *= 2) {
        run_pass1<2, true, 1>(w, nsm, pts); run_pass1<4, true, 1>(w, nsm, pts); run_pass1<8, true, 1>(w, nsm, pts); run_pass1<16, true, 1>(w, nsm, pts);
        run_pass1<8, true, 0>(w, nsm, pts); run_pass1<16, true, 0>(w, nsm, pts); run_pass1<8, false, 0>(w, nsm, pts); run_pass1<16, false, 0>(w, nsm, pts);
    }
    for (int w = 4; w <= 16; w *= 2) {
        run_pass1<1, false>(w, nsm, pts); run_pass1<2, false>(w, nsm, pts); if (w <= 8) run_pass1<4, false>(w, nsm, pts);
        run_pass1<1, true>(w, nsm, pts);  run_pass1<2, true>(w, nsm, pts);  if (w <= 8) run_pass1<4, true>(w, nsm, pts);
    }
    for (int w = 4; w <= 16; w *= 2) {
        run_alu<0>("VIMNMX.U16x2 compare-exchange (2 instr)", 32, w, nsm);   // 4 stages x 8 pairs
        run_alu<1>("HMNMX2.BF16 compare-exchange (2 instr)", 32, w, nsm);
        run_alu<2>("VIMNMX.U32 compare-exchange (2 instr)", 32, w, nsm);
        run_alu<3>("64-bit key compare-exchange", 12, w, nsm);           // pairs with both indices < 8: stages 1,2,4 x 4
        run_alu<4>("u16x2 CE + 4 dependent FFMA", 32, w, nsm);
    }
    return 0;
}
