// pipe_probe.cu -- which FP32 instructions share an issue pipe on B200?  (tools/ only; run on the GPU box)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/pipe_probe tools/pipe_probe.cu && tools/bin/pipe_probe
// Every probe runs NCH independent chains per thread, so latency never limits; the result is warp-instructions per cycle per SM
// (4.0 = one instruction per scheduler per cycle).
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ float ffma(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0,%1,%2,%3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float fadd(float a, float b) { float r; asm volatile("add.rn.f32 %0,%1,%2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fmul(float a, float b) { float r; asm volatile("mul.rn.f32 %0,%1,%2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
constexpr int NCH = 16, ITERS = 2048;

// MODE: 0 FFMA (acc = a*b+acc, 3 distinct sources)   1 FADD   2 FMUL   3 FFMA+FADD 1:1   4 FFMA+FMUL 1:1   5 FADD+FMUL 1:1
//       6 2 FFMA + 1 FADD   7 2 FFMA + 1 FMUL + 1 FADD (balanced Gram)   8 3 FADD + 1 FMUL + 2 FFMA (direct distance)
//       9 FFMA with acc in place and a shared multiplier (a*s+acc: 2 register reads)
template <int MODE>
__global__ void __launch_bounds__(1024) probe(float* out, float seed, long long* cyc) {
    float a[NCH], b[NCH], c[NCH];
    for (int i = 0; i < NCH; ++i) { a[i] = seed + i + threadIdx.x; b[i] = a[i] * 0.5f + 1.f; c[i] = a[i] - 3.f; }
    const float s = seed * 0.999f;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            if (MODE == 0) c[i] = ffma(a[i], b[i], c[i]);
            if (MODE == 1) c[i] = fadd(c[i], a[i]);
            if (MODE == 2) c[i] = fmul(c[i], a[i]);
            if (MODE == 3) { c[i] = ffma(a[i], b[i], c[i]); b[i] = fadd(b[i], a[i]); }
            if (MODE == 4) { c[i] = ffma(a[i], b[i], c[i]); b[i] = fmul(b[i], a[i]); }
            if (MODE == 5) { c[i] = fadd(c[i], a[i]); b[i] = fmul(b[i], a[i]); }
            if (MODE == 6) { c[i] = ffma(a[i], b[i], c[i]); b[i] = ffma(a[i], c[i], b[i]); a[i] = fadd(a[i], s); }
            if (MODE == 7) { const float t = fmul(a[i], b[i]); const float u = ffma(a[i], c[i], s); const float v = ffma(b[i], c[i], t); c[i] = fadd(u, v); }
            if (MODE == 8) { const float d = fadd(a[i], s), e = fadd(b[i], s), f = fadd(c[i], s); float t = fmul(e, e); t = ffma(d, d, t); c[i] = ffma(f, f, t); }
            if (MODE == 9) c[i] = ffma(a[i], s, c[i]);
        }
    }
    const long long t1 = clock64();
    float acc = 0;
    for (int i = 0; i < NCH; ++i) acc += a[i] + b[i] + c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
int run(const char* name, double instr_per_chain, int warps, int nsm) {
    float* out; long long* cyc;
    CK(cudaMalloc(&out, sizeof(float) * nsm * 1024)); CK(cudaMalloc(&cyc, sizeof(long long) * nsm));
    probe<MODE><<<nsm, warps * 32>>>(out, 1.0f, cyc); CK(cudaDeviceSynchronize());
    probe<MODE><<<nsm, warps * 32>>>(out, 1.0f, cyc); CK(cudaDeviceSynchronize());
    long long* h = new long long[nsm]; CK(cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
    double mean = 0; for (int i = 0; i < nsm; ++i) mean += h[i]; mean /= nsm;
    printf("%-52s warps/SM=%2d  warp-instr/cycle/SM %.3f\n", name, warps, (double)ITERS * NCH * instr_per_chain * warps / mean);
    cudaFree(out); cudaFree(cyc); delete[] h; return 0;
}
int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    const int nsm = pr.multiProcessorCount;
    for (int w = 8; w <= 32; w *= 2) {
        run<0>("FFMA c=a*b+c (3 distinct sources)", 1, w, nsm);
        run<9>("FFMA c=a*s+c (shared multiplier)", 1, w, nsm);
        run<1>("FADD", 1, w, nsm);
        run<2>("FMUL", 1, w, nsm);
        run<3>("FFMA + FADD 1:1", 2, w, nsm);
        run<4>("FFMA + FMUL 1:1", 2, w, nsm);
        run<5>("FADD + FMUL 1:1", 2, w, nsm);
        run<6>("2 FFMA + 1 FADD", 3, w, nsm);
        run<7>("2 FFMA + FMUL + FADD (balanced Gram dot)", 4, w, nsm);
        run<8>("3 FADD + FMUL + 2 FFMA (direct distance)", 6, w, nsm);
    }
    return 0;
}
