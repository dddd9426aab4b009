// microbench.cu -- pipe-throughput probes for the design decisions in DESIGN.md (run on the B200 box):
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/microbench tools/microbench.cu && gpurun_out/microbench
// Reports warp-instructions per cycle per SM for scalar FFMA, packed FFMA2/FADD2/FMUL2 (fma.rn.f32x2 ...),
// FMNMX3, mixes of FMA-pipe and ALU-pipe work, and CREDUX, so that the FP32 roofline used in bench.py is
// the measured one and the packed-vs-scalar choice in the distance kernels is evidenced.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)
__device__ __forceinline__ u64 pk(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void upk(u64 v, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(v)); }
__device__ __forceinline__ u64 add2(u64 a,u64 b){u64 r; asm volatile("add.rn.f32x2 %0,%1,%2;":"=l"(r):"l"(a),"l"(b)); return r;}
__device__ __forceinline__ u64 mul2(u64 a,u64 b){u64 r; asm volatile("mul.rn.f32x2 %0,%1,%2;":"=l"(r):"l"(a),"l"(b)); return r;}
__device__ __forceinline__ u64 fma2(u64 a,u64 b,u64 c){u64 r; asm volatile("fma.rn.f32x2 %0,%1,%2,%3;":"=l"(r):"l"(a),"l"(b),"l"(c)); return r;}
__device__ __forceinline__ float ffma(float a,float b,float c){float r; asm volatile("fma.rn.f32 %0,%1,%2,%3;":"=f"(r):"f"(a),"f"(b),"f"(c)); return r;}
__device__ __forceinline__ float fadd(float a,float b){float r; asm volatile("add.rn.f32 %0,%1,%2;":"=f"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ float fmin3(float a,float b,float c){float r; asm volatile("min.f32 %0,%1,%2,%3;":"=f"(r):"f"(a),"f"(b),"f"(c)); return r;}
__device__ __forceinline__ float fmin2(float a,float b){float r; asm volatile("min.f32 %0,%1,%2;":"=f"(r):"f"(a),"f"(b)); return r;}

constexpr int ITERS = 4096;
constexpr int NCH = 8;   // independent chains per thread

// mode 0: scalar FFMA   1: FFMA2   2: FADD2   3: FMUL2   4: FMNMX3   5: FMNMX   6: scalar FADD
// mode 7: 6 FFMA2-class + 2 FMNMX3 per group (CD inner-loop mix, packed)   8: 6 scalar + 1 FMNMX3 (scalar mix)
// mode 9: CREDUX.MIN   10: 3 FADD2 + FMUL2 + 2 FFMA2 (exact distance chain, packed)   11: same chain scalar
template <int MODE>
__global__ void __launch_bounds__(1024) probe(float* out, float seed, long long* cyc) {
    float a[NCH]; u64 p[NCH];
    for (int i = 0; i < NCH; ++i) { a[i] = seed + i + threadIdx.x; p[i] = pk(a[i], a[i] + 1.f); }
    float s = seed; u64 s2 = pk(seed, seed * 0.5f);
    long long t0 = clock64();
    #pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < NCH; ++i) {
            if (MODE == 0) a[i] = ffma(a[i], s, s);
            if (MODE == 1) p[i] = fma2(p[i], s2, s2);
            if (MODE == 2) p[i] = add2(p[i], s2);
            if (MODE == 3) p[i] = mul2(p[i], s2);
            if (MODE == 4) a[i] = fmin3(a[i], s, a[(i + 1) % NCH]);
            if (MODE == 5) a[i] = fmin2(a[i], s);
            if (MODE == 6) a[i] = fadd(a[i], s);
            if (MODE == 7) { // per chain: 6 packed fma-pipe + 2 alu min3
                u64 d = add2(p[i], s2); u64 e = add2(p[i], d); u64 f = add2(d, e);
                u64 t = mul2(e, e); t = fma2(d, d, t); t = fma2(f, f, t); p[i] = t;
                float lo, hi; upk(t, lo, hi); a[i] = fmin3(a[i], lo, hi); a[(i + 1) % NCH] = fmin3(a[(i + 1) % NCH], lo, hi);
            }
            if (MODE == 8) {
                float d = fadd(a[i], s), e = fadd(a[i], d), f = fadd(d, e);
                float t = e * e; t = ffma(d, d, t); t = ffma(f, f, t); a[i] = fmin3(a[i], t, s);
            }
            if (MODE == 9) { unsigned r = __reduce_min_sync(0xffffffffu, __float_as_uint(a[i])); a[i] = __uint_as_float(r) + 1.f; }
            if (MODE == 10) {
                u64 d = add2(p[i], s2); u64 e = add2(p[i], d); u64 f = add2(d, e);
                u64 t = mul2(e, e); t = fma2(d, d, t); t = fma2(f, f, t); p[i] = t;
            }
            if (MODE == 11) {
                float d = fadd(a[i], s), e = fadd(a[i], d), f = fadd(d, e);
                float t = e * e; t = ffma(d, d, t); t = ffma(f, f, t); a[i] = t;
            }
        }
    }
    long long t1 = clock64();
    float acc = 0; for (int i = 0; i < NCH; ++i) { float lo, hi; upk(p[i], lo, hi); acc += a[i] + lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
int run(const char* name, double instr_per_chain_iter, int warps, int nsm) {
    // ONE CTA per SM (grid = #SMs) with `warps` warps, so every CTA is co-resident and clock64 deltas are exact.
    float* out; long long* cyc;
    int grid = nsm; int threads = warps * 32;
    CK(cudaMalloc(&out, sizeof(float) * grid * 1024)); CK(cudaMalloc(&cyc, sizeof(long long) * grid));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<grid, threads>>>(out, 1.0f, cyc); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); probe<MODE><<<grid, threads>>>(out, 1.0f, cyc); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long* h = new long long[grid]; CK(cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    double mean = 0; for (int i = 0; i < grid; ++i) mean += h[i]; mean /= grid;
    double winstr_per_sm = (double)ITERS * NCH * instr_per_chain_iter * warps;
    printf("%-44s warps/SM=%2d  warp-instr/cycle/SM=%.3f  (cycles/CTA %.0f, %.3f ms, implied clock %.0f MHz)\n", name, warps,
           winstr_per_sm / mean, mean, ms, mean / (ms * 1e3));
    cudaFree(out); cudaFree(cyc); delete[] h; return 0;
}

int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    int nsm = pr.multiProcessorCount;
    printf("device %s, %d SMs, clockRate %d kHz\n", pr.name, nsm, pr.clockRate);
    for (int c = 8; c <= 32; c *= 2) {
        run<0>("scalar FFMA", 1, c, nsm);
        run<6>("scalar FADD", 1, c, nsm);
        run<1>("packed FFMA2", 1, c, nsm);
        run<2>("packed FADD2", 1, c, nsm);
        run<3>("packed FMUL2", 1, c, nsm);
        run<4>("FMNMX3", 1, c, nsm);
        run<5>("FMNMX", 1, c, nsm);
        run<11>("scalar distance chain (6 fma-pipe)", 6, c, nsm);
        run<10>("packed distance chain (6 x f32x2)", 6, c, nsm);
        run<8>("scalar chain + 1 FMNMX3 (7 instr)", 7, c, nsm);
        run<7>("packed chain + 2 FMNMX3 (8 instr)", 8, c, nsm);
        run<9>("CREDUX.MIN + FADD (2 instr)", 2, c, nsm);
    }
    return 0;
}
