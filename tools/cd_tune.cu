// cd_tune.cu -- variant sweep for the all-pairs Chamfer kernel (tools/; not part of the library).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o /tmp/cd_tune tools/cd_tune.cu && /tmp/cd_tune [clouds]
// Every variant is checked against variant (16,2,2,0) bit for bit (minima are exact; only the final float sums
// could differ, and they use the same order) and timed with CUDA events.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../pdgn_b200/csrc/cd_kernel.cuh"
using namespace pdgn;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int R, int NH, int MINB, int VAR>
float run(const char* name, const float* PA, const float* PB_soa, const float* PB_aos, int n, int npts, int npad, float* out, const float* ref, int sms) {
    auto kern = cd_allpairs_kernel<R, NH, MINB, VAR>;
    const float* PB = (VAR & CDV_AOS) ? PB_aos : PB_soa;
    const size_t smem = cd_smem_bytes<NH, VAR>(npad);
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NH * CD_HALF, smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
    const int spairs = (n + NH - 1) / NH;
    const int target = 20 * occ * sms;
    int strips = (target + spairs - 1) / spairs; if (strips > n) strips = n; if (strips < 1) strips = 1;
    const int rstrip = (n + strips - 1) / strips; strips = (n + rstrip - 1) / rstrip;
    CK(cudaMemset(out, 0, sizeof(float) * n * n));
    kern<<<dim3(strips, spairs), NH * CD_HALF, smem>>>(PA, PB, n, n, npts, npad, rstrip, out, n);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int it = 0; it < 3; ++it) {
        CK(cudaEventRecord(e0));
        kern<<<dim3(strips, spairs), NH * CD_HALF, smem>>>(PA, PB, n, n, npts, npad, rstrip, out, n);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    std::vector<float> h(n * (size_t)n);
    CK(cudaMemcpy(h.data(), out, sizeof(float) * n * n, cudaMemcpyDeviceToHost));
    double maxrel = 0; if (ref) for (size_t i = 0; i < h.size(); ++i) maxrel = fmax(maxrel, fabs(h[i] - ref[i]) / fmax(1e-30, fabs(ref[i])));
    const double pairs = (double)n * n, rate = pairs / (best * 1e-3);
    const double frac = rate * npts * (double)npts * 6.0 / (sms * 128.0 * 1.965e9);
    printf("%-28s R=%2d NH=%d regs=%3d spill=%3zuB occ=%d grid=%dx%d  %8.3f ms  %9.0f pairs/s  issue-roofline %.3f  maxrel-vs-base %.2e\n", name, R, NH,
           fa.numRegs, (size_t)fa.localSizeBytes, occ, strips, spairs, best, rate, frac, maxrel);
    return best;
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 296, npts = argc > 2 ? atoi(argv[2]) : 2048, npad = (npts + 15) & ~15;
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0)); const int sms = pr.multiProcessorCount;
    std::vector<float> hA((size_t)n * npts * 3), hB((size_t)n * npts * 3);
    srand(1);
    auto fill = [&](std::vector<float>& v) { for (size_t i = 0; i < v.size(); i += 3) { float x = rand() / (float)RAND_MAX * 2 - 1, y = rand() / (float)RAND_MAX * 2 - 1, z = rand() / (float)RAND_MAX * 2 - 1; float r = sqrtf(x * x + y * y + z * z) + 1e-6f; v[i] = x / r; v[i + 1] = y / r; v[i + 2] = z / r; } };
    fill(hA); fill(hB);
    float *dA, *dB, *PA, *PB, *PB4, *out;
    CK(cudaMalloc(&dA, hA.size() * 4)); CK(cudaMalloc(&dB, hB.size() * 4));
    CK(cudaMalloc(&PA, (size_t)n * 3 * npad * 4)); CK(cudaMalloc(&PB, (size_t)n * 3 * npad * 4)); CK(cudaMalloc(&out, (size_t)n * n * 4)); CK(cudaMalloc(&PB4, (size_t)n * 4 * npad * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
    cd_pack_kernel<<<dim3((npad + 255) / 256, n), 256>>>(dA, 0, npts, npad, PA);
    cd_pack_kernel<<<dim3((npad + 255) / 256, n), 256>>>(dB, 0, npts, npad, PB);
    cd_pack4_kernel<<<dim3((npad + 255) / 256, n), 256>>>(dB, 0, npts, npad, reinterpret_cast<float4*>(PB4));
    CK(cudaDeviceSynchronize());
    printf("%d x %d clouds of %d points, %d SMs\n", n, n, npts, sms);
    run<16, 2, 2, 0>("base", PA, PB, PB4, n, npts, npad, out, nullptr, sms);
    std::vector<float> ref((size_t)n * n); CK(cudaMemcpy(ref.data(), out, ref.size() * 4, cudaMemcpyDeviceToHost));
    run<16, 2, 2, 3>("pred-red+prefetch", PA, PB, PB4, n, npts, npad, out, ref.data(), sms);
    run<16, 2, 2, 18>("prefetch+warpcol (shipped)", PA, PB, PB4, n, npts, npad, out, ref.data(), sms);
    run<16, 2, 2, 16>("warpcol", PA, PB, PB4, n, npts, npad, out, ref.data(), sms);
    run<16, 2, 2, 16 + 128>("warpcol + integer min3", PA, PB, PB4, n, npts, npad, out, ref.data(), sms);
    run<16, 2, 2, 18 + 128>("prefetch+warpcol + integer min3", PA, PB, PB4, n, npts, npad, out, ref.data(), sms);
    run<16, 2, 2, 18 + 256>("prefetch+warpcol + imin rows", PA, PB, PB4, n, npts, npad, out, ref.data(), sms);
    run<16, 2, 2, 18 + 512>("prefetch+warpcol + imin cols", PA, PB, PB4, n, npts, npad, out, ref.data(), sms);
    run<16, 2, 2, 16 + 256>("warpcol + imin rows", PA, PB, PB4, n, npts, npad, out, ref.data(), sms);
    run<16, 2, 2, 16 + 512>("warpcol + imin cols", PA, PB, PB4, n, npts, npad, out, ref.data(), sms);
    run<16, 2, 2, 3 + 128>("pred-red+prefetch + imin", PA, PB, PB4, n, npts, npad, out, ref.data(), sms);
    run<16, 2, 2, 18 + 256 + 1024>("prefetch+warpcol+imin rows+unroll2", PA, PB, PB4, n, npts, npad, out, ref.data(), sms);
    run<16, 2, 2, 18 + 1024>("prefetch+warpcol+unroll2", PA, PB, PB4, n, npts, npad, out, ref.data(), sms);
    run<16, 2, 2, 18 + 128 + 1024>("prefetch+warpcol+imin+unroll2", PA, PB, PB4, n, npts, npad, out, ref.data(), sms);
    run<16, 2, 2, 18 + 32>("ABLATION no col-min", PA, PB, PB4, n, npts, npad, out, ref.data(), sms);
    run<16, 2, 2, 18 + 64>("ABLATION no row-min", PA, PB, PB4, n, npts, npad, out, ref.data(), sms);
    run<16, 2, 2, 8>("aos", PA, PB, PB4, n, npts, npad, out, ref.data(), sms);
    run<16, 2, 2, 9>("aos+pred-red", PA, PB, PB4, n, npts, npad, out, ref.data(), sms);
    run<16, 2, 2, 11>("aos+pred-red+prefetch", PA, PB, PB4, n, npts, npad, out, ref.data(), sms);
    run<16, 1, 3, 9>("aos 1 half x3 pred-red", PA, PB, PB4, n, npts, npad, out, ref.data(), sms);
    run<8, 2, 4, 9>("aos R8 x4 pred-red", PA, PB, PB4, n, npts, npad, out, ref.data(), sms);
    return 0;
}
