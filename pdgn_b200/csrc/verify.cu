// verify.cu -- PDGN_B200_VERIFY=1 index range check for the gather / scatter entry points (debugging aid, off by default).
#include "common.cuh"

namespace pdgn {

template <typename I>
__global__ void verify_idx_kernel(const I* __restrict__ idx, size_t count, int n, unsigned long long* bad) {
    unsigned long long mine = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        const long long v = (long long)idx[i];
        mine += (v < 0 || v >= (long long)n) ? 1ull : 0ull;
    }
    mine = __reduce_add_sync(kFull, (unsigned)mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(bad, mine);
}

__device__ unsigned long long g_verify_bad;

template <typename I>
static int verify_idx(const I* idx, size_t count, int n, cudaStream_t st) {
    if (count == 0) return PDGN_OK;
    if (!idx) return PDGN_ERR_BAD_ARG;
    unsigned long long* bad = nullptr;
    PDGN_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&bad), g_verify_bad));
    PDGN_CUDA(cudaMemsetAsync(bad, 0, sizeof(unsigned long long), st));
    const int blocks = (int)((count + 1023) / 1024 < 1184 ? (count + 1023) / 1024 : 1184);
    verify_idx_kernel<I><<<blocks, 256, 0, st>>>(idx, count, n, bad);
    PDGN_CHECK_LAUNCH();
    unsigned long long host = 0;
    PDGN_CUDA(cudaMemcpyAsync(&host, bad, sizeof(host), cudaMemcpyDeviceToHost, st));
    PDGN_CUDA(cudaStreamSynchronize(st));
    return host ? PDGN_ERR_INDEX : PDGN_OK;
}

int verify_idx32(const int* idx, size_t count, int n, cudaStream_t st) { return verify_idx<int>(idx, count, n, st); }
int verify_idx64(const int64_t* idx, size_t count, int n, cudaStream_t st) { return verify_idx<int64_t>(idx, count, n, st); }

}  // namespace pdgn
