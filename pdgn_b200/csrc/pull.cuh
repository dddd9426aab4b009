// pull.cuh -- list-walk helpers shared by the backward (pull) kernels of gather.cu and pull_stream.cu.
#pragma once
#include "common.cuh"

namespace pdgn {

// Multi-channel pull: acc[ch] += value(ch, e) for the entries pb[a..b) in list order (index loads four at a time,
// every entry feeds all CC channel accumulators, so the list is read once per channel chunk).
template <int CC, class F>
__device__ __forceinline__ void pull_list(const int* __restrict__ pb, int a, int b, float (&acc)[CC], F value) {
    int q = a;
    for (; q + 4 <= b; q += 4) {
        const int e0 = __ldg(pb + q), e1 = __ldg(pb + q + 1), e2 = __ldg(pb + q + 2), e3 = __ldg(pb + q + 3);
#pragma unroll
        for (int ch = 0; ch < CC; ++ch) {
            acc[ch] += value(ch, e0);
            acc[ch] += value(ch, e1);
            acc[ch] += value(ch, e2);
            acc[ch] += value(ch, e3);
        }
    }
    for (; q < b; ++q) {
        const int e = __ldg(pb + q);
#pragma unroll
        for (int ch = 0; ch < CC; ++ch) acc[ch] += value(ch, e);
    }
}

#ifndef PDGN_PULL_ECACHE
#define PDGN_PULL_ECACHE 24   // 16 / 24 / 32 entries: grouping bwd 169 / 165 / 175 us, edge-feature bwd 272 / 261 / 273 us at C=256
#endif
constexpr int PULL_LONG = 64;  // lists longer than this are summed by a whole warp
constexpr int PULL_CC = 4;     // channel rows staged per CTA (upper bound)

// Long list: lane l sums entries a+l, a+l+32, ... in order, then a fixed butterfly combines the 32 partial sums.
template <int CC, class F>
__device__ __forceinline__ void pull_list_warp(const int* __restrict__ pb, int a, int b, int lane, float (&acc)[CC], F value) {
    for (int q = a + lane; q < b; q += 32) {
        const int e = __ldg(pb + q);
#pragma unroll
        for (int ch = 0; ch < CC; ++ch) acc[ch] += value(ch, e);
    }
#pragma unroll
    for (int ch = 0; ch < CC; ++ch)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[ch] += __shfl_xor_sync(kFull, acc[ch], o);
}


// Streaming pull over the inverse index (pull_stream.cu).  mode 0: grouping backward, 1: edge-feature backward, 2: interpolation
// backward.  *launched = false (and PDGN_OK) when the shape does not fit its shared-memory ring: the caller falls back.
int pull_stream_launch(int mode, const float* src, const int* offs, const int* pos, int b, int c, int ntargets, int rowlen, int k,
                       float* dst, cudaStream_t st, bool* launched, const float* wgt);

}  // namespace pdgn
