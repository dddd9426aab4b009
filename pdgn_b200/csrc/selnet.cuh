// selnet.cuh -- lane-private bitonic networks on registers (shared by knn_gram.cu and knn_feat_tc.cu).
#pragma once
#include "common.cuh"

namespace pdgn {

__device__ __forceinline__ unsigned kq_min2(unsigned a, unsigned b) { unsigned r; asm("min.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ unsigned kq_max2(unsigned a, unsigned b) { unsigned r; asm("max.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }


// ascending bitonic sort of N registers with the given compare-exchange
template <int N, typename CE>
__device__ __forceinline__ void kq_bitonic_sort(CE ce) {
#pragma unroll
    for (int size = 2; size <= N; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const int p = i ^ stride;
                if (p > i) ce(i, p, (i & size) == 0 || size == N);
            }
        }
    }
}
// ascending bitonic MERGE of a bitonic sequence of N registers
template <int N, typename CE>
__device__ __forceinline__ void kq_bitonic_merge(CE ce) {
#pragma unroll
    for (int stride = N >> 1; stride > 0; stride >>= 1) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int p = i ^ stride;
            if (p > i) ce(i, p, true);
        }
    }
}


// k-th smallest (k <= 32) of 64 group minima given as 16-bit patterns that order like unsigned integers (bf16 bits of
// non-negative floats): grp(i), i = 0..63.  Packed u16x2 network: the upper half is stored complemented, so one VIMNMX.U16x2
// sorts one half ascending and the other descending.
template <class G>
__device__ __forceinline__ unsigned kq_kth_of_64(G grp, int k) {
    unsigned v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (grp(i) & 0xffffu) | (((grp(i + 32) & 0xffffu) ^ 0xffffu) << 16);
    kq_bitonic_sort<32>([&](int i, int p, bool up) {
        const unsigned lo = kq_min2(v[i], v[p]), hi = kq_max2(v[i], v[p]);
        v[i] = up ? lo : hi;
        v[p] = up ? hi : lo;
    });
    // (lower halves ascending) ++ (upper halves descending) is bitonic: the element-wise minimum holds the 32 smallest of the 64
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = min(v[i] & 0xffffu, (v[i] >> 16) ^ 0xffffu);
    kq_bitonic_merge<32>([&](int i, int p, bool) {
        const unsigned lo = min(v[i], v[p]), hi = max(v[i], v[p]);
        v[i] = lo;
        v[p] = hi;
    });
    unsigned t = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) t = (i < k) ? max(t, v[i]) : t;   // v ascending: v[k-1] without a dynamically indexed array
    return t;
}

}  // namespace pdgn
