// gather.cu -- bandwidth kernels: grouping / interpolation / edge-feature gathers and their backward scatters.
//
// Replaces grouping_forward_cuda_kernel_fast / grouping_backward_cuda_kernel
// (lib/pointops/src/grouping/grouping_cuda_kernel.cu:60-75, :28-46), interpolation_forward_cuda_kernel_fast /
// interpolation_backward_cuda_kernel (lib/pointops/src/interpolation/interpolation_cuda_kernel.cu:181-195,
// :90-114) and the index_select/repeat/cat composition of get_edge_features{,_xyz}
// (models/PDGNet_v2.py:461-477, :505-525).
//
// These are HBM-bound: the output (or grad_out) stream dominates the bytes.  Each thread owns 4 consecutive
// output positions (one 16-byte streaming store per channel), reads its 4 indices ONCE and walks the
// channels of its chunk, so idx is read once per channel chunk instead of once per channel (the reference
// re-reads it C times) and the random 4-byte gathers hit L1/L2-resident feature rows.
#include <cstdlib>
#include "common.cuh"
#include "pull.cuh"

namespace pdgn {

constexpr int GT = 256;  // threads per CTA

__device__ __forceinline__ void st_stream4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }
__device__ __forceinline__ float4 ld_stream4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }

// ---------------------------------------------------------------- grouping forward
// out[b,ch,e] = points[b,ch,idx[b,e]]   (e = j*k+s flattened, mk = m*k)
template <bool VEC>
__global__ void __launch_bounds__(GT) group_fwd_kernel(const float* __restrict__ points, const int* __restrict__ idx, int c,
                                                      int n, int mk, int cpb, float* __restrict__ out) {
    const int bz = blockIdx.z;
    const int c0 = blockIdx.y * cpb, c1 = min(c, c0 + cpb);
    const long long e = ((long long)blockIdx.x * GT + threadIdx.x) * (VEC ? 4 : 1);
    if (e >= mk) return;
    const int* ip = idx + (size_t)bz * mk + e;
    if (VEC) {
        const int4 id = *reinterpret_cast<const int4*>(ip);
        const float* src = points + ((size_t)bz * c + c0) * n;
        float* dst = out + ((size_t)bz * c + c0) * mk + e;
#pragma unroll 4
        for (int ch = c0; ch < c1; ++ch, src += n, dst += mk)
            st_stream4(dst, make_float4(__ldg(src + id.x), __ldg(src + id.y), __ldg(src + id.z), __ldg(src + id.w)));
    } else {
        const int id = *ip;
        for (int ch = c0; ch < c1; ++ch) out[((size_t)bz * c + ch) * mk + e] = __ldg(points + ((size_t)bz * c + ch) * n + id);
    }
}

// ---------------------------------------------------------------- grouping backward
// grad_points[b,ch,idx[b,e]] += grad_out[b,ch,e]   (FP32 RED.ADD, like the reference's atomicAdd)
template <bool VEC>
__global__ void __launch_bounds__(GT) group_bwd_kernel(const float* __restrict__ grad_out, const int* __restrict__ idx, int c,
                                                      int n, int mk, int cpb, float* __restrict__ grad_points) {
    const int bz = blockIdx.z;
    const int c0 = blockIdx.y * cpb, c1 = min(c, c0 + cpb);
    const long long e = ((long long)blockIdx.x * GT + threadIdx.x) * (VEC ? 4 : 1);
    if (e >= mk) return;
    const int* ip = idx + (size_t)bz * mk + e;
    if (VEC) {
        const int4 id = *reinterpret_cast<const int4*>(ip);
        const float* src = grad_out + ((size_t)bz * c + c0) * mk + e;
        float* dst = grad_points + ((size_t)bz * c + c0) * n;
#pragma unroll 4
        for (int ch = c0; ch < c1; ++ch, src += mk, dst += n) {
            const float4 g = ld_stream4(src);
            atomicAdd(dst + id.x, g.x);
            atomicAdd(dst + id.y, g.y);
            atomicAdd(dst + id.z, g.z);
            atomicAdd(dst + id.w, g.w);
        }
    } else {
        const int id = *ip;
        for (int ch = c0; ch < c1; ++ch)
            atomicAdd(grad_points + ((size_t)bz * c + ch) * n + id, grad_out[((size_t)bz * c + ch) * mk + e]);
    }
}

// ---------------------------------------------------------------- grouping, shared-memory staged (C >= 8)
// The random 4-byte gathers of the kernels above are served by L1 at ~10+ cycles per warp load; for feature-sized C that,
// not HBM, is the limit (3.3 TB/s = 51 % of the measured copy peak at B=35, C=256, n=m=1024, k=10).  Staging the CC
// feature rows of one (batch, channel chunk) in shared memory turns them into 32-bank gathers, leaving the streaming of
// the [B,C,m,k] tensor as the only HBM traffic.  Small row chunks (16 KB => ~12 CTAs/SM) and a software-pipelined index
// load keep enough stores in flight to cover the index-load -> LDS -> store latency chain.
static int gs_row_bytes() {  // shared budget for staged rows per CTA (tuning hook: PDGN_GS_ROW_KB)
    static int v = 0;
    if (v == 0) {
        const char* e = tune_env("PDGN_GS_ROW_KB");
        v = (e ? atoi(e) : 16) * 1024;  // 16 KB: 84.5 % of HBM at the C=256 stress shape (64 KB: 75 %), profiles/r01_gather_tune.txt
    }
    return v;
}
#define GS_ROW_BYTES gs_row_bytes()

__global__ void __launch_bounds__(GT) group_fwd_smem_kernel(const float* __restrict__ points, const int* __restrict__ idx, int c,
                                                           int n, int mk, int cc_max, int jpart, float* __restrict__ out) {
    extern __shared__ __align__(16) float rows[];  // [cc][n]
    const int bz = blockIdx.z;
    const int c0 = blockIdx.y * cc_max, cc = min(cc_max, c - c0);
    const float* src = points + ((size_t)bz * c + c0) * n;  // cc consecutive rows are one contiguous block
    if (((cc * n) & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        for (int e = threadIdx.x; e < (cc * n) >> 2; e += GT) reinterpret_cast<float4*>(rows)[e] = __ldg(reinterpret_cast<const float4*>(src) + e);
    } else {
        for (int e = threadIdx.x; e < cc * n; e += GT) rows[e] = __ldg(src + e);
    }
    __syncthreads();
    const long long e_end = min((long long)mk, (long long)(blockIdx.x + 1) * jpart);
    const int* ip = idx + (size_t)bz * mk;
    float* dst0 = out + ((size_t)bz * c + c0) * mk;
    long long e = (long long)blockIdx.x * jpart + threadIdx.x * 4;
    int4 id = e < e_end ? *reinterpret_cast<const int4*>(ip + e) : make_int4(0, 0, 0, 0);
    for (; e < e_end; e += GT * 4) {
        const long long en = e + GT * 4;  // software-pipelined index load: the next quad is in flight while this one is gathered
        const int4 idn = en < e_end ? *reinterpret_cast<const int4*>(ip + en) : make_int4(0, 0, 0, 0);
        const float* r = rows;
        float* dst = dst0 + e;
#pragma unroll 4
        for (int ch = 0; ch < cc; ++ch, r += n, dst += mk) st_stream4(dst, make_float4(r[id.x], r[id.y], r[id.z], r[id.w]));
        id = idn;
    }
}

// Inverse index of idx[b] (values in [0,n), mk entries): offs[b][n+1], pos[b][mk] with every list in ascending
// position order, so the pull kernels add each target's contributions in a fixed order (deterministic, unlike the
// reference's float atomics).  The fill is a STABLE counting sort with no block-wide barrier inside its loops: the entry
// list is cut into one contiguous slice per warp and every warp keeps its OWN 16-bit counter row (cnt[w][target]); a column
// prefix over the rows turns the counts into each warp's first slot inside a target's list, and the fill pass then only
// needs warp-level ordering (__match_any_sync ranks equal targets by lane, the group leader advances the warp's cursor).
// Equal targets keep their position order however skewed the index distribution is (feature-space kNN graphs have hubs
// with thousands of incoming edges).  Round 1's version took the warps of a 1024-entry tile in turn (32 barriers per
// tile): 25-38 us per call; this one is bounded by the three passes over idx.
constexpr int CSR_T = 1024;

__host__ __device__ inline int csr_warps(int n) {   // counter rows that fit ~200 KB (power of two, <= 32)
    int w = 32;
    while (w > 1 && (size_t)w * (size_t)((n + 1) & ~1) * 3 > 160 * 1024) w >>= 1;   // 16-bit counter + 8-bit owner per (warp, target)
    return w;
}
static size_t csr_smem_bytes(int n) { return (size_t)csr_warps(n) * (size_t)((n + 1) & ~1) * 3 + (size_t)n * 4 + 32 * 4 + 16; }
// staging area for the lists (mk ints) when it fits beside the counters
static bool csr_stage(int n, long long mk) { return csr_smem_bytes(n) + (size_t)mk * 4 <= 200 * 1024; }
static size_t csr_smem_total(int n, long long mk) { return csr_smem_bytes(n) + (csr_stage(n, mk) ? (size_t)mk * 4 : 0); }
// shapes the kernel takes: counters of a (warp, target) pair are 16 bit, the counter block must fit shared memory
static bool csr_ok(int n, long long mk) {
    const int w = csr_warps(n);
    const long long per = ((mk + w - 1) / w + 31) / 32 * 32;
    return csr_smem_bytes(n) <= 200 * 1024 && per <= 65535 && mk <= 0x7fffffffLL;
}

template <typename IdxT>
__global__ void __launch_bounds__(CSR_T) csr_build_kernel(const IdxT* __restrict__ idx, int n, int mk, int* __restrict__ offs,
                                                         int* __restrict__ pos, int stage_pos) {
    extern __shared__ __align__(16) unsigned char csr_raw[];
    const int W = csr_warps(n), np = (n + 1) & ~1;
    unsigned short* cnt = reinterpret_cast<unsigned short*>(csr_raw);                 // [W][np]
    int* base = reinterpret_cast<int*>(csr_raw + (size_t)W * np * 2);               // [n] first slot of each target's list
    int* wsum = base + n;                                                            // [32]
    unsigned char* owner = reinterpret_cast<unsigned char*>(wsum + 32);              // [W][np] duplicate detection, fill pass
    // stage_pos: the lists are assembled in shared memory and written out coalesced.  Scattered 4-byte global stores from ONE SM
    // per batch element (32 sectors per warp store) were the fill pass's bound: 6 of the kernel's 11 us.
    int* pos_s = reinterpret_cast<int*>(csr_raw + (((size_t)W * np * 3 + (size_t)n * 4 + 32 * 4 + 15) & ~(size_t)15));
    const int bz = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const IdxT* ip = idx + (size_t)bz * mk;
    int* ob = offs + (size_t)bz * (n + 1);
    int* pb = pos + (size_t)bz * mk;
    {
        unsigned* z = reinterpret_cast<unsigned*>(cnt);
        for (int i = t; i < W * np / 2; i += CSR_T) z[i] = 0u;
    }
    __syncthreads();
    // slice of warp w (w < W): entries [w * per, (w + 1) * per), per a multiple of 32
    const int per = ((mk + W - 1) / W + 31) / 32 * 32;
    const int e_lo = min(mk, warp * per), e_hi = warp < W ? min(mk, e_lo + per) : e_lo;
    unsigned* cw = reinterpret_cast<unsigned*>(cnt + (size_t)(warp < W ? warp : 0) * np);
    // The slice is read from global memory ONCE, all loads in flight together, and kept in registers for the fill pass
    // (slices of up to CSR_RC * 32 entries; the fill loop used to issue one dependent load per 32 entries: ~5 us of latency).
    constexpr int CSR_RC = 24;
    const bool cached = e_hi - e_lo <= CSR_RC * 32;
    int tc[CSR_RC];
#pragma unroll
    for (int i = 0; i < CSR_RC; ++i) {
        const int e = e_lo + lane + 32 * i;
        tc[i] = (cached && e < e_hi) ? (int)ip[e] : -1;
    }
    if (cached) {
#pragma unroll
        for (int i = 0; i < CSR_RC; ++i)
            if (tc[i] >= 0) atomicAdd(&cw[tc[i] >> 1], 1u << ((tc[i] & 1) * 16));   // integer counts: order irrelevant
    } else {
        for (int e = e_lo + lane; e < e_hi; e += 32) {
            const int tgt = (int)ip[e];
            atomicAdd(&cw[tgt >> 1], 1u << ((tgt & 1) * 16));
        }
    }
    __syncthreads();
    // column prefix: cnt[w][p] <- entries of target p in the slices before w; base[p] <- total for now
    for (int p = t; p < n; p += CSR_T) {
        int run = 0;
        for (int w = 0; w < W; ++w) {
            const int v = cnt[(size_t)w * np + p];
            cnt[(size_t)w * np + p] = (unsigned short)run;
            run += v;
        }
        base[p] = run;
    }
    __syncthreads();
    // exclusive scan of the totals: each thread owns a contiguous slice of targets
    const int tper = (n + CSR_T - 1) / CSR_T;
    const int lo = min(n, t * tper), hi = min(n, lo + tper);
    int local = 0;
    for (int p = lo; p < hi; ++p) local += base[p];
    int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (t < 32) {
        int v = wsum[t];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(kFull, v, o);
            if (t >= o) v += u;
        }
        wsum[t] = v;  // inclusive over warps
    }
    __syncthreads();
    int run = incl - local + (warp ? wsum[warp - 1] : 0);
    for (int p = lo; p < hi; ++p) {
        const int v = base[p];
        base[p] = run;
        ob[p] = run;
        run += v;
    }
    if (t == CSR_T - 1) ob[n] = mk;
    __syncthreads();
    // fill: the warp walks its slice in order, 32 entries at a time
    const unsigned lt = (1u << lane) - 1u;
    unsigned short* cur = cnt + (size_t)(warp < W ? warp : 0) * np;
    // Rank of a lane among the lanes of its 32-entry group with the same target.  __match_any_sync costs ~1000 cycles when the
    // 32 targets are (nearly) all different -- the common case, and 70 % of this kernel's time in the first version -- so the
    // lanes find out whether they share a target by writing their lane number into a per-warp byte array indexed by target
    // and reading it back: a lane that reads another lane's number shares its target.  Only those targets (usually none or one,
    // or one hub) are ranked, one ballot each.
    unsigned char* own = owner + (size_t)(warp < W ? warp : 0) * np;
    auto place = [&](int e, bool valid, int tgt_in) {
        const int tgt = valid ? tgt_in : 0;
        if (valid) own[tgt] = (unsigned char)lane;
        __syncwarp();
        const bool lost = valid && own[tgt] != (unsigned char)lane;
        unsigned dm = __ballot_sync(kFull, lost);
        unsigned peers = valid ? (1u << lane) : 0u;                      // alone unless shown otherwise
        while (dm) {                                                      // warp-uniform: one round per target that occurs twice or more
            const int src = __ffs((int)dm) - 1;
            const int tdup = __shfl_sync(kFull, tgt, src);
            const unsigned same = __ballot_sync(kFull, valid && tgt == tdup);
            if (valid && tgt == tdup) peers = same;
            dm &= ~same;
        }
        const int rank = __popc(peers & lt);
        int first = 0;
        if (valid && rank == 0) {                                         // one leader per distinct target of the group
            first = cur[tgt];
            cur[tgt] = (unsigned short)(first + __popc(peers));
        }
        first = __shfl_sync(kFull, first, peers ? __ffs((int)peers) - 1 : lane);
        if (valid) {
            if (stage_pos) pos_s[base[tgt] + first + rank] = e;
            else pb[base[tgt] + first + rank] = e;
        }
        __syncwarp();
    };
    if (cached) {
#pragma unroll
        for (int i = 0; i < CSR_RC; ++i) {
            const int e0 = e_lo + 32 * i;
            if (e0 >= e_hi) break;                                        // warp-uniform
            place(e0 + lane, tc[i] >= 0, tc[i]);
        }
    } else {
        for (int e0 = e_lo; e0 < e_hi; e0 += 32) {
            const int e = e0 + lane;
            place(e, e < e_hi, e < e_hi ? (int)ip[e] : 0);
        }
    }
    if (stage_pos) {
        __syncthreads();
        for (int i = t; i < mk; i += CSR_T) pb[i] = pos_s[i];
    }
}

// Shared skeleton of the three pull kernels.  `value(ch, e)` = contribution of entry e to channel ch (ch < cc is the
// caller's business: rows beyond cc are zero-filled), `extra(ch, p)` = per-target term added first, `out(ch, p)` = address.
// Every (channel, target) is written by exactly one thread with one RED.ADD onto the caller's buffer => deterministic.
template <class V, class X, class O>
__device__ __forceinline__ void pull_targets(const int* __restrict__ ob, const int* __restrict__ pb, int ntargets, int cc,
                                             int* nlong, int* longlist, V value, X extra, O out) {
    for (int p = threadIdx.x; p < ntargets; p += blockDim.x) {
        const int a = ob[p], b = ob[p + 1];
        float acc[PULL_CC];
#pragma unroll
        for (int ch = 0; ch < PULL_CC; ++ch) acc[ch] = ch < cc ? extra(ch, p) : 0.f;
        bool deferred = false;
        if (b - a > PULL_LONG) {
            const int slot = atomicAdd(nlong, 1);
            if (slot < 512) {
                longlist[slot] = p;
                deferred = true;
            }
        }
        if (!deferred) pull_list<PULL_CC>(pb, a, b, acc, value);
#pragma unroll
        for (int ch = 0; ch < PULL_CC; ++ch)
            if (ch < cc) atomicAdd(out(ch, p), acc[ch]);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int nl = min(*nlong, 512);
    for (int i = warp; i < nl; i += nw) {
        const int p = longlist[i];
        float acc[PULL_CC];
#pragma unroll
        for (int ch = 0; ch < PULL_CC; ++ch) acc[ch] = 0.f;
        pull_list_warp<PULL_CC>(pb, ob[p], ob[p + 1], lane, acc, value);
        if (lane == 0) {
#pragma unroll
            for (int ch = 0; ch < PULL_CC; ++ch)
                if (ch < cc) atomicAdd(out(ch, p), acc[ch]);
        }
    }
}

// grad_points[b,ch,p] += sum over the positions e with idx[b,e] == p of grad_out[b,ch,e]: the rows of one (batch,
// channel chunk) are streamed once into shared memory, every thread pulls the lists of its targets from there.
__global__ void __launch_bounds__(512) group_bwd_pull_kernel(const float* __restrict__ grad_out, const int* __restrict__ offs,
                                                            const int* __restrict__ pos, int c, int n, int mk, int cc_max,
                                                            float* __restrict__ grad_points) {
    extern __shared__ __align__(16) float rows[];  // [PULL_CC][mk] (rows beyond cc zero-filled)
    __shared__ int nlong, longlist[512];
    const int bz = blockIdx.y;
    const int c0 = blockIdx.x * cc_max, cc = min(cc_max, c - c0);
    const float* src = grad_out + ((size_t)bz * c + c0) * mk;
    if (threadIdx.x == 0) nlong = 0;
    if ((mk & 3) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* r4 = reinterpret_cast<float4*>(rows);
        for (int e = threadIdx.x; e < cc * (mk >> 2); e += blockDim.x) r4[e] = __ldcs(s4 + e);
    } else {
        for (int e = threadIdx.x; e < cc * mk; e += blockDim.x) rows[e] = __ldcs(src + e);
    }
    for (int e = cc * mk + threadIdx.x; e < cc_max * mk; e += blockDim.x) rows[e] = 0.f;
    __syncthreads();
    float* dst = grad_points + ((size_t)bz * c + c0) * n;
    const int cm = cc_max;
    pull_targets(offs + (size_t)bz * (n + 1), pos + (size_t)bz * mk, n, cc, &nlong, longlist,
                 [&](int ch, int e) { return ch < cm ? rows[(size_t)ch * mk + e] : 0.f; },
                 [&](int, int) { return 0.f; },
                 [&](int ch, int p) { return dst + (size_t)ch * n + p; });
}

// ---------------------------------------------------------------- interpolation forward
// out[b,ch,j] = fma(w2,p[i2], fma(w0,p[i0], w1*p[i1]))  -- the compiled order of
// weight[0]*points[idx[0]] + weight[1]*points[idx[1]] + weight[2]*points[idx[2]] (interpolation_cuda_kernel.cu:194).
__device__ __forceinline__ float interp3(const float* __restrict__ src, int i0, int i1, int i2, float w0, float w1, float w2) {
    return __fmaf_rn(w2, __ldg(src + i2), __fmaf_rn(w0, __ldg(src + i0), __fmul_rn(w1, __ldg(src + i1))));
}

template <bool VEC>
__global__ void __launch_bounds__(GT) interp_fwd_kernel(const float* __restrict__ points, const int* __restrict__ idx,
                                                       const float* __restrict__ weight, int c, int m, int n, int cpb,
                                                       float* __restrict__ out) {
    const int bz = blockIdx.z;
    const int c0 = blockIdx.y * cpb, c1 = min(c, c0 + cpb);
    const long long j = ((long long)blockIdx.x * GT + threadIdx.x) * (VEC ? 4 : 1);
    if (j >= n) return;
    const int* ip = idx + ((size_t)bz * n + j) * 3;
    const float* wp = weight + ((size_t)bz * n + j) * 3;
    if (VEC) {
        int id[12];
        float w[12];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int4 a = reinterpret_cast<const int4*>(ip)[q];
            const float4 f = reinterpret_cast<const float4*>(wp)[q];
            id[4 * q] = a.x; id[4 * q + 1] = a.y; id[4 * q + 2] = a.z; id[4 * q + 3] = a.w;
            w[4 * q] = f.x; w[4 * q + 1] = f.y; w[4 * q + 2] = f.z; w[4 * q + 3] = f.w;
        }
        const float* src = points + ((size_t)bz * c + c0) * m;
        float* dst = out + ((size_t)bz * c + c0) * n + j;
#pragma unroll 2
        for (int ch = c0; ch < c1; ++ch, src += m, dst += n) {
            float4 o;
            o.x = interp3(src, id[0], id[1], id[2], w[0], w[1], w[2]);
            o.y = interp3(src, id[3], id[4], id[5], w[3], w[4], w[5]);
            o.z = interp3(src, id[6], id[7], id[8], w[6], w[7], w[8]);
            o.w = interp3(src, id[9], id[10], id[11], w[9], w[10], w[11]);
            st_stream4(dst, o);
        }
    } else {
        const int i0 = ip[0], i1 = ip[1], i2 = ip[2];
        const float w0 = wp[0], w1 = wp[1], w2 = wp[2];
        for (int ch = c0; ch < c1; ++ch)
            out[((size_t)bz * c + ch) * n + j] = interp3(points + ((size_t)bz * c + ch) * m, i0, i1, i2, w0, w1, w2);
    }
}

// grad_points[b,ch,idx_t] += grad_out[b,ch,j] * w_t   (each product rounded, then RED.ADD; reference :90-114)
__global__ void __launch_bounds__(GT) interp_bwd_kernel(const float* __restrict__ grad_out, const int* __restrict__ idx,
                                                       const float* __restrict__ weight, int c, int n, int m, int cpb,
                                                       float* __restrict__ grad_points) {
    const int bz = blockIdx.z;
    const int c0 = blockIdx.y * cpb, c1 = min(c, c0 + cpb);
    const long long j = (long long)blockIdx.x * GT + threadIdx.x;
    if (j >= n) return;
    const int* ip = idx + ((size_t)bz * n + j) * 3;
    const float* wp = weight + ((size_t)bz * n + j) * 3;
    const int i0 = ip[0], i1 = ip[1], i2 = ip[2];
    const float w0 = wp[0], w1 = wp[1], w2 = wp[2];
    const float* src = grad_out + ((size_t)bz * c + c0) * n + j;
    float* dst = grad_points + ((size_t)bz * c + c0) * m;
#pragma unroll 4
    for (int ch = c0; ch < c1; ++ch, src += n, dst += m) {
        const float g = __ldcs(src);
        atomicAdd(dst + i0, __fmul_rn(g, w0));
        atomicAdd(dst + i1, __fmul_rn(g, w1));
        atomicAdd(dst + i2, __fmul_rn(g, w2));
    }
}

// ---------------------------------------------------------------- edge features
// ee[b,ch,i,s] = x[b,ch,i] ; ee[b,c+ch,i,s] = x[b,ch,idx[b,i,s]] - x[b,ch,i]      (PDGNet_v2.py:470-475)
template <bool VEC>
__global__ void __launch_bounds__(GT) edge_fwd_kernel(const float* __restrict__ x, const long long* __restrict__ idx, int c,
                                                     int n, int k, int cpb, float* __restrict__ ee) {
    const int bz = blockIdx.z;
    const int c0 = blockIdx.y * cpb, c1 = min(c, c0 + cpb);
    const long long nk = (long long)n * k;
    const long long e = ((long long)blockIdx.x * GT + threadIdx.x) * (VEC ? 4 : 1);
    if (e >= nk) return;
    const long long* ip = idx + (size_t)bz * nk + e;
    constexpr int V = VEC ? 4 : 1;
    int nb[V], ce[V];
#pragma unroll
    for (int q = 0; q < V; ++q) {
        nb[q] = (int)ip[q];
        ce[q] = (int)((e + q) / k);
    }
    for (int ch = c0; ch < c1; ++ch) {
        const float* src = x + ((size_t)bz * c + ch) * n;
        float cen[V], dif[V];
#pragma unroll
        for (int q = 0; q < V; ++q) {
            cen[q] = __ldg(src + ce[q]);
            dif[q] = __fsub_rn(__ldg(src + nb[q]), cen[q]);
        }
        float* d0 = ee + ((size_t)bz * 2 * c + ch) * nk + e;
        float* d1 = ee + ((size_t)bz * 2 * c + c + ch) * nk + e;
        if (VEC) {
            st_stream4(d0, make_float4(cen[0], cen[1 % V], cen[2 % V], cen[3 % V]));
            st_stream4(d1, make_float4(dif[0], dif[1 % V], dif[2 % V], dif[3 % V]));
        } else {
            d0[0] = cen[0];
            d1[0] = dif[0];
        }
    }
}

// Shared-memory staged form (C >= 8), like group_fwd_smem_kernel: the cc feature rows of one (batch, channel chunk) sit in shared
// memory, so the random neighbour gathers and the per-point centre reads are 32-bank LDS instead of L1/L2 gathers, and the two
// [B,2C,N,k] halves stream out as 16-byte stores.  Index loads (int64, 32 bytes per thread) are software-pipelined.
__global__ void __launch_bounds__(GT) edge_fwd_smem_kernel(const float* __restrict__ x, const long long* __restrict__ idx, int c, int n,
                                                          int k, int cc_max, int jpart, float* __restrict__ ee) {
    extern __shared__ __align__(16) float rows[];  // [cc][n]
    const int bz = blockIdx.z;
    const int c0 = blockIdx.y * cc_max, cc = min(cc_max, c - c0);
    const int nk = n * k;
    const float* src = x + ((size_t)bz * c + c0) * n;
    if (((cc * n) & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        for (int e = threadIdx.x; e < (cc * n) >> 2; e += GT) reinterpret_cast<float4*>(rows)[e] = __ldg(reinterpret_cast<const float4*>(src) + e);
    } else {
        for (int e = threadIdx.x; e < cc * n; e += GT) rows[e] = __ldg(src + e);
    }
    __syncthreads();
    const long long e_end = min((long long)nk, (long long)(blockIdx.x + 1) * jpart);
    const long long* ip = idx + (size_t)bz * nk;
    float* d0b = ee + ((size_t)bz * 2 * c + c0) * nk;          // central half
    float* d1b = d0b + (size_t)c * nk;                          // neighbour - central half
    long long e = (long long)blockIdx.x * jpart + threadIdx.x * 4;
    longlong2 ia = make_longlong2(0, 0), ib = make_longlong2(0, 0);
    if (e < e_end) {
        ia = *reinterpret_cast<const longlong2*>(ip + e);
        ib = *reinterpret_cast<const longlong2*>(ip + e + 2);
    }
    for (; e < e_end; e += GT * 4) {
        const long long en = e + GT * 4;
        longlong2 na = make_longlong2(0, 0), nb2 = make_longlong2(0, 0);
        if (en < e_end) {
            na = *reinterpret_cast<const longlong2*>(ip + en);
            nb2 = *reinterpret_cast<const longlong2*>(ip + en + 2);
        }
        const int nb0 = (int)ia.x, nb1 = (int)ia.y, nb2i = (int)ib.x, nb3 = (int)ib.y;
        const int ce0 = (int)(e / k), ce1 = (int)((e + 1) / k), ce2 = (int)((e + 2) / k), ce3 = (int)((e + 3) / k);
        const float* r = rows;
        float* d0 = d0b + e;
        float* d1 = d1b + e;
#pragma unroll 4
        for (int ch = 0; ch < cc; ++ch, r += n, d0 += nk, d1 += nk) {
            const float a0 = r[ce0], a1 = r[ce1], a2 = r[ce2], a3 = r[ce3];
            st_stream4(d0, make_float4(a0, a1, a2, a3));
            st_stream4(d1, make_float4(__fsub_rn(r[nb0], a0), __fsub_rn(r[nb1], a1), __fsub_rn(r[nb2i], a2), __fsub_rn(r[nb3], a3)));
        }
        ia = na;
        ib = nb2;
    }
}

// grad_x[b,ch,i] += sum_s (g_central[i,s] - g_nbr[i,s]) ;  grad_x[b,ch,idx[i,s]] += g_nbr[i,s]
__global__ void __launch_bounds__(GT) edge_bwd_kernel(const float* __restrict__ grad_ee, const long long* __restrict__ idx, int c,
                                                     int n, int k, int cpb, float* __restrict__ grad_x) {
    const int bz = blockIdx.z;
    const int c0 = blockIdx.y * cpb, c1 = min(c, c0 + cpb);
    const int i = blockIdx.x * GT + threadIdx.x;
    if (i >= n) return;
    const long long nk = (long long)n * k;
    const long long* ip = idx + (size_t)bz * nk + (long long)i * k;
    for (int ch = c0; ch < c1; ++ch) {
        const float* g0 = grad_ee + ((size_t)bz * 2 * c + ch) * nk + (long long)i * k;
        const float* g1 = grad_ee + ((size_t)bz * 2 * c + c + ch) * nk + (long long)i * k;
        float* dst = grad_x + ((size_t)bz * c + ch) * n;
        float acc = 0.f;
        for (int s = 0; s < k; ++s) {
            const float gn = __ldcs(g1 + s);
            acc += __ldcs(g0 + s) - gn;
            atomicAdd(dst + (int)ip[s], gn);
        }
        atomicAdd(dst + i, acc);
    }
}

// ---------------------------------------------------------------- interpolation / edge features, shared-memory staged
// Same two ideas as the staged grouping kernels: gathers out of shared rows, backward as a deterministic pull
// through the inverse index (csr_build_kernel) instead of float atomics.

// out[b,ch,j] for 4 consecutive j per thread; rows = cc feature rows of m floats.
__global__ void __launch_bounds__(GT) interp_fwd_smem_kernel(const float* __restrict__ points, const int* __restrict__ idx,
                                                            const float* __restrict__ weight, int c, int m, int n, int cc_max,
                                                            int jpart, float* __restrict__ out) {
    extern __shared__ __align__(16) float rows[];  // [cc][m]
    const int bz = blockIdx.z;
    const int c0 = blockIdx.y * cc_max, cc = min(cc_max, c - c0);
    const float* src = points + ((size_t)bz * c + c0) * m;
    if (((cc * m) & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        for (int e = threadIdx.x; e < (cc * m) >> 2; e += GT) reinterpret_cast<float4*>(rows)[e] = __ldg(reinterpret_cast<const float4*>(src) + e);
    } else {
        for (int e = threadIdx.x; e < cc * m; e += GT) rows[e] = __ldg(src + e);
    }
    __syncthreads();
    const int j_end = min(n, (int)(blockIdx.x + 1) * jpart);
    float* dst0 = out + ((size_t)bz * c + c0) * n;
    for (int j = blockIdx.x * jpart + threadIdx.x * 4; j < j_end; j += GT * 4) {
        const int* ip = idx + ((size_t)bz * n + j) * 3;
        const float* wp = weight + ((size_t)bz * n + j) * 3;
        int id[12];
        float w[12];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int4 a = reinterpret_cast<const int4*>(ip)[q];
            const float4 f = reinterpret_cast<const float4*>(wp)[q];
            id[4 * q] = a.x; id[4 * q + 1] = a.y; id[4 * q + 2] = a.z; id[4 * q + 3] = a.w;
            w[4 * q] = f.x; w[4 * q + 1] = f.y; w[4 * q + 2] = f.z; w[4 * q + 3] = f.w;
        }
        const float* r = rows;
        float* dst = dst0 + j;
#pragma unroll 2
        for (int ch = 0; ch < cc; ++ch, r += m, dst += n) {
            float4 o;
            o.x = __fmaf_rn(w[2], r[id[2]], __fmaf_rn(w[0], r[id[0]], __fmul_rn(w[1], r[id[1]])));
            o.y = __fmaf_rn(w[5], r[id[5]], __fmaf_rn(w[3], r[id[3]], __fmul_rn(w[4], r[id[4]])));
            o.z = __fmaf_rn(w[8], r[id[8]], __fmaf_rn(w[6], r[id[6]], __fmul_rn(w[7], r[id[7]])));
            o.w = __fmaf_rn(w[11], r[id[11]], __fmaf_rn(w[9], r[id[9]], __fmul_rn(w[10], r[id[10]])));
            st_stream4(dst, o);
        }
    }
}

// grad_points[b,ch,p] += sum over entries e = 3*j+t with idx[b,e] == p of fmul(grad_out[b,ch,j], weight[b,e])
// (each product rounded like the reference's atomicAdd operand; added in ascending e: deterministic).
__global__ void __launch_bounds__(512) interp_bwd_pull_kernel(const float* __restrict__ grad_out, const float* __restrict__ weight,
                                                             const int* __restrict__ offs, const int* __restrict__ pos, int c, int n,
                                                             int m, int cc_max, float* __restrict__ grad_points) {
    extern __shared__ __align__(16) float rows[];  // [cc_max][n] grad_out rows (beyond cc zero-filled), then [3n] weights
    __shared__ int nlong, longlist[512];
    const int bz = blockIdx.y;
    const int c0 = blockIdx.x * cc_max, cc = min(cc_max, c - c0);
    const float* src = grad_out + ((size_t)bz * c + c0) * n;
    if (threadIdx.x == 0) nlong = 0;
    for (int e = threadIdx.x; e < cc * n; e += blockDim.x) rows[e] = __ldcs(src + e);
    for (int e = cc * n + threadIdx.x; e < cc_max * n; e += blockDim.x) rows[e] = 0.f;
    float* wsm = rows + (size_t)cc_max * n;
    const float* wsrc = weight + (size_t)bz * n * 3;
    for (int e = threadIdx.x; e < 3 * n; e += blockDim.x) wsm[e] = __ldg(wsrc + e);
    __syncthreads();
    float* dst = grad_points + ((size_t)bz * c + c0) * m;
    const int cm = cc_max;
    pull_targets(offs + (size_t)bz * (m + 1), pos + (size_t)bz * 3 * n, m, cc, &nlong, longlist,
                 [&](int ch, int e) { return ch < cm ? __fmul_rn(rows[(size_t)ch * n + e / 3], wsm[e]) : 0.f; },
                 [&](int, int) { return 0.f; },
                 [&](int ch, int p) { return dst + (size_t)ch * m + p; });
}

// edge features backward as a pull: grad_x[b,ch,i] += sum_s (g0[i,s] - g1[i,s]) + sum over entries e with idx[b,e] == i of g1[e]
// (g0 = grad_ee[b,ch], g1 = grad_ee[b,c+ch], both [n*k]); the g1 rows are staged in shared memory, g0 is read in place.
__global__ void __launch_bounds__(512) edge_bwd_pull_kernel(const float* __restrict__ grad_ee, const int* __restrict__ offs,
                                                           const int* __restrict__ pos, int c, int n, int k, int cc_max,
                                                           float* __restrict__ grad_x) {
    extern __shared__ __align__(16) float rows[];  // [cc_max][n*k]: g1 per channel (beyond cc zero-filled)
    __shared__ int nlong, longlist[512];
    const int bz = blockIdx.y;
    const int c0 = blockIdx.x * cc_max, cc = min(cc_max, c - c0);
    const int nk = n * k;
    if (threadIdx.x == 0) nlong = 0;
    const float* s1 = grad_ee + ((size_t)bz * 2 * c + c + c0) * nk;  // cc consecutive g1 rows are contiguous
    if ((nk & 3) == 0) {
        for (int e = threadIdx.x; e < cc * (nk >> 2); e += blockDim.x) reinterpret_cast<float4*>(rows)[e] = __ldcs(reinterpret_cast<const float4*>(s1) + e);
    } else {
        for (int e = threadIdx.x; e < cc * nk; e += blockDim.x) rows[e] = __ldcs(s1 + e);
    }
    for (int e = cc * nk + threadIdx.x; e < cc_max * nk; e += blockDim.x) rows[e] = 0.f;
    __syncthreads();
    const float* g0b = grad_ee + ((size_t)bz * 2 * c + c0) * nk;
    float* dst = grad_x + ((size_t)bz * c + c0) * n;
    const int cm = cc_max;
    pull_targets(offs + (size_t)bz * (n + 1), pos + (size_t)bz * nk, n, cc, &nlong, longlist,
                 [&](int ch, int e) { return ch < cm ? rows[(size_t)ch * nk + e] : 0.f; },
                 [&](int ch, int i) {
                     const float* g0 = g0b + (size_t)ch * nk + (size_t)i * k;
                     const float* r1 = rows + (size_t)ch * nk + (size_t)i * k;
                     float acc = 0.f;
                     for (int s = 0; s < k; ++s) acc += __ldcs(g0 + s) - r1[s];
                     return acc;
                 },
                 [&](int ch, int p) { return dst + (size_t)ch * n + p; });
}

// channels per CTA: keep >= ~4 waves of CTAs while amortising the index read over as many channels as possible
static int pick_cpb(long long ctas_per_channel_group, int c) {
    int cpb = c;
    while (cpb > 4 && ctas_per_channel_group * ((c + cpb - 1) / cpb) < 4 * 148 * 4) cpb = (cpb + 1) / 2;
    return cpb < 1 ? 1 : cpb;
}

}  // namespace pdgn

using namespace pdgn;

#define PDGN_GATHER_ARGS_OK(...) \
    if (!(__VA_ARGS__)) return PDGN_ERR_BAD_ARG

extern "C" int pdgn_group_fwd(const float* points, const int* idx, int b, int c, int n, int m, int k, float* out, void* stream) {
    PDGN_RANGE("pdgn_group_fwd");
    PDGN_GATHER_ARGS_OK(b >= 0 && c >= 0 && n >= 0 && m >= 0 && k >= 0);
    const long long mk = (long long)m * k;
    if (b == 0 || c == 0 || mk == 0) return PDGN_OK;  // empty output
    PDGN_GATHER_ARGS_OK(points && idx && out && n > 0);
    PDGN_VERIFY_IDX32(idx, (size_t)b * m * k, n, (cudaStream_t)stream);
    if (mk > 0x7fffffffLL || b > 65535) return PDGN_ERR_UNSUPPORTED;
    const bool vec = (mk % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(idx)) % 16 == 0);
    if (vec && c >= 8 && (size_t)n * 4 * 4 <= GS_ROW_BYTES) {
        // shared-memory staged path: cc rows of n floats per CTA, position range split so that >= ~4 waves of CTAs exist
        int cc = 4;
        while (cc * 2 <= c && (size_t)cc * 2 * n * 4 <= GS_ROW_BYTES) cc *= 2;
        const int chunks = (c + cc - 1) / cc;
        long long parts = (4LL * 148 * 3 + (long long)chunks * b - 1) / ((long long)chunks * b);
        const long long max_parts = (mk + GT * 4 - 1) / (GT * 4);
        if (parts > max_parts) parts = max_parts;
        if (parts < 1) parts = 1;
        long long jpart = (mk + parts - 1) / parts;
        jpart = (jpart + GT * 4 - 1) / (GT * 4) * (GT * 4);
        parts = (mk + jpart - 1) / jpart;
        const size_t smem = (size_t)cc * n * 4;
        PDGN_CUDA(cudaFuncSetAttribute(group_fwd_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((unsigned)parts, chunks, b);
        if (grid.y > 65535) return PDGN_ERR_UNSUPPORTED;
        group_fwd_smem_kernel<<<grid, GT, smem, (cudaStream_t)stream>>>(points, idx, c, n, (int)mk, cc, (int)jpart, out);
        PDGN_CHECK_LAUNCH();
        return PDGN_OK;
    }
    const long long per = vec ? 4 : 1;
    const unsigned gx = (unsigned)((mk + GT * per - 1) / (GT * per));
    const int cpb = pick_cpb((long long)gx * b, c);
    dim3 grid(gx, (c + cpb - 1) / cpb, b);
    if (grid.y > 65535) return PDGN_ERR_UNSUPPORTED;
    if (vec) group_fwd_kernel<true><<<grid, GT, 0, (cudaStream_t)stream>>>(points, idx, c, n, (int)mk, cpb, out);
    else group_fwd_kernel<false><<<grid, GT, 0, (cudaStream_t)stream>>>(points, idx, c, n, (int)mk, cpb, out);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}

// Workspace form of the backward: deterministic pull through an inverse index built in `workspace`
// (pdgn_group_bwd_workspace bytes).  Falls back to the atomic kernel when the rows do not fit shared memory.
extern "C" size_t pdgn_group_bwd_workspace(int b, int n, int m, int k) {
    if (b < 0 || n < 0 || m < 0 || k < 0) return 0;
    return ((size_t)b * ((size_t)n + 1) + (size_t)b * m * k) * sizeof(int) + 256;
}

extern "C" int pdgn_group_bwd_ws(const float* grad_out, const int* idx, int b, int c, int n, int m, int k, float* grad_points,
                                 void* workspace, size_t workspace_bytes, void* stream) {
    PDGN_RANGE("pdgn_group_bwd_ws");
    PDGN_GATHER_ARGS_OK(b >= 0 && c >= 0 && n >= 0 && m >= 0 && k >= 0);
    const long long mk = (long long)m * k;
    if (b == 0 || c == 0 || mk == 0) return PDGN_OK;  // nothing to add
    PDGN_GATHER_ARGS_OK(grad_out && idx && grad_points && n > 0);
    PDGN_VERIFY_IDX32(idx, (size_t)b * m * k, n, (cudaStream_t)stream);
    if (mk > 0x7fffffffLL || b > 65535) return PDGN_ERR_UNSUPPORTED;
    const size_t row_bytes = (size_t)mk * 4;
    const size_t csr_smem = csr_smem_total(n, mk);
    if (!workspace || c < 4 || row_bytes > 200 * 1024 || !csr_ok(n, mk))
        return pdgn_group_bwd(grad_out, idx, b, c, n, m, k, grad_points, stream);
    if (workspace_bytes < pdgn_group_bwd_workspace(b, n, m, k) - 256 || (reinterpret_cast<uintptr_t>(workspace) & 3)) return PDGN_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    int* offs = reinterpret_cast<int*>(workspace);
    int* pos = offs + (size_t)b * (n + 1);
    PDGN_CUDA(cudaFuncSetAttribute(csr_build_kernel<int>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csr_smem));
    csr_build_kernel<int><<<b, CSR_T, csr_smem, st>>>(idx, n, (int)mk, offs, pos, csr_stage(n, mk) ? 1 : 0);
    PDGN_CHECK_LAUNCH();
    bool launched = false;
    const int rc_stream = pull_stream_launch(0, grad_out, offs, pos, b, c, n, (int)mk, 0, grad_points, st, &launched, nullptr);
    if (rc_stream != PDGN_OK || launched) return rc_stream;
    int cc = 1;
    while (cc < 4 && (size_t)cc * 2 * row_bytes <= 100 * 1024) cc *= 2;
    const size_t smem = (size_t)cc * row_bytes;
    PDGN_CUDA(cudaFuncSetAttribute(group_bwd_pull_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((c + cc - 1) / cc, b);
    group_bwd_pull_kernel<<<grid, 512, smem, st>>>(grad_out, offs, pos, c, n, (int)mk, cc, grad_points);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}

extern "C" int pdgn_group_bwd(const float* grad_out, const int* idx, int b, int c, int n, int m, int k, float* grad_points,
                              void* stream) {
    PDGN_RANGE("pdgn_group_bwd");
    PDGN_GATHER_ARGS_OK(b >= 0 && c >= 0 && n >= 0 && m >= 0 && k >= 0);
    const long long mk = (long long)m * k;
    if (b == 0 || c == 0 || mk == 0) return PDGN_OK;  // nothing to add
    PDGN_GATHER_ARGS_OK(grad_out && idx && grad_points && n > 0);
    PDGN_VERIFY_IDX32(idx, (size_t)b * m * k, n, (cudaStream_t)stream);
    if (mk > 0x7fffffffLL || b > 65535) return PDGN_ERR_UNSUPPORTED;
    const bool vec = (mk % 4 == 0) && ((reinterpret_cast<uintptr_t>(grad_out) | reinterpret_cast<uintptr_t>(idx)) % 16 == 0);
    const long long per = vec ? 4 : 1;
    const unsigned gx = (unsigned)((mk + GT * per - 1) / (GT * per));
    const int cpb = pick_cpb((long long)gx * b, c);
    dim3 grid(gx, (c + cpb - 1) / cpb, b);
    if (grid.y > 65535) return PDGN_ERR_UNSUPPORTED;
    if (vec) group_bwd_kernel<true><<<grid, GT, 0, (cudaStream_t)stream>>>(grad_out, idx, c, n, (int)mk, cpb, grad_points);
    else group_bwd_kernel<false><<<grid, GT, 0, (cudaStream_t)stream>>>(grad_out, idx, c, n, (int)mk, cpb, grad_points);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}

extern "C" int pdgn_interp_fwd(const float* points, const int* idx, const float* weight, int b, int c, int m, int n, float* out,
                               void* stream) {
    PDGN_RANGE("pdgn_interp_fwd");
    PDGN_GATHER_ARGS_OK(b >= 0 && c >= 0 && m >= 0 && n >= 0);
    if (b == 0 || c == 0 || n == 0) return PDGN_OK;
    PDGN_GATHER_ARGS_OK(points && idx && weight && out && m > 0);
    PDGN_VERIFY_IDX32(idx, (size_t)b * n * 3, m, (cudaStream_t)stream);
    if (b > 65535) return PDGN_ERR_UNSUPPORTED;
    const bool vec = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(idx) |
                                       reinterpret_cast<uintptr_t>(weight)) % 16 == 0);
    if (vec && c >= 8 && (size_t)m * 4 * 4 <= GS_ROW_BYTES) {
        int cc = 4;
        while (cc * 2 <= c && (size_t)cc * 2 * m * 4 <= GS_ROW_BYTES) cc *= 2;
        const int chunks = (c + cc - 1) / cc;
        long long parts = (4LL * 148 * 3 + (long long)chunks * b - 1) / ((long long)chunks * b);
        const long long max_parts = ((long long)n + GT * 4 - 1) / (GT * 4);
        if (parts > max_parts) parts = max_parts;
        if (parts < 1) parts = 1;
        long long jpart = (n + parts - 1) / parts;
        jpart = (jpart + GT * 4 - 1) / (GT * 4) * (GT * 4);
        parts = (n + jpart - 1) / jpart;
        const size_t smem = (size_t)cc * m * 4;
        PDGN_CUDA(cudaFuncSetAttribute(interp_fwd_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((unsigned)parts, chunks, b);
        if (grid.y > 65535) return PDGN_ERR_UNSUPPORTED;
        interp_fwd_smem_kernel<<<grid, GT, smem, (cudaStream_t)stream>>>(points, idx, weight, c, m, n, cc, (int)jpart, out);
        PDGN_CHECK_LAUNCH();
        return PDGN_OK;
    }
    const int per = vec ? 4 : 1;
    const unsigned gx = (unsigned)((n + GT * per - 1) / (GT * per));
    const int cpb = pick_cpb((long long)gx * b, c);
    dim3 grid(gx, (c + cpb - 1) / cpb, b);
    if (grid.y > 65535) return PDGN_ERR_UNSUPPORTED;
    if (vec) interp_fwd_kernel<true><<<grid, GT, 0, (cudaStream_t)stream>>>(points, idx, weight, c, m, n, cpb, out);
    else interp_fwd_kernel<false><<<grid, GT, 0, (cudaStream_t)stream>>>(points, idx, weight, c, m, n, cpb, out);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}

extern "C" size_t pdgn_interp_bwd_workspace(int b, int n, int m) {
    if (b < 0 || n < 0 || m < 0) return 0;
    return ((size_t)b * ((size_t)m + 1) + (size_t)b * n * 3) * sizeof(int) + 256;
}

extern "C" int pdgn_interp_bwd_ws(const float* grad_out, const int* idx, const float* weight, int b, int c, int n, int m,
                                  float* grad_points, void* workspace, size_t workspace_bytes, void* stream) {
    PDGN_RANGE("pdgn_interp_bwd_ws");
    PDGN_GATHER_ARGS_OK(b >= 0 && c >= 0 && m >= 0 && n >= 0);
    if (b == 0 || c == 0 || n == 0) return PDGN_OK;
    PDGN_GATHER_ARGS_OK(grad_out && idx && weight && grad_points && m > 0);
    PDGN_VERIFY_IDX32(idx, (size_t)b * n * 3, m, (cudaStream_t)stream);
    if (b > 65535) return PDGN_ERR_UNSUPPORTED;
    const size_t csr_smem = csr_smem_total(m, (long long)n * 3);
    int cc = PULL_CC;
    while (cc > 1 && ((size_t)cc * n + 3 * (size_t)n) * 4 > 96 * 1024) cc >>= 1;
    const size_t smem = ((size_t)cc * n + 3 * (size_t)n) * 4;
    if (!workspace || c < 4 || smem > 200 * 1024 || !csr_ok(m, (long long)n * 3))
        return pdgn_interp_bwd(grad_out, idx, weight, b, c, n, m, grad_points, stream);
    if (workspace_bytes < pdgn_interp_bwd_workspace(b, n, m) - 256 || (reinterpret_cast<uintptr_t>(workspace) & 3)) return PDGN_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    int* offs = reinterpret_cast<int*>(workspace);
    int* pos = offs + (size_t)b * (m + 1);
    PDGN_CUDA(cudaFuncSetAttribute(csr_build_kernel<int>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csr_smem));
    csr_build_kernel<int><<<b, CSR_T, csr_smem, st>>>(idx, m, 3 * n, offs, pos, csr_stage(m, (long long)n * 3) ? 1 : 0);
    PDGN_CHECK_LAUNCH();
    {   // streaming form (TMA double-buffered rows, register-cached lists and weights): any shape it accepts
        bool launched = false;
        const int rc = pull_stream_launch(2, grad_out, offs, pos, b, c, m, n, 3, grad_points, st, &launched, weight);
        if (rc != PDGN_OK) return rc;
        if (launched) return PDGN_OK;
    }
    PDGN_CUDA(cudaFuncSetAttribute(interp_bwd_pull_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((c + cc - 1) / cc, b);
    interp_bwd_pull_kernel<<<grid, 512, smem, st>>>(grad_out, weight, offs, pos, c, n, m, cc, grad_points);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}

extern "C" int pdgn_interp_bwd(const float* grad_out, const int* idx, const float* weight, int b, int c, int n, int m,
                               float* grad_points, void* stream) {
    PDGN_RANGE("pdgn_interp_bwd");
    PDGN_GATHER_ARGS_OK(b >= 0 && c >= 0 && m >= 0 && n >= 0);
    if (b == 0 || c == 0 || n == 0) return PDGN_OK;
    PDGN_GATHER_ARGS_OK(grad_out && idx && weight && grad_points && m > 0);
    PDGN_VERIFY_IDX32(idx, (size_t)b * n * 3, m, (cudaStream_t)stream);
    if (b > 65535) return PDGN_ERR_UNSUPPORTED;
    const unsigned gx = (unsigned)((n + GT - 1) / GT);
    const int cpb = pick_cpb((long long)gx * b, c);
    dim3 grid(gx, (c + cpb - 1) / cpb, b);
    if (grid.y > 65535) return PDGN_ERR_UNSUPPORTED;
    interp_bwd_kernel<<<grid, GT, 0, (cudaStream_t)stream>>>(grad_out, idx, weight, c, n, m, cpb, grad_points);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}

extern "C" int pdgn_edge_feat_fwd(const float* x, const int64_t* idx, int b, int c, int n, int k, float* ee, void* stream) {
    PDGN_RANGE("pdgn_edge_feat_fwd");
    PDGN_GATHER_ARGS_OK(b >= 0 && c >= 0 && n >= 0 && k >= 0);
    const long long nk = (long long)n * k;
    if (b == 0 || c == 0 || nk == 0) return PDGN_OK;
    PDGN_GATHER_ARGS_OK(x && idx && ee);
    PDGN_VERIFY_IDX64(idx, (size_t)b * n * k, n, (cudaStream_t)stream);
    if (b > 65535 || nk > 0x7fffffffLL) return PDGN_ERR_UNSUPPORTED;
    const bool vec = (nk % 4 == 0) && (reinterpret_cast<uintptr_t>(ee) % 16 == 0);
    const long long* ip = reinterpret_cast<const long long*>(idx);
    // measured (tools/edge_fwd_time.py, B=35, k=10): C=256 N=1024 167 -> 143 us (5.4 TB/s); the generator's small stages are launch-
    // latency sized and stay on the plain kernel
    if (vec && c >= 8 && (long long)c * n >= 131072 && (size_t)n * 4 * 4 <= GS_ROW_BYTES && (reinterpret_cast<uintptr_t>(idx) % 16 == 0)) {
        // shared-memory staged path (same plan as pdgn_group_fwd): cc rows of n floats per CTA, positions split into parts
        int cc = 4;
        while (cc * 2 <= c && (size_t)cc * 2 * n * 4 <= GS_ROW_BYTES) cc *= 2;
        const int chunks = (c + cc - 1) / cc;
        long long parts = (4LL * 148 * 3 + (long long)chunks * b - 1) / ((long long)chunks * b);
        const long long max_parts = (nk + GT * 4 - 1) / (GT * 4);
        if (parts > max_parts) parts = max_parts;
        if (parts < 1) parts = 1;
        long long jpart = (nk + parts - 1) / parts;
        jpart = (jpart + GT * 4 - 1) / (GT * 4) * (GT * 4);
        parts = (nk + jpart - 1) / jpart;
        const size_t smem = (size_t)cc * n * 4;
        dim3 sgrid((unsigned)parts, chunks, b);
        static const char* ef_env = tune_env("PDGN_EDGE_FWD_IMPL");   // "plain": the unstaged kernel (A/B)
        if (sgrid.y <= 65535 && !(ef_env && ef_env[0] == 'p')) {
            PDGN_CUDA(cudaFuncSetAttribute(edge_fwd_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            edge_fwd_smem_kernel<<<sgrid, GT, smem, (cudaStream_t)stream>>>(x, ip, c, n, k, cc, (int)jpart, ee);
            PDGN_CHECK_LAUNCH();
            return PDGN_OK;
        }
    }
    const long long per = vec ? 4 : 1;
    const unsigned gx = (unsigned)((nk + GT * per - 1) / (GT * per));
    const int cpb = pick_cpb((long long)gx * b, c);
    dim3 grid(gx, (c + cpb - 1) / cpb, b);
    if (grid.y > 65535) return PDGN_ERR_UNSUPPORTED;
    if (vec) edge_fwd_kernel<true><<<grid, GT, 0, (cudaStream_t)stream>>>(x, ip, c, n, k, cpb, ee);
    else edge_fwd_kernel<false><<<grid, GT, 0, (cudaStream_t)stream>>>(x, ip, c, n, k, cpb, ee);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}

extern "C" int pdgn_edge_feat_bwd(const float* grad_ee, const int64_t* idx, int b, int c, int n, int k, float* grad_x,
                                  void* stream) {
    PDGN_RANGE("pdgn_edge_feat_bwd");
    PDGN_GATHER_ARGS_OK(b >= 0 && c >= 0 && n >= 0 && k >= 0);
    if (b == 0 || c == 0 || n == 0 || k == 0) return PDGN_OK;
    PDGN_GATHER_ARGS_OK(grad_ee && idx && grad_x);
    PDGN_VERIFY_IDX64(idx, (size_t)b * n * k, n, (cudaStream_t)stream);
    if (b > 65535) return PDGN_ERR_UNSUPPORTED;
    const unsigned gx = (unsigned)((n + GT - 1) / GT);
    const int cpb = pick_cpb((long long)gx * b, c);
    dim3 grid(gx, (c + cpb - 1) / cpb, b);
    if (grid.y > 65535) return PDGN_ERR_UNSUPPORTED;
    edge_bwd_kernel<<<grid, GT, 0, (cudaStream_t)stream>>>(grad_ee, reinterpret_cast<const long long*>(idx), c, n, k, cpb, grad_x);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}

extern "C" size_t pdgn_edge_feat_bwd_workspace(int b, int n, int k) {
    if (b < 0 || n < 0 || k < 0) return 0;
    return ((size_t)b * ((size_t)n + 1) + (size_t)b * n * k) * sizeof(int) + 256;
}

extern "C" int pdgn_edge_feat_bwd_ws(const float* grad_ee, const int64_t* idx, int b, int c, int n, int k, float* grad_x,
                                     void* workspace, size_t workspace_bytes, void* stream) {
    PDGN_RANGE("pdgn_edge_feat_bwd_ws");
    PDGN_GATHER_ARGS_OK(b >= 0 && c >= 0 && n >= 0 && k >= 0);
    if (b == 0 || c == 0 || n == 0 || k == 0) return PDGN_OK;
    PDGN_GATHER_ARGS_OK(grad_ee && idx && grad_x);
    PDGN_VERIFY_IDX64(idx, (size_t)b * n * k, n, (cudaStream_t)stream);
    if (b > 65535) return PDGN_ERR_UNSUPPORTED;
    const long long nk = (long long)n * k;
    const size_t csr_smem = csr_smem_total(n, nk);
    const size_t row_bytes = (size_t)nk * 4;  // the g1 row of one channel
    if (!workspace || row_bytes > 200 * 1024 || !csr_ok(n, nk))
        return pdgn_edge_feat_bwd(grad_ee, idx, b, c, n, k, grad_x, stream);
    if (workspace_bytes < pdgn_edge_feat_bwd_workspace(b, n, k) - 256 || (reinterpret_cast<uintptr_t>(workspace) & 3)) return PDGN_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    int* offs = reinterpret_cast<int*>(workspace);
    int* pos = offs + (size_t)b * (n + 1);
    PDGN_CUDA(cudaFuncSetAttribute(csr_build_kernel<long long>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csr_smem));
    csr_build_kernel<long long><<<b, CSR_T, csr_smem, st>>>(reinterpret_cast<const long long*>(idx), n, (int)nk, offs, pos, csr_stage(n, nk) ? 1 : 0);
    PDGN_CHECK_LAUNCH();
    bool launched = false;
    const int rc_stream = pull_stream_launch(1, grad_ee, offs, pos, b, c, n, (int)nk, k, grad_x, st, &launched, nullptr);
    if (rc_stream != PDGN_OK || launched) return rc_stream;
    int cc = 1;
    while (cc < 4 && (size_t)cc * 2 * row_bytes <= 100 * 1024) cc *= 2;
    const size_t smem = (size_t)cc * row_bytes;
    PDGN_CUDA(cudaFuncSetAttribute(edge_bwd_pull_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((c + cc - 1) / cc, b);
    edge_bwd_pull_kernel<<<grid, 512, smem, st>>>(grad_ee, offs, pos, c, n, k, cc, grad_x);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}
