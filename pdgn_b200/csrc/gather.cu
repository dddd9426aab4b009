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
#include "common.cuh"

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
    PDGN_GATHER_ARGS_OK(points && idx && out && b >= 0 && c >= 0 && n > 0 && m >= 0 && k >= 0);
    const long long mk = (long long)m * k;
    if (b == 0 || c == 0 || mk == 0) return PDGN_OK;
    if (mk > 0x7fffffffLL || b > 65535) return PDGN_ERR_UNSUPPORTED;
    const bool vec = (mk % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(idx)) % 16 == 0);
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

extern "C" int pdgn_group_bwd(const float* grad_out, const int* idx, int b, int c, int n, int m, int k, float* grad_points,
                              void* stream) {
    PDGN_GATHER_ARGS_OK(grad_out && idx && grad_points && b >= 0 && c >= 0 && n > 0 && m >= 0 && k >= 0);
    const long long mk = (long long)m * k;
    if (b == 0 || c == 0 || mk == 0) return PDGN_OK;
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
    PDGN_GATHER_ARGS_OK(points && idx && weight && out && b >= 0 && c >= 0 && m > 0 && n >= 0);
    if (b == 0 || c == 0 || n == 0) return PDGN_OK;
    if (b > 65535) return PDGN_ERR_UNSUPPORTED;
    const bool vec = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(idx) |
                                       reinterpret_cast<uintptr_t>(weight)) % 16 == 0);
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

extern "C" int pdgn_interp_bwd(const float* grad_out, const int* idx, const float* weight, int b, int c, int n, int m,
                               float* grad_points, void* stream) {
    PDGN_GATHER_ARGS_OK(grad_out && idx && weight && grad_points && b >= 0 && c >= 0 && m > 0 && n >= 0);
    if (b == 0 || c == 0 || n == 0) return PDGN_OK;
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
    PDGN_GATHER_ARGS_OK(x && idx && ee && b >= 0 && c >= 0 && n >= 0 && k >= 0);
    const long long nk = (long long)n * k;
    if (b == 0 || c == 0 || nk == 0) return PDGN_OK;
    if (b > 65535 || nk > 0x7fffffffLL) return PDGN_ERR_UNSUPPORTED;
    const bool vec = (nk % 4 == 0) && (reinterpret_cast<uintptr_t>(ee) % 16 == 0);
    const long long per = vec ? 4 : 1;
    const unsigned gx = (unsigned)((nk + GT * per - 1) / (GT * per));
    const int cpb = pick_cpb((long long)gx * b, c);
    dim3 grid(gx, (c + cpb - 1) / cpb, b);
    if (grid.y > 65535) return PDGN_ERR_UNSUPPORTED;
    const long long* ip = reinterpret_cast<const long long*>(idx);
    if (vec) edge_fwd_kernel<true><<<grid, GT, 0, (cudaStream_t)stream>>>(x, ip, c, n, k, cpb, ee);
    else edge_fwd_kernel<false><<<grid, GT, 0, (cudaStream_t)stream>>>(x, ip, c, n, k, cpb, ee);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}

extern "C" int pdgn_edge_feat_bwd(const float* grad_ee, const int64_t* idx, int b, int c, int n, int k, float* grad_x,
                                  void* stream) {
    PDGN_GATHER_ARGS_OK(grad_ee && idx && grad_x && b >= 0 && c >= 0 && n >= 0 && k >= 0);
    if (b == 0 || c == 0 || n == 0 || k == 0) return PDGN_OK;
    if (b > 65535) return PDGN_ERR_UNSUPPORTED;
    const unsigned gx = (unsigned)((n + GT - 1) / GT);
    const int cpb = pick_cpb((long long)gx * b, c);
    dim3 grid(gx, (c + cpb - 1) / cpb, b);
    if (grid.y > 65535) return PDGN_ERR_UNSUPPORTED;
    edge_bwd_kernel<<<grid, GT, 0, (cudaStream_t)stream>>>(grad_ee, reinterpret_cast<const long long*>(idx), c, n, k, cpb, grad_x);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}
