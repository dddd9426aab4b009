// chamfer.cu -- paired (per batch element) directional nearest-neighbour distance with argmin, and its gradient.
//
// Replaces NmDistanceKernel / NmDistanceGradKernel (evaluation/pytorch_structural_losses/src/nndistance.cu:2-154)
// for d == 3 and the bmm + min composition of ChamferLoss (utils/chamfer_loss.py:13-38) and distChamfer
// (evaluation/evaluation_metrics.py:35-45) for any small d.  Distances are direct differences in FP32
// (d == 3: the reference's native rounding, d2_xyz; otherwise an fma chain over the channels), so the result is
// closer to the FP64 truth than the reference's Gram form and bit-identical to nndistance for d == 3.
// Strict '<' while scanning candidates in index order => the lowest index among equal minima, as
// nndistance.cu:29-32,116-119 orders them.
//
// Layout: a CTA owns CH_Q*128 query points of one batch element (CH_Q per thread, in registers) and streams the
// candidate set through shared memory as SoA planes (warp-broadcast LDS.128 = 4 candidates per plane per load).
#include "common.cuh"
#include "multi.cuh"

namespace pdgn {

constexpr int CH_T = 128;      // threads per CTA
constexpr int CH_Q = 2;        // queries per thread
// candidates per tile: SoA planes must fit the 48 KB static shared-memory window
template <int D> struct ChTile { static constexpr int value = D <= 8 ? 1024 : 512; };
constexpr int CH_DMAX = 16;

template <int D>
__device__ __forceinline__ float sqdist(const float (&q)[D], const float (&p)[D]) {
    if (D == 3) return d2_xyz(q[0], q[1], q[2], p[0], p[1], p[2]);
    float diff = __fsub_rn(q[0], p[0]);
    float t = __fmul_rn(diff, diff);
#pragma unroll
    for (int c = 1; c < D; ++c) {
        diff = __fsub_rn(q[c], p[c]);
        t = __fmaf_rn(diff, diff, t);
    }
    return t;
}

// x [b,nx,D] queries, y [b,ny,D] candidates -> mind [b,nx], argm [b,nx] (argm may be null)
// S "slices" of CH_T threads share the queries of a CTA: slice s scans the candidate quads s, s+S, s+2S, ... of every tile
// and the S partial (minimum, index) pairs of a query are merged lexicographically at the end (lowest index among equal
// minima, exactly what a single in-order scan gives).  The training shapes have only a few hundred queries per batch
// element: with S = 1 they put one warp on each scheduler (27 us for 35 x 1024 x 1024, 21 % of the issue roofline).
template <int D, int S>
__device__ __forceinline__ void nn_min_body(const float* __restrict__ x, const float* __restrict__ y, int nx, int ny,
                                            float* __restrict__ mind, int* __restrict__ argm, int bx, int bz) {
    constexpr int CH_TILE = ChTile<D>::value;
    constexpr int MERGE = S > 1 ? 2 * S * CH_T * CH_Q : 0;          // floats needed to merge the slices
    constexpr int SMEM = D * CH_TILE > MERGE ? D * CH_TILE : MERGE;
    __shared__ __align__(16) float tile[SMEM];
    const int qt = threadIdx.x % CH_T, slice = threadIdx.x / CH_T;  // slice is warp-uniform (CH_T is a multiple of 32)
    const int q0 = (bx * CH_T + qt) * CH_Q;
    float q[CH_Q][D];
    float best[CH_Q];
    int besti[CH_Q];
#pragma unroll
    for (int r = 0; r < CH_Q; ++r) {
        const int qi = min(q0 + r, nx - 1);
#pragma unroll
        for (int c = 0; c < D; ++c) q[r][c] = x[((size_t)bz * nx + qi) * D + c];
        best[r] = __int_as_float(0x7f800000);
        besti[r] = 0;
    }
    const float* yb = y + (size_t)bz * ny * D;
    for (int j0 = 0; j0 < ny; j0 += CH_TILE) {
        const int cnt = min(CH_TILE, ny - j0);
        const int cnt4 = (cnt + 3) & ~3;
        __syncthreads();
        // AoS global -> SoA shared (coalesced global reads); pad the tail quad with the tile's first candidate
        for (int e = threadIdx.x; e < cnt4 * D; e += CH_T * S) {
            const int j = e / D, c = e - j * D;
            tile[c * CH_TILE + j] = yb[(size_t)(j0 + (j < cnt ? j : 0)) * D + c];
        }
        __syncthreads();
#pragma unroll 1
        for (int j = 4 * slice; j < cnt4; j += 4 * S) {
            float4 P[D];
#pragma unroll
            for (int c = 0; c < D; ++c) P[c] = *reinterpret_cast<const float4*>(tile + c * CH_TILE + j);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float p[D];
#pragma unroll
                for (int c = 0; c < D; ++c) p[c] = u == 0 ? P[c].x : u == 1 ? P[c].y : u == 2 ? P[c].z : P[c].w;
#pragma unroll
                for (int r = 0; r < CH_Q; ++r) {
                    const float d = sqdist<D>(q[r], p);
                    // a padded duplicate is never strictly smaller than its original within a slice, and across slices the
                    // original wins the merge with its lower index
                    if (d < best[r]) {
                        best[r] = d;
                        besti[r] = j0 + j + u;
                    }
                }
            }
        }
    }
    if (S > 1) {
        __syncthreads();  // the last tile is consumed: its storage becomes the merge buffer
        float* sb = tile;                                        // [S][CH_T*CH_Q]
        int* si = reinterpret_cast<int*>(tile + S * CH_T * CH_Q);  // [S][CH_T*CH_Q]
#pragma unroll
        for (int r = 0; r < CH_Q; ++r) {
            sb[slice * CH_T * CH_Q + r * CH_T + qt] = best[r];
            si[slice * CH_T * CH_Q + r * CH_T + qt] = besti[r];
        }
        __syncthreads();
        if (slice != 0) return;
#pragma unroll
        for (int r = 0; r < CH_Q; ++r)
#pragma unroll
            for (int o = 1; o < S; ++o) {
                const float d = sb[o * CH_T * CH_Q + r * CH_T + qt];
                const int i = si[o * CH_T * CH_Q + r * CH_T + qt];
                if (d < best[r] || (d == best[r] && i < besti[r])) {
                    best[r] = d;
                    besti[r] = i;
                }
            }
    }
#pragma unroll
    for (int r = 0; r < CH_Q; ++r)
        if (q0 + r < nx) {
            mind[(size_t)bz * nx + q0 + r] = best[r];
            if (argm) argm[(size_t)bz * nx + q0 + r] = besti[r];
        }
}

template <int D, int S>
__global__ void __launch_bounds__(CH_T * S) nn_min_kernel(const float* __restrict__ x, const float* __restrict__ y, int nx, int ny,
                                                         float* __restrict__ mind, int* __restrict__ argm) {
    nn_min_body<D, S>(x, y, nx, ny, mind, argm, blockIdx.x, blockIdx.y);
}

// problem-descriptor launches (multi.cuh)
template <int D, int S>
__global__ void __launch_bounds__(CH_T * S) nn_min_multi_kernel(const __grid_constant__ MinTable tb) {
    const MinProb& pr = tb.p[multi_find(tb, blockIdx.x)];
    nn_min_body<D, S>(pr.x, pr.y, pr.nx, pr.ny, pr.mind, pr.argm, blockIdx.x - pr.cta0, blockIdx.y);
}

// backward of sum_i g * inv_m * min_i for every problem of the table: d/dx_i = 2 g inv_m (x_i - y_a(i)), opposite sign on y_a(i)
__global__ void __launch_bounds__(256) chamfer_bwd_multi_kernel(const __grid_constant__ MinTable tb, int d) {
    const MinProb& pr = tb.p[multi_find(tb, blockIdx.x)];
    const int bz = blockIdx.y;
    const int i = (blockIdx.x - pr.cta0) * 256 + threadIdx.x;
    if (i >= pr.nx) return;
    const int j = pr.argm[(size_t)bz * pr.nx + i];
    const float g = 2.f * __ldg(pr.gscale) * pr.inv_m;
    const float* xp = pr.x + ((size_t)bz * pr.nx + i) * d;
    const float* yp = pr.y + ((size_t)bz * pr.ny + j) * d;
    float* gx = pr.gx + ((size_t)bz * pr.nx + i) * d;
    float* gy = pr.gy + ((size_t)bz * pr.ny + j) * d;
    for (int c = 0; c < d; ++c) {
        const float v = g * (xp[c] - yp[c]);
        atomicAdd(gx + c, v);
        atomicAdd(gy + c, -v);
    }
}

// d/dx_i of w_i * |x_i - y_a(i)|^2 = 2 w_i (x_i - y_a(i)), and the opposite sign on y_a(i) (nndistance.cu:129-148).
__global__ void __launch_bounds__(256) chamfer_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, int nx, int ny,
                                                         int d, const float* __restrict__ w, const int* __restrict__ arg,
                                                         float* __restrict__ grad_x, float* __restrict__ grad_y) {
    const int bz = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nx) return;
    const int j = arg[(size_t)bz * nx + i];
    const float g = 2.f * w[(size_t)bz * nx + i];
    const float* xp = x + ((size_t)bz * nx + i) * d;
    const float* yp = y + ((size_t)bz * ny + j) * d;
    float* gx = grad_x + ((size_t)bz * nx + i) * d;
    float* gy = grad_y + ((size_t)bz * ny + j) * d;
    for (int c = 0; c < d; ++c) {
        const float v = g * (xp[c] - yp[c]);
        atomicAdd(gx + c, v);
        atomicAdd(gy + c, -v);
    }
}

template <int D>
static int launch_nn_min(const float* x, const float* y, int b, int nx, int ny, float* mind, int* argm, cudaStream_t st) {
    dim3 grid((nx + CH_T * CH_Q - 1) / (CH_T * CH_Q), b);
    // slices per CTA: enough warps to give every scheduler of the chip ~4, but at least 16 candidate quads per slice
    static const int sms = [] {
        int dev = 0, v = 148;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        return v;
    }();
    const long long ctas = (long long)grid.x * grid.y;
    constexpr int kMaxSlices = D <= 8 ? 8 : 4;  // 1024 threads leave 64 registers: not enough for D > 8 channels
    int slices = 1;
    while (slices < kMaxSlices && ctas * (CH_T / 32) * slices < 16LL * sms && ny >= 64 * (2 * slices)) slices <<= 1;
    switch (slices) {
        case 8:
            if constexpr (kMaxSlices >= 8) nn_min_kernel<D, 8><<<grid, CH_T * 8, 0, st>>>(x, y, nx, ny, mind, argm);
            break;
        case 4: nn_min_kernel<D, 4><<<grid, CH_T * 4, 0, st>>>(x, y, nx, ny, mind, argm); break;
        case 2: nn_min_kernel<D, 2><<<grid, CH_T * 2, 0, st>>>(x, y, nx, ny, mind, argm); break;
        default: nn_min_kernel<D, 1><<<grid, CH_T, 0, st>>>(x, y, nx, ny, mind, argm); break;
    }
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}

static int dispatch_nn_min(const float* x, const float* y, int b, int nx, int ny, int d, float* mind, int* argm, cudaStream_t st) {
    switch (d) {
#define PDGN_CASE(D_) case D_: return launch_nn_min<D_>(x, y, b, nx, ny, mind, argm, st);
        PDGN_CASE(1) PDGN_CASE(2) PDGN_CASE(3) PDGN_CASE(4) PDGN_CASE(5) PDGN_CASE(6) PDGN_CASE(7) PDGN_CASE(8)
        PDGN_CASE(9) PDGN_CASE(10) PDGN_CASE(11) PDGN_CASE(12) PDGN_CASE(13) PDGN_CASE(14) PDGN_CASE(15) PDGN_CASE(16)
#undef PDGN_CASE
    }
    return PDGN_ERR_UNSUPPORTED;
}

template <int D>
static int nn_min_multi_launch_d(MinTable& tb, int b, cudaStream_t st) {
    constexpr int S = 4;                    // slices per CTA: the training shapes have few queries per batch element
    int ctas = 0;
    for (int i = 0; i < tb.count; ++i) {
        tb.p[i].cta0 = ctas;
        ctas += (tb.p[i].nx + CH_T * CH_Q - 1) / (CH_T * CH_Q);
    }
    nn_min_multi_kernel<D, S><<<dim3(ctas, b), CH_T * S, 0, st>>>(tb);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}
int nn_min_multi_launch(MinTable& tb, int b, int d, cudaStream_t st) {
    if (tb.count < 1 || tb.count > MULTI_MAX) return PDGN_ERR_UNSUPPORTED;
    if (d == 3) return nn_min_multi_launch_d<3>(tb, b, st);
    if (d == 6) return nn_min_multi_launch_d<6>(tb, b, st);
    if (d == 9) return nn_min_multi_launch_d<9>(tb, b, st);
    return PDGN_ERR_UNSUPPORTED;
}
int chamfer_bwd_multi_launch(MinTable& tb, int b, int d, cudaStream_t st) {
    if (tb.count < 1 || tb.count > MULTI_MAX) return PDGN_ERR_UNSUPPORTED;
    int ctas = 0;
    for (int i = 0; i < tb.count; ++i) {
        tb.p[i].cta0 = ctas;
        ctas += (tb.p[i].nx + 255) / 256;
    }
    chamfer_bwd_multi_kernel<<<dim3(ctas, b), 256, 0, st>>>(tb, d);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}

}  // namespace pdgn

using namespace pdgn;

extern "C" int pdgn_chamfer_min(const float* x, const float* y, int b, int nx, int ny, int d, float* min_xy, int* arg_xy,
                                float* min_yx, int* arg_yx, void* stream) {
    PDGN_RANGE("pdgn_chamfer_min");
    if (b < 0 || nx < 0 || ny < 0) return PDGN_ERR_BAD_ARG;
    if (d < 1 || d > CH_DMAX || b > 65535) return PDGN_ERR_UNSUPPORTED;
    if ((!min_xy && arg_xy) || (!min_yx && arg_yx)) return PDGN_ERR_BAD_ARG;
    if (b == 0) return PDGN_OK;
    if (!x || !y) return (nx == 0 && ny == 0) ? PDGN_OK : PDGN_ERR_BAD_ARG;
    if ((nx == 0) != (ny == 0)) return PDGN_ERR_BAD_ARG;  // a minimum over an empty set is undefined
    if (nx == 0) return PDGN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = PDGN_OK;
    if (min_xy) rc = dispatch_nn_min(x, y, b, nx, ny, d, min_xy, arg_xy, st);
    if (rc == PDGN_OK && min_yx) rc = dispatch_nn_min(y, x, b, ny, nx, d, min_yx, arg_yx, st);
    return rc;
}

extern "C" int pdgn_chamfer_bwd(const float* x, const float* y, int b, int nx, int ny, int d, const float* w_xy, const int* arg_xy,
                                const float* w_yx, const int* arg_yx, float* grad_x, float* grad_y, void* stream) {
    PDGN_RANGE("pdgn_chamfer_bwd");
    if (!x || !y || !grad_x || !grad_y || b < 0 || nx < 0 || ny < 0 || d < 1) return PDGN_ERR_BAD_ARG;
    if ((w_xy == nullptr) != (arg_xy == nullptr) || (w_yx == nullptr) != (arg_yx == nullptr)) return PDGN_ERR_BAD_ARG;
    if (b > 65535) return PDGN_ERR_UNSUPPORTED;
    if (b == 0 || nx == 0 || ny == 0) return PDGN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (arg_xy) PDGN_VERIFY_IDX32(arg_xy, (size_t)b * nx, ny, st);
    if (arg_yx) PDGN_VERIFY_IDX32(arg_yx, (size_t)b * ny, nx, st);
    if (w_xy) {
        chamfer_bwd_kernel<<<dim3((nx + 255) / 256, b), 256, 0, st>>>(x, y, nx, ny, d, w_xy, arg_xy, grad_x, grad_y);
        PDGN_CHECK_LAUNCH();
    }
    if (w_yx) {
        chamfer_bwd_kernel<<<dim3((ny + 255) / 256, b), 256, 0, st>>>(y, x, ny, nx, d, w_yx, arg_yx, grad_y, grad_x);
        PDGN_CHECK_LAUNCH();
    }
    return PDGN_OK;
}
