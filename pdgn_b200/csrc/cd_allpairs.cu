// cd_allpairs.cu -- all-pairs Chamfer-distance matrix (the evaluation hot path).
//
// Replaces the Python double loop _pairwise_EMD_CD_ -> distChamfer of the reference
// (evaluation/evaluation_metrics.py:85-121 and :35-45; three bmm + two min over a [bs,2048,2048] matrix per
// iteration, 20 000 iterations per 1000x1000 matrix) with one launch per matrix tile.
//
// Design (DESIGN.md "cd_allpairs"):
//   * FP32 SIMT, direct-difference distance with the reference's native rounding (d2_xyz, common.cuh): the
//     kernel is bound by the FP32 pipe (6 FMA-pipe instructions per point pair), not HBM and not tensor cores.
//   * One CTA = two "halves" of 128 threads.  A half owns one cloud of A: every thread keeps 16 of its points
//     (rows) in registers for the whole strip, so 128 threads hold a 2048-point cloud.  Both halves scan the
//     same B cloud, streamed tile by tile into shared memory by the TMA engine (cp.async.bulk + mbarrier,
//     double buffered), read back as warp-broadcast LDS.128 of SoA planes (4 candidates per load).
//   * Each distance is computed once and feeds both directions: the row minimum stays in the thread's
//     registers (FMNMX3 over two candidates), the column minimum is folded over the thread's 16 rows
//     (FMNMX3), reduced across the warp with one CREDUX.MIN on the bit pattern (distances are >= 0 so uint
//     order == float order) and merged across the half's 4 warps with one shared-memory ATOMS.MIN.
//   * Per cloud pair only one scalar leaves the SM: (sum_i rowmin + sum_j colmin) / npts.
#include "common.cuh"

namespace pdgn {

constexpr int CD_R = 16;                    // rows (points of the A cloud) per thread
constexpr int CD_HALF = 128;                // threads per half
constexpr int CD_THREADS = 2 * CD_HALF;
constexpr int CD_ROWS = CD_R * CD_HALF;     // 2048 rows per half per row block
constexpr int CD_TILE = 2048;               // candidates per shared-memory stage
constexpr unsigned CD_INF_BITS = 0x7f800000u;

// AoS [cloud][npts][3] -> SoA planes [cloud][3][npad]; pad entries replicate point 0 (harmless for minima).
__global__ void cd_pack_kernel(const float* __restrict__ src, int cloud0, int npts, int npad, float* __restrict__ dst) {
    const int cl = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= npad) return;
    const float* p = src + ((size_t)(cloud0 + cl) * npts + (j < npts ? j : 0)) * 3;
    float* d = dst + (size_t)cl * 3 * npad + j;
    d[0] = p[0];
    d[npad] = p[1];
    d[2 * (size_t)npad] = p[2];
}

// Two candidates against the thread's 16 rows.
__device__ __forceinline__ void cd_two_candidates(const float (&qx)[CD_R], const float (&qy)[CD_R], const float (&qz)[CD_R],
                                                  float (&rowmin)[CD_R], float x0, float y0, float z0, float x1, float y1,
                                                  float z1, unsigned* col, int lane) {
    float c0, c1;
    {
        const float a0 = d2_xyz(qx[0], qy[0], qz[0], x0, y0, z0), a1 = d2_xyz(qx[0], qy[0], qz[0], x1, y1, z1);
        const float b0 = d2_xyz(qx[1], qy[1], qz[1], x0, y0, z0), b1 = d2_xyz(qx[1], qy[1], qz[1], x1, y1, z1);
        rowmin[0] = min3(rowmin[0], a0, a1);
        rowmin[1] = min3(rowmin[1], b0, b1);
        c0 = fminf(a0, b0);
        c1 = fminf(a1, b1);
    }
#pragma unroll
    for (int k = 2; k < CD_R; k += 2) {
        const float a0 = d2_xyz(qx[k], qy[k], qz[k], x0, y0, z0), a1 = d2_xyz(qx[k], qy[k], qz[k], x1, y1, z1);
        const float b0 = d2_xyz(qx[k + 1], qy[k + 1], qz[k + 1], x0, y0, z0);
        const float b1 = d2_xyz(qx[k + 1], qy[k + 1], qz[k + 1], x1, y1, z1);
        rowmin[k] = min3(rowmin[k], a0, a1);
        rowmin[k + 1] = min3(rowmin[k + 1], b0, b1);
        c0 = min3(c0, a0, b0);
        c1 = min3(c1, a1, b1);
    }
    const unsigned r0 = __reduce_min_sync(kFull, __float_as_uint(c0));
    const unsigned r1 = __reduce_min_sync(kFull, __float_as_uint(c1));
    if (lane < 2) atomicMin(col + lane, lane ? r1 : r0);
}

__global__ void __launch_bounds__(CD_THREADS, 2)
cd_allpairs_kernel(const float* __restrict__ PA, const float* __restrict__ PB, int nrows, int ncols, int npts, int npad,
                   int rstrip, float* __restrict__ out, long long ld_out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tile = reinterpret_cast<float*>(smem_raw);                 // [2 stages][3 planes][CD_TILE]
    unsigned* colmin = reinterpret_cast<unsigned*>(tile + 2 * 3 * CD_TILE);  // [2 halves][npad]
    float* red = reinterpret_cast<float*>(colmin + 2 * (size_t)npad);        // [2 halves][4 warps]
    uint64_t* bars = reinterpret_cast<uint64_t*>(red + 8);                  // [2 stages]

    const int tid = threadIdx.x, half = tid >> 7, ht = tid & (CD_HALF - 1), lane = tid & 31, hw = ht >> 5;
    int s = blockIdx.y * 2 + half;
    const bool s_valid = s < nrows;
    if (!s_valid) s = nrows - 1;  // odd row count: the spare half recomputes the last cloud and discards it
    const int r_begin = blockIdx.x * rstrip;
    const int r_end = min(ncols, r_begin + rstrip);
    const int nrb = (npts + CD_ROWS - 1) / CD_ROWS;
    const int ncb = (npad + CD_TILE - 1) / CD_TILE;
    const int ntiles = (r_end - r_begin) * nrb * ncb;
    unsigned* mycol = colmin + (size_t)half * npad;

    for (int j = ht; j < npad; j += CD_HALF) mycol[j] = CD_INF_BITS;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // tile t (flattened over r, row block, candidate block) -> stage t&1
    auto issue = [&](int t) {
        const int cb = t % ncb;
        const int r = r_begin + t / (ncb * nrb);
        const int c0 = cb * CD_TILE;
        const unsigned bytes = (unsigned)min(CD_TILE, npad - c0) * 4u;
        uint64_t* bar = &bars[t & 1];
        float* dst = tile + (t & 1) * 3 * CD_TILE;
        const float* src = PB + (size_t)r * 3 * npad + c0;
        mbar_expect_tx(bar, 3u * bytes);
        bulk_g2s(dst, src, bytes, bar);
        bulk_g2s(dst + CD_TILE, src + npad, bytes, bar);
        bulk_g2s(dst + 2 * CD_TILE, src + 2 * (size_t)npad, bytes, bar);
    };
    if (tid == 0) {
        issue(0);
        if (ntiles > 1) issue(1);
    }

    float qx[CD_R], qy[CD_R], qz[CD_R], rowmin[CD_R];
    const float* arow = PA + (size_t)s * 3 * npad;
    const float inv_n = 1.0f / (float)npts;
    int t = 0;
    for (int r = r_begin; r < r_end; ++r) {
        float total = 0.f;
        for (int rb = 0; rb < nrb; ++rb) {
            const int i0 = rb * CD_ROWS + ht * CD_R;
            if (nrb > 1 || r == r_begin) {
#pragma unroll
                for (int k = 0; k < CD_R; ++k) {
                    const int i = (i0 + k < npts) ? i0 + k : 0;
                    qx[k] = arow[i];
                    qy[k] = arow[npad + i];
                    qz[k] = arow[2 * (size_t)npad + i];
                }
            }
#pragma unroll
            for (int k = 0; k < CD_R; ++k) rowmin[k] = __int_as_float(CD_INF_BITS);

            for (int cb = 0; cb < ncb; ++cb, ++t) {
                const float* st = tile + (t & 1) * 3 * CD_TILE;
                const int cnt = min(CD_TILE, npad - cb * CD_TILE);
                unsigned* col = mycol + cb * CD_TILE;
                mbar_wait(&bars[t & 1], (unsigned)((t >> 1) & 1));
#pragma unroll 1
                for (int j = 0; j < cnt; j += 4) {
                    const float4 X = *reinterpret_cast<const float4*>(st + j);
                    const float4 Y = *reinterpret_cast<const float4*>(st + CD_TILE + j);
                    const float4 Z = *reinterpret_cast<const float4*>(st + 2 * CD_TILE + j);
                    cd_two_candidates(qx, qy, qz, rowmin, X.x, Y.x, Z.x, X.y, Y.y, Z.y, col + j, lane);
                    cd_two_candidates(qx, qy, qz, rowmin, X.z, Y.z, Z.z, X.w, Y.w, Z.w, col + j + 2, lane);
                }
                __syncthreads();  // stage drained by all 8 warps; this tile's column atomics are done
                if (tid == 0 && t + 2 < ntiles) {
                    fence_proxy_async();
                    issue(t + 2);
                }
            }
            const int nvalid = npts - i0;
#pragma unroll
            for (int k = 0; k < CD_R; ++k)
                if (k < nvalid) total += rowmin[k];
        }
        // cloud pair (s, r) complete: fold this half's column minima, reset them for the next r
        for (int j = ht; j < npad; j += CD_HALF) {
            if (j < npts) total += __uint_as_float(mycol[j]);
            mycol[j] = CD_INF_BITS;
        }
        total = warp_sum(total);
        if (lane == 0) red[half * 4 + hw] = total;
        __syncthreads();
        if (ht == 0 && s_valid) {
            const float* rr = red + half * 4;
            out[(size_t)s * ld_out + r] = (rr[0] + rr[1] + rr[2] + rr[3]) * inv_n;
        }
    }
}

static int cd_num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;
    }
    return sms;
}

static inline int cd_npad(int npts) { return (npts + 15) & ~15; }

}  // namespace pdgn

using namespace pdgn;

extern "C" size_t pdgn_cd_allpairs_workspace(int na, int nb, int npts) {
    if (na < 0 || nb < 0 || npts <= 0) return 0;
    return ((size_t)na + (size_t)nb) * 3 * (size_t)cd_npad(npts) * sizeof(float) + 256;
}

extern "C" int pdgn_cd_allpairs(const float* A, const float* B, int na, int nb, int npts, int row0, int row1, int col0,
                                int col1, float* out, long long ld_out, void* workspace, size_t workspace_bytes,
                                void* stream) {
    if (!A || !B || !out || na < 0 || nb < 0 || npts <= 0) return PDGN_ERR_BAD_ARG;
    if (row0 < 0 || row1 > na || row0 > row1 || col0 < 0 || col1 > nb || col0 > col1) return PDGN_ERR_BAD_ARG;
    if (npts > 16384) return PDGN_ERR_UNSUPPORTED;
    const int nrows = row1 - row0, ncols = col1 - col0;
    if (nrows == 0 || ncols == 0) return PDGN_OK;
    if (ld_out < ncols) return PDGN_ERR_BAD_ARG;
    const int npad = cd_npad(npts);
    const size_t need = ((size_t)nrows + ncols) * 3 * npad * sizeof(float);
    if (!workspace || workspace_bytes < need) return PDGN_ERR_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 15) != 0) return PDGN_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    float* PA = reinterpret_cast<float*>(workspace);
    float* PB = PA + (size_t)nrows * 3 * npad;

    dim3 pb(256), pga((npad + 255) / 256, nrows), pgb((npad + 255) / 256, ncols);
    cd_pack_kernel<<<pga, pb, 0, st>>>(A, row0, npts, npad, PA);
    PDGN_CHECK_LAUNCH();
    cd_pack_kernel<<<pgb, pb, 0, st>>>(B, col0, npts, npad, PB);
    PDGN_CHECK_LAUNCH();

    const size_t smem = (size_t)(2 * 3 * CD_TILE + 2 * (size_t)npad + 8) * 4 + 2 * sizeof(uint64_t);
    PDGN_CUDA(cudaFuncSetAttribute(cd_allpairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int spairs = (nrows + 1) / 2;
    if (spairs > 65535) return PDGN_ERR_UNSUPPORTED;
    // enough CTAs for >= ~20 waves of 2 CTAs/SM so the tail is small; each CTA walks `rstrip` B clouds
    const int target = 20 * 2 * cd_num_sms();
    int strips = (target + spairs - 1) / spairs;
    if (strips > ncols) strips = ncols;
    if (strips < 1) strips = 1;
    const int rstrip = (ncols + strips - 1) / strips;
    strips = (ncols + rstrip - 1) / rstrip;
    cd_allpairs_kernel<<<dim3(strips, spairs), CD_THREADS, smem, st>>>(PA, PB, nrows, ncols, npts, npad, rstrip, out, ld_out);
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}

extern "C" int pdgn_cd_allpairs_host(const float* A_host, const float* B_host, int na, int nb, int npts, int row0, int row1,
                                     int col0, int col1, float* out_host, long long ld_out, void* stream) {
    if (!A_host || !B_host || !out_host || na < 0 || nb < 0 || npts <= 0) return PDGN_ERR_BAD_ARG;
    if (row0 < 0 || row1 > na || row0 > row1 || col0 < 0 || col1 > nb || col0 > col1) return PDGN_ERR_BAD_ARG;
    const int nrows = row1 - row0, ncols = col1 - col0;
    if (nrows == 0 || ncols == 0) return PDGN_OK;
    if (ld_out < ncols) return PDGN_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t abytes = (size_t)nrows * npts * 3 * sizeof(float), bbytes = (size_t)ncols * npts * 3 * sizeof(float);
    const size_t obytes = (size_t)nrows * ncols * sizeof(float);
    const size_t wbytes = pdgn_cd_allpairs_workspace(nrows, ncols, npts);
    float *dA = nullptr, *dB = nullptr, *dO = nullptr;
    void* ws = nullptr;
    int rc = PDGN_OK;
    cudaError_t e;
    if ((e = cudaMallocAsync(&dA, abytes, st)) != cudaSuccess) return (int)e;
    if ((e = cudaMallocAsync(&dB, bbytes, st)) != cudaSuccess) { cudaFreeAsync(dA, st); return (int)e; }
    if ((e = cudaMallocAsync(&dO, obytes, st)) != cudaSuccess) { cudaFreeAsync(dA, st); cudaFreeAsync(dB, st); return (int)e; }
    if ((e = cudaMallocAsync(&ws, wbytes, st)) != cudaSuccess) { cudaFreeAsync(dA, st); cudaFreeAsync(dB, st); cudaFreeAsync(dO, st); return (int)e; }
    e = cudaMemcpyAsync(dA, A_host + (size_t)row0 * npts * 3, abytes, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dB, B_host + (size_t)col0 * npts * 3, bbytes, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        rc = pdgn_cd_allpairs(dA, dB, nrows, ncols, npts, 0, nrows, 0, ncols, dO, ncols, ws, wbytes, stream);
        if (rc == PDGN_OK)
            e = cudaMemcpy2DAsync(out_host, (size_t)ld_out * sizeof(float), dO, (size_t)ncols * sizeof(float),
                                  (size_t)ncols * sizeof(float), nrows, cudaMemcpyDeviceToHost, st);
    }
    cudaFreeAsync(dA, st);
    cudaFreeAsync(dB, st);
    cudaFreeAsync(dO, st);
    cudaFreeAsync(ws, st);
    cudaError_t es = cudaStreamSynchronize(st);
    if (rc != PDGN_OK) return rc;
    if (e != cudaSuccess) return (int)e;
    return (int)es;
}
