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
//     order == float order) and stored to the warp's own column array (one STS.128 per 4 candidates); the half's
//     4 warps are merged when the cloud pair is folded.  Clouds of more than 2048 points walk several row blocks and
//     merge through shared-memory ATOMS.MIN instead.
//   * Per cloud pair only one scalar leaves the SM: (sum_i rowmin + sum_j colmin) / npts.
#include <cstdlib>
#include "cd_kernel.cuh"

namespace pdgn {

// shipped configuration (chosen with tools/cd_tune.cu on B200; see DESIGN.md / profiles/)
constexpr int CD_R = 16;      // rows per thread
constexpr int CD_NH = 2;      // halves (A clouds) per CTA
constexpr int CD_MINB = 2;    // CTAs per SM the register allocation must allow
constexpr int CD_VARIANT_BIG = CDV_PRED_RED | CDV_PREFETCH;   // npts > 2048: several row blocks merge their column minima by ATOMS.MIN
// npts <= 2048 (every PDGN shape): per-warp column arrays written with plain STS.128 (no atomics, no branches in the inner
// loop); row minima as VIMNMX3.U32 on the bit patterns (d2 >= +0, so the order is the same; ptxas schedules this mix best:
// 0.707 vs 0.695 of the issue roofline, profiles/r01_cd_tune_c.txt)
constexpr int CD_VARIANT = CDV_PREFETCH | CDV_WARPCOL | CDV_IMIN_ROW;
constexpr int CD_THREADS = CD_NH * CD_HALF;

template <int VAR>
static int cd_launch(bool sym, dim3 grid, size_t smem, cudaStream_t st, const float* PA, const float* PB, int nrows, int ncols,
                     int npts, int npad, int rstrip, float* out, long long ld_out, const unsigned* gate) {
    if (sym) {
        PDGN_CUDA(cudaFuncSetAttribute(cd_allpairs_kernel<CD_R, CD_NH, CD_MINB, VAR, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cd_allpairs_kernel<CD_R, CD_NH, CD_MINB, VAR, true><<<grid, CD_THREADS, smem, st>>>(PA, PB, nrows, ncols, npts, npad, rstrip, out, ld_out, gate);
    } else {
        PDGN_CUDA(cudaFuncSetAttribute(cd_allpairs_kernel<CD_R, CD_NH, CD_MINB, VAR, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cd_allpairs_kernel<CD_R, CD_NH, CD_MINB, VAR, false><<<grid, CD_THREADS, smem, st>>>(PA, PB, nrows, ncols, npts, npad, rstrip, out, ld_out, gate);
    }
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}

template <int OPT>
static int cdg_launch_opt(bool sym, dim3 grid, size_t smem, cudaStream_t st, const float* PA, const float* PB, int nrows, int ncols,
                          int npts, int npad, int rstrip, float* out, long long ld_out, const unsigned* gate) {
    if (sym) {
        PDGN_CUDA(cudaFuncSetAttribute(cd_gram_kernel<CD_R, CD_NH, CD_MINB, true, OPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cd_gram_kernel<CD_R, CD_NH, CD_MINB, true, OPT><<<grid, CD_THREADS, smem, st>>>(PA, PB, nrows, ncols, npts, npad, rstrip, out, ld_out, gate);
    } else {
        PDGN_CUDA(cudaFuncSetAttribute(cd_gram_kernel<CD_R, CD_NH, CD_MINB, false, OPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cd_gram_kernel<CD_R, CD_NH, CD_MINB, false, OPT><<<grid, CD_THREADS, smem, st>>>(PA, PB, nrows, ncols, npts, npad, rstrip, out, ld_out, gate);
    }
    PDGN_CHECK_LAUNCH();
    return PDGN_OK;
}

constexpr int CDG_OPT = 0;   // shipped inner-loop form (bit 0: unroll 2, bit 1: software-pipelined loads; profiles/r02_cd_gram_tune.txt)

static int cdg_launch(bool sym, dim3 grid, size_t smem, cudaStream_t st, const float* PA, const float* PB, int nrows, int ncols,
                      int npts, int npad, int rstrip, float* out, long long ld_out, const unsigned* gate) {
    static const char* opt = tune_env("PDGN_CDG_OPT");   // tuning hook
    const int o = opt ? atoi(opt) : CDG_OPT;
    switch (o) {
        case 1: return cdg_launch_opt<1>(sym, grid, smem, st, PA, PB, nrows, ncols, npts, npad, rstrip, out, ld_out, gate);
        case 2: return cdg_launch_opt<2>(sym, grid, smem, st, PA, PB, nrows, ncols, npts, npad, rstrip, out, ld_out, gate);
        case 3: return cdg_launch_opt<3>(sym, grid, smem, st, PA, PB, nrows, ncols, npts, npad, rstrip, out, ld_out, gate);
        case 4: return cdg_launch_opt<4>(sym, grid, smem, st, PA, PB, nrows, ncols, npts, npad, rstrip, out, ld_out, gate);
        case 8: return cdg_launch_opt<8>(sym, grid, smem, st, PA, PB, nrows, ncols, npts, npad, rstrip, out, ld_out, gate);
        case 12: return cdg_launch_opt<12>(sym, grid, smem, st, PA, PB, nrows, ncols, npts, npad, rstrip, out, ld_out, gate);
        default: return cdg_launch_opt<0>(sym, grid, smem, st, PA, PB, nrows, ncols, npts, npad, rstrip, out, ld_out, gate);
    }
}

// PDGN_B200_CD_EXACT=1 (a product switch, read once): every tile uses the direct-form kernel, whose minima are bit-identical to
// the reference's NmDistanceKernel; by default centred clouds of <= 2048 points take the Gram-form kernel (cd_kernel.cuh).
static bool cd_exact_only() {
    static const bool on = [] { const char* e = getenv("PDGN_B200_CD_EXACT"); return e && e[0] == '1'; }();
    return on;
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
    // 3 planes per cloud for the direct-form kernel + 4 for the Gram-form one + the gate words
    return ((size_t)na + (size_t)nb) * 7 * (size_t)cd_npad(npts) * sizeof(float) + 512;
}

extern "C" int pdgn_cd_allpairs(const float* A, const float* B, int na, int nb, int npts, int row0, int row1, int col0,
                                int col1, float* out, long long ld_out, void* workspace, size_t workspace_bytes,
                                void* stream) {
    PDGN_RANGE("pdgn_cd_allpairs");
    if (na < 0 || nb < 0 || npts <= 0) return PDGN_ERR_BAD_ARG;
    if (row0 < 0 || row1 > na || row0 > row1 || col0 < 0 || col1 > nb || col0 > col1) return PDGN_ERR_BAD_ARG;
    if (npts > 16384) return PDGN_ERR_UNSUPPORTED;
    const int nrows = row1 - row0, ncols = col1 - col0;
    if (nrows == 0 || ncols == 0) return PDGN_OK;  // empty tile (pointers may be null)
    if (!A || !B || !out) return PDGN_ERR_BAD_ARG;
    if (ld_out < ncols) return PDGN_ERR_BAD_ARG;
    const int npad = cd_npad(npts);
    const bool gram = npts <= CD_R * CD_HALF && !cd_exact_only();   // Gram-form kernel eligible (the gate still decides per tile)
    const size_t need = ((size_t)nrows + ncols) * (gram ? 7 : 3) * npad * sizeof(float) + (gram ? 256 : 0);
    if (!workspace || workspace_bytes < need) return PDGN_ERR_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 15) != 0) return PDGN_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    float* PA = reinterpret_cast<float*>(workspace);
    float* PB = PA + (size_t)nrows * 3 * npad;
    float* GA = PB + (size_t)ncols * 3 * npad;                       // Gram-form packs: [cloud][4][npad]
    float* GB = GA + (size_t)nrows * 4 * npad;
    unsigned* stats = reinterpret_cast<unsigned*>(GB + (size_t)ncols * 4 * npad);   // [0] max |p|^2, [1] mean NN d^2, [2] gate

    // same set against itself (the rr / ss matrices of compute_all_metrics, evaluation_metrics.py:187-188): pack once,
    // compute the upper triangle, mirror it
    const bool sym = (A == B) && row0 == col0 && row1 == col1 && nrows > 1;
    dim3 pb(256), pga((npad + 255) / 256, nrows), pgb((npad + 255) / 256, ncols);
    cd_pack_kernel<<<pga, pb, 0, st>>>(A, row0, npts, npad, PA);
    PDGN_CHECK_LAUNCH();
    if (sym) {
        PB = PA;
    } else {
        cd_pack_kernel<<<pgb, pb, 0, st>>>(B, col0, npts, npad, PB);
        PDGN_CHECK_LAUNCH();
    }
    const unsigned* gate = nullptr;
    if (gram) {
        PDGN_CUDA(cudaMemsetAsync(stats, 0, 3 * sizeof(unsigned), st));
        cd_pack_gram_kernel<<<pga, pb, 0, st>>>(A, row0, npts, npad, GA, stats);
        PDGN_CHECK_LAUNCH();
        if (sym) {
            GB = GA;
        } else {
            cd_pack_gram_kernel<<<pgb, pb, 0, st>>>(B, col0, npts, npad, GB, stats);
            PDGN_CHECK_LAUNCH();
        }
        cd_scale_kernel<<<1, 256, 0, st>>>(A, row0, npts, stats);
        PDGN_CHECK_LAUNCH();
        cd_gate_kernel<<<1, 1, 0, st>>>(stats, npts, stats + 2);
        PDGN_CHECK_LAUNCH();
        gate = stats + 2;
    }

    const int spairs = (nrows + CD_NH - 1) / CD_NH;
    if (spairs > 65535) return PDGN_ERR_UNSUPPORTED;
    // enough CTAs that the last partial wave is a small fraction of the run; each CTA walks `rstrip` B clouds
    static const int waves = [] {  // tuning hook
        const char* e = tune_env("PDGN_CD_WAVES");
        const int v = e ? atoi(e) : 0;
        return v > 0 ? v : 64;
    }();
    const int target = (sym ? 2 : 1) * waves * CD_MINB * cd_num_sms();  // ~64 waves: tail <= ~1.5 %; sym: half the CTAs exit at once
    int strips = (target + spairs - 1) / spairs;
    if (strips > ncols) strips = ncols;
    if (strips < 1) strips = 1;
    const int rstrip = (ncols + strips - 1) / strips;
    strips = (ncols + rstrip - 1) / rstrip;
    const dim3 grid(strips, spairs);
    static const bool force_big = tune_env("PDGN_CD_ATOMIC_COLMIN") != nullptr;  // tuning hook: the shared-atomic variant for every size
    int rc;
    // two launches, one of which returns at once: the gate is decided on the device, the host never waits for it
    if (npts <= CD_R * CD_HALF && !force_big)
        rc = cd_launch<CD_VARIANT>(sym, grid, cd_smem_bytes<CD_NH, CD_VARIANT>(npad), st, PA, PB, nrows, ncols, npts, npad, rstrip, out, ld_out, gate);
    else
        rc = cd_launch<CD_VARIANT_BIG>(sym, grid, cd_smem_bytes<CD_NH, CD_VARIANT_BIG>(npad), st, PA, PB, nrows, ncols, npts, npad, rstrip, out, ld_out, gate);
    if (rc != PDGN_OK) return rc;
    if (gram) {
        rc = cdg_launch(sym, grid, cdg_smem_bytes<CD_NH>(npad), st, GA, GB, nrows, ncols, npts, npad, rstrip, out, ld_out, gate);
        if (rc != PDGN_OK) return rc;
    }
    if (sym) {
        cd_mirror_kernel<<<dim3((nrows + 31) / 32, (nrows + 7) / 8), dim3(32, 8), 0, st>>>(out, nrows, ld_out);
        PDGN_CHECK_LAUNCH();
    }
    return PDGN_OK;
}

extern "C" int pdgn_cd_allpairs_host(const float* A_host, const float* B_host, int na, int nb, int npts, int row0, int row1,
                                     int col0, int col1, float* out_host, long long ld_out, void* stream) {
    PDGN_RANGE("pdgn_cd_allpairs_host");
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
