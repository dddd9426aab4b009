// pull_stream.cu -- streaming backward of the gathers: grouping, edge-feature and interpolation backward as a deterministic PULL
// through the inverse index (csr_build_kernel, gather.cu) instead of the reference's float atomicAdd scatter
// (grouping_cuda_kernel.cu:28-46, interpolation_cuda_kernel.cu:90-114, the index_select backward of PDGNet_v2.py:461-477).
//
// A persistent grid (one CTA per SM) walks the flattened list of (batch element, channel chunk) items in contiguous, balanced
// ranges.  The rows of an item arrive by TMA bulk copy (cp.async.bulk + mbarrier) into a ring of shared-memory stages; the
// consumers are the CTA's 32 warps, each synchronising ONLY through the stage's full / empty mbarriers (no block-wide barrier
// in the steady state: round 1's kernel spent three __syncthreads and an exposed global round trip per 80 KB chunk).
// Thread layout: the 1024 threads form G = 1024 / tp channel groups of tp threads (tp = the target count rounded up to a power
// of two), so a 128-point stage of the generator still fills the CTA; thread (g, tl) owns target tl and the CC channels
// g*CC.. of every chunk.  The inverse-index list of a target is the same for every chunk of a batch element: it is kept in
// registers as byte offsets (EC entries), and the per-chunk work is one LDS + one FADD / FFMA per (entry, channel).  That only
// holds if the address of channel ch's row is `entry offset + immediate`: the staged rows therefore sit in fixed-size SLOTS
// (template parameter STRIDE bytes; 0 = rows packed at their own length, address arithmetic at run time), one bulk copy per
// row.  The loop over the entries is bounded by the warp's longest list.  Lists longer than EC (hubs of a kNN graph) are
// summed by whole warps after the register pass (lane l takes entries l, l+32, ..; fixed butterfly), their targets collected
// once per batch element.  Every (channel, target) is written by exactly one thread, in a fixed order of additions:
// deterministic, unlike the reference.
// MODE 0: rows = grad_out[b,ch,:].  MODE 1: rows = g1 = grad_ee[b,c+ch,:] plus the central term sum_s (g0 - g1)[i,s] from the
// g0 rows staged beside them.  MODE 2 (interpolation backward): the lists hold entries e = 3 j + t of idx[b, 3 rowlen], the
// contribution of an entry is rows[ch][e / 3] * wgt[b][e], accumulated with one FFMA (the reference rounds the product before
// its atomicAdd; the fused form is the more accurate one and the sum order differs from the reference's anyway).
#include "pull.cuh"

namespace pdgn {

constexpr int PS_MAXST = 4;
constexpr int PS_LCAP = 4096;   // entries of long lists cached in shared memory per batch element (beyond: read from global)

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// consumer-side wait with a suspend-time hint: the warp sleeps in hardware instead of spinning through the issue slots the
// working warps need (round-2 ncu: a quarter of the executed instructions of the un-hinted kernel were try_wait retries)
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAITS_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONES_%=;\n\t"
        "bra WAITS_%=;\n\t"
        "DONES_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// sum_s (q0[s] - q1[s]) over the k contiguous values of one target, in index order; the widest shared-memory load the
// alignment allows (a warp's 32 segments of k = 10 or 20 floats are bank-conflict free as 8- / 16-byte loads, 2-way as scalars)
__device__ __forceinline__ float central_sum(const float* __restrict__ q0, const float* __restrict__ q1, int k) {
    float acc = 0.f;
    if ((k & 3) == 0) {
        for (int s = 0; s < (k >> 2); ++s) {
            const float4 a = reinterpret_cast<const float4*>(q0)[s], d = reinterpret_cast<const float4*>(q1)[s];
            acc += a.x - d.x;
            acc += a.y - d.y;
            acc += a.z - d.z;
            acc += a.w - d.w;
        }
    } else if ((k & 1) == 0) {
        for (int s = 0; s < (k >> 1); ++s) {
            const float2 a = reinterpret_cast<const float2*>(q0)[s], d = reinterpret_cast<const float2*>(q1)[s];
            acc += a.x - d.x;
            acc += a.y - d.y;
        }
    } else {
        for (int s = 0; s < k; ++s) acc += q0[s] - q1[s];
    }
    return acc;
}

template <int MODE> struct PullCfg { static constexpr int EC = MODE == 2 ? 16 : 24; };

template <int CC, int MODE, int STRIDE>
__global__ void __launch_bounds__(1024, 1) pull_stream_kernel(const float* __restrict__ src, const int* __restrict__ offs,
                                                          const int* __restrict__ pos, int b, int c, int ntargets, int rowlen,
                                                          int k, int tp, int nstage, float* __restrict__ dst,
                                                          const float* __restrict__ wgt, int parts, int longcap, int dbg) {
    constexpr int NR = MODE == 1 ? 2 : 1;          // row sets per channel (MODE 1: g0 and g1)
    constexpr int EC = PullCfg<MODE>::EC;          // list entries cached in registers across the chunks
    extern __shared__ __align__(128) unsigned char praw[];   // [nstage][NR][SC slots] | long-target list
    __shared__ uint64_t full[PS_MAXST], empty[PS_MAXST];
    __shared__ int nlong;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = 1024 / tp, SC = G * CC;          // channel groups, channels per stage
    const int nchunks = (c + SC - 1) / SC;
    const long long total = (long long)b * nchunks;
    // Item range of this CTA.  parts > 0 (b <= #SMs): CTA = (batch element, part of its chunks), so the lists are loaded once,
    // at kernel start, before the memory system fills up with bulk copies (a list reload in mid-stream is three dependent
    // round trips behind ~25 MB of queued TMA traffic: 13-19 us measured).  parts == 0: balanced split of the flattened list.
    int it0, nloc;
    if (parts > 0) {
        const int bzc = blockIdx.x / parts, part = blockIdx.x - bzc * parts;
        const int c_lo = nchunks * part / parts, c_hi = nchunks * (part + 1) / parts;
        it0 = bzc * nchunks + c_lo;
        nloc = c_hi - c_lo;
    } else {
        it0 = (int)(total * blockIdx.x / gridDim.x);
        nloc = (int)(total * (blockIdx.x + 1) / gridDim.x) - it0;
    }
    if (nloc <= 0) return;
    const uint32_t rowb = (uint32_t)rowlen * 4u;                       // bytes of a row
    const uint32_t slot = STRIDE ? (uint32_t)STRIDE : rowb;           // bytes between staged rows
    const uint32_t set_b = (uint32_t)SC * slot, stage_b = NR * set_b; // bytes of a row set / of a stage
    const uint32_t praw_s = smem_u32(praw);
    // after the ring: targets with long lists, where each one's entries sit in the shared cache below (-1: not cached), the cache
    int* longlist = reinterpret_cast<int*>(praw + (size_t)nstage * stage_b);
    int* loff = longlist + longcap;
    int* llen = loff + longcap;
    unsigned short* lce = reinterpret_cast<unsigned short*>(llen + longcap);          // [PS_LCAP] word offsets inside a row
    float* lcw = reinterpret_cast<float*>(lce + PS_LCAP);                             // [PS_LCAP] weights (MODE 2 only)
    const int entries = MODE == 2 ? 3 * rowlen : rowlen;
    const int g = tid / tp, tl = tid - g * tp;

    auto issue = [&](int item, int s) {            // thread 0 only
        const int bz = item / nchunks, ch0 = (item - bz * nchunks) * SC;
        const int nrow = min(SC, c - ch0);
        unsigned char* d = praw + (size_t)s * stage_b;
        mbar_expect_tx(&full[s], NR * (unsigned)nrow * rowb);
        const float* s0 = src + ((size_t)bz * (MODE == 1 ? 2 : 1) * c + ch0) * rowlen;
        if (STRIDE) {
            for (int r = 0; r < nrow; ++r) {
                bulk_g2s(d + (size_t)r * slot, s0 + (size_t)r * rowlen, rowb, &full[s]);
                if (MODE == 1) bulk_g2s(d + set_b + (size_t)r * slot, s0 + ((size_t)c + r) * rowlen, rowb, &full[s]);
            }
        } else {
            bulk_g2s(d, s0, (unsigned)nrow * rowb, &full[s]);
            if (MODE == 1) bulk_g2s(d + set_b, s0 + (size_t)c * rowlen, (unsigned)nrow * rowb, &full[s]);
        }
    };
    if (tid == 0) {
        for (int s = 0; s < nstage; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 32);
        }
        mbar_fence_init();
        nlong = 0;
    }
    __syncthreads();
    if (tid == 0)
        for (int i = 0; i < nstage - 1 && i < nloc; ++i) issue(it0 + i, i);

    int cur_bz = -1, clen = 0, wlen = 0;
    uint32_t ecp[EC / 2];                          // cached entries: WORD offset inside a staged row, two 16-bit values per register
    float wc[MODE == 2 ? EC : 1];
    // running state instead of divisions: (bz, chunk) of the item, ring stage + phase of the item and of its predecessor
    int bz = it0 / nchunks, chunk = it0 - bz * nchunks;
    int s = 0, ph = 0, sp = nstage - 1, php = 1;   // (sp, php): stage and phase of item i-1 (valid from i = 1)

    for (int i = 0; i < nloc; ++i) {
        const int ch0 = chunk * SC;
        if (tid == 0 && i + nstage - 1 < nloc) {   // producer duty: refill the stage that item i-1 has just left
            if (i >= 1) mbar_wait(&empty[sp], (unsigned)php);
            fence_proxy_async();
            issue(it0 + i + nstage - 1, i >= 1 ? sp : nstage - 1);
        }
        const int* ob = offs + (size_t)bz * (ntargets + 1);
        const int* pb = pos + (size_t)bz * entries;
        const float* wb = MODE == 2 ? wgt + (size_t)bz * entries : nullptr;
        if (bz != cur_bz) {                        // block-uniform: new batch element, new lists
            __syncthreads();                       // every warp has finished the previous element's long-target list
            cur_bz = bz;
            if (tid == 0) nlong = 0;
            __syncthreads();
            int ca = 0;
            clen = 0;
            if (tl < ntargets && !(dbg & 4)) {
                ca = ob[tl];
                clen = ob[tl + 1] - ca;
            }
#pragma unroll
            for (int u = 0; u < EC; u += 2) {
                const int e0 = (u < clen && clen <= EC) ? __ldg(pb + ca + u) : 0;
                const int e1 = (u + 1 < clen && clen <= EC) ? __ldg(pb + ca + u + 1) : 0;
                ecp[u >> 1] = (uint32_t)(MODE == 2 ? e0 / 3 : e0) | ((uint32_t)(MODE == 2 ? e1 / 3 : e1) << 16);
                if (MODE == 2) {
                    wc[u] = (u < clen && clen <= EC) ? __ldg(wb + e0) : 0.f;
                    wc[u + 1] = (u + 1 < clen && clen <= EC) ? __ldg(wb + e1) : 0.f;
                }
            }
            wlen = __reduce_max_sync(kFull, clen <= EC ? clen : 0);
            if (dbg & 1) wlen = 0;
            if (g == 0 && !(dbg & 4))
                for (int p = tl; p < ntargets; p += tp)
                    if (ob[p + 1] - ob[p] > EC) longlist[atomicAdd(&nlong, 1)] = p;   // order irrelevant: one warp sums a whole list
            __syncthreads();
            // The long lists are walked once per chunk: copy them (and their weights) into shared memory now.  Read from global
            // memory in the chunk loop, a single hub made its CTA the kernel's tail (two dependent round trips per chunk).
            const int nlg = nlong;
            if (nlg > 0) {
                if (warp == 0) {
                    int run = 0;
                    for (int j0 = 0; j0 < nlg; j0 += 32) {
                        const int j = j0 + lane;
                        int len = 0;
                        if (j < nlg) {
                            const int p = longlist[j];
                            len = ob[p + 1] - ob[p];
                        }
                        int incl = len;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int v = __shfl_up_sync(kFull, incl, o);
                            if (lane >= o) incl += v;
                        }
                        const int off = run + incl - len;
                        if (j < nlg) {
                            loff[j] = off + len <= PS_LCAP ? off : -1;
                            llen[j] = len;
                        }
                        run += __shfl_sync(kFull, incl, 31);
                    }
                }
                __syncthreads();
                for (int j = warp; j < nlg; j += 32) {
                    const int off = loff[j];
                    if (off < 0) continue;
                    const int p = longlist[j], a = ob[p], len = ob[p + 1] - a;
                    for (int q = lane; q < len; q += 32) {
                        const int e = __ldg(pb + a + q);
                        lce[off + q] = (unsigned short)(MODE == 2 ? e / 3 : e);
                        if (MODE == 2) lcw[off + q] = __ldg(wb + e);
                    }
                }
                __syncthreads();
            }
        }
        float* dst_i = dst + ((size_t)bz * c + ch0) * ntargets;        // channel ch0 of this batch element
        const uint32_t stage_s = praw_s + (uint32_t)s * stage_b;       // shared address of the stage (g0 set in MODE 1)
        const unsigned char* stage = praw + (size_t)s * stage_b;
        const int chg = g * CC, ccg = min(CC, c - ch0 - chg);           // this thread's channels of the chunk
        const bool own = ccg > 0 && tl < ntargets && clen <= EC;
        // the caller's buffer is ADDED into: read it before waiting for the stage (coalesced), so the round trip overlaps the wait
        float old[CC];
#pragma unroll
        for (int ch = 0; ch < CC; ++ch) old[ch] = 0.f;
        float* const dp = dst_i + (size_t)chg * ntargets + tl;         // this thread's (first channel, target) of the chunk
        if (own && !(dbg & 2)) {
            const float* q = dp;
#pragma unroll
            for (int ch = 0; ch < CC; ++ch) {                           // pointer walk: two integer adds per channel
                if (ch < ccg) old[ch] = *q;
                q += ntargets;
            }
        }
        // same for the first long-list target of this warp (lane 0 writes it): its round trip must not sit between the stage
        // becoming full and the warp releasing it
        const int nl = (dbg & 8) ? 0 : nlong * G;
        float lold[CC];
#pragma unroll
        for (int ch = 0; ch < CC; ++ch) lold[ch] = 0.f;
        if (warp < nl && lane == 0 && !(dbg & 2)) {
            const int jj = warp / G, gg = warp - jj * G;
            const int lchg = gg * CC, lccg = min(CC, c - ch0 - lchg);
            const float* q = dst_i + (size_t)lchg * ntargets + longlist[jj];
#pragma unroll
            for (int ch = 0; ch < CC; ++ch) {
                if (ch < lccg) lold[ch] = *q;
                q += ntargets;
            }
        }
        mbar_wait_sleep(&full[s], (unsigned)ph);
        if (own) {
            // ---- the thread's own target, list in registers
            const uint32_t rows_s = stage_s + (uint32_t)chg * slot + (MODE == 1 ? set_b : 0);   // rows the lists index
            float acc[CC];
#pragma unroll
            for (int ch = 0; ch < CC; ++ch) {
                acc[ch] = 0.f;
                if (MODE == 1) {
                    const float* q0 = reinterpret_cast<const float*>(stage + (size_t)(chg + ch) * slot) + tl * k;
                    acc[ch] = central_sum(q0, reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(q0) + set_b), k);
                }
            }
#pragma unroll
            for (int u0 = 0; u0 < EC; u0 += 4) {
                if (u0 >= wlen) break;                                  // warp-uniform
#pragma unroll
                for (int u = u0; u < u0 + 4; ++u) {
                    if (u < clen) {
                        const uint32_t w16 = (u & 1) ? (ecp[u >> 1] >> 16) : (ecp[u >> 1] & 0xffffu);
                        const uint32_t ea = rows_s + (w16 << 2);
#pragma unroll
                        for (int ch = 0; ch < CC; ++ch) {
                            const float v = lds_f32(ea + (uint32_t)ch * slot);   // STRIDE != 0: an immediate offset
                            acc[ch] = MODE == 2 ? __fmaf_rn(v, wc[u], acc[ch]) : acc[ch] + v;
                        }
                    }
                }
            }
            float* q = dp;
#pragma unroll
            for (int ch = 0; ch < CC; ++ch) {
                if (ch < ccg && !(dbg & 2)) *q = old[ch] + acc[ch];
                q += ntargets;
            }
        }
        if (ccg > 0) {
            // ---- further targets of this thread (more than tp targets): lists from global memory
            for (int p = tl + tp; p < ntargets; p += tp) {
                const int a = ob[p], len = ob[p + 1] - a;
                if (len > EC) continue;                                 // summed by a whole warp below
                const float* rows = reinterpret_cast<const float*>(stage + (MODE == 1 ? set_b : 0) + (size_t)chg * slot);
                const int sf = (int)(slot >> 2);
                float* dp = dst_i + (size_t)chg * ntargets + p;
                float acc[CC];
#pragma unroll
                for (int ch = 0; ch < CC; ++ch)
                    acc[ch] = MODE == 1 ? central_sum(rows + ch * sf + p * k - (set_b >> 2), rows + ch * sf + p * k, k) : 0.f;
                pull_list<CC>(pb, a, a + len, acc, [&](int ch, int e) {
                    return MODE == 2 ? __fmul_rn(rows[ch * sf + e / 3], __ldg(wb + e)) : rows[ch * sf + e];
                });
#pragma unroll
                for (int ch = 0; ch < CC; ++ch)
                    if (ch < ccg) dp[(size_t)ch * ntargets] += acc[ch];
            }
        }
        // long lists: one warp per (target, channel group)
        for (int w = warp; w < nl; w += 32) {
            const int jj = w / G, gg = w - jj * G;
            const int p = longlist[jj];
            const int chg = gg * CC, ccg = min(CC, c - ch0 - chg);
            if (ccg <= 0) continue;
            const float* r0 = reinterpret_cast<const float*>(stage + (size_t)chg * slot);
            const float* rows = reinterpret_cast<const float*>(stage + (MODE == 1 ? set_b : 0) + (size_t)chg * slot);
            const int sf = (int)(slot >> 2);
            float acc[CC];
#pragma unroll
            for (int ch = 0; ch < CC; ++ch) acc[ch] = 0.f;
            const int off = loff[jj];
            if (off >= 0) {                        // same order of additions as the global-memory walk below
                const int len = llen[jj];
                for (int q = lane; q < len; q += 32) {
                    const int w16 = lce[off + q];
#pragma unroll
                    for (int ch = 0; ch < CC; ++ch)
                        acc[ch] += MODE == 2 ? __fmul_rn(rows[ch * sf + w16], lcw[off + q]) : rows[ch * sf + w16];
                }
#pragma unroll
                for (int ch = 0; ch < CC; ++ch)
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc[ch] += __shfl_xor_sync(kFull, acc[ch], o);
            } else {
                pull_list_warp<CC>(pb, ob[p], ob[p + 1], lane, acc, [&](int ch, int e) {
                    return MODE == 2 ? __fmul_rn(rows[ch * sf + e / 3], __ldg(wb + e)) : rows[ch * sf + e];
                });
            }
            if (lane == 0) {
#pragma unroll
                for (int ch = 0; ch < CC; ++ch) {
                    if (ch < ccg) {
                        const float cen = MODE == 1 ? central_sum(r0 + ch * sf + p * k, rows + ch * sf + p * k, k) : 0.f;
                        float* o = dst_i + (size_t)(chg + ch) * ntargets + p;
                        if (!(dbg & 2)) *o = (w == warp ? lold[ch] : *o) + (cen + acc[ch]);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);     // this warp is done with the stage
        sp = s;
        php = ph;
        if (++s == nstage) {
            s = 0;
            ph ^= 1;
        }
        if (++chunk == nchunks) {
            chunk = 0;
            ++bz;
        }
    }
}

// Slot sizes the kernel is instantiated for (bytes): interpolation rows of <= 2048 points, k = 10 graphs of <= 512 / <= 1024
// points; anything else packs its rows (STRIDE 0).
constexpr int PS_SLOT_A = 8 * 1024, PS_SLOT_B = 20 * 1024, PS_SLOT_C = 40 * 1024;

template <int MODE, int STRIDE>
static int pull_stream_try(const float* src, const int* offs, const int* pos, int b, int c, int ntargets, int rowlen, int k,
                           float* dst, cudaStream_t st, bool* launched, const float* wgt) {
    constexpr int NR = MODE == 1 ? 2 : 1;
    const size_t slot = STRIDE ? (size_t)STRIDE : (size_t)rowlen * 4;
    int tp = 32;
    while (tp < 1024 && tp < ntargets) tp <<= 1;
    int G = 1024 / tp;
    while (G > 1 && G > c) G >>= 1;                 // no more channel groups than channels
    tp = 1024 / G;
    const long long entries = (long long)rowlen * (MODE == 2 ? 3 : 1);
    const int longcap = (int)(entries / (PullCfg<MODE>::EC + 1)) + 1;   // no more targets than this can have a long list
    const size_t long_bytes = (size_t)longcap * 12 + (size_t)PS_LCAP * (MODE == 2 ? 6 : 2) + 16;
    const size_t budget = 212 * 1024 - long_bytes;
    static const char* cc_env = tune_env("PDGN_PULL_CC");
    static const char* st_env = tune_env("PDGN_PULL_STAGES");
    const int want_st = st_env ? atoi(st_env) : 2;
    static const char* dbg_env = tune_env("PDGN_PULL_DBG");      // ablation only (bits): 1 skip the list walk, 2 skip the dst
    // read-modify-write, 4 skip the list loads, 8 skip the long-list warps, 32 inverse-index build only (profiles/r02_pull_stream_ablation.txt)
    const int dbg = dbg_env ? atoi(dbg_env) : 0;
    int CCsel = 0, nstage = 0;
    for (int pass = 0; pass < 2 && !CCsel; ++pass) {            // first a ring of >= want_st stages, then settle for 2
        for (int cc = cc_env ? atoi(cc_env) : (MODE == 1 ? 4 : 8); cc >= 1; cc >>= 1) {
            if (cc != 1 && G * cc > c) continue;
            const size_t stage_bytes = (size_t)NR * G * cc * slot;
            const int fit = (int)(budget / stage_bytes);
            const long long items = (long long)b * ((c + G * cc - 1) / (G * cc));   // enough items for every SM before widening
            if (fit >= (pass == 0 ? want_st : 2) && (items >= 2 * 148 || cc == 1)) {
                CCsel = cc;
                nstage = fit < PS_MAXST ? fit : PS_MAXST;
                break;
            }
        }
    }
    if (!CCsel) return PDGN_OK;                     // rows too long for two stages
    const size_t stage_bytes = (size_t)NR * G * CCsel * slot;
    if ((size_t)NR * G * CCsel * rowlen * 4 >= (1u << 20)) return PDGN_OK;   // mbarrier tx-count range
    const long long items = (long long)b * ((c + G * CCsel - 1) / (G * CCsel));
    const int nch = (c + G * CCsel - 1) / (G * CCsel);
    // CTAs per batch element: 0 = balanced split of the flattened item list over all SMs (default: measured 4 % faster at B = 35
    // than 4 x 35 = 140 batch-aligned CTAs, which idle 8 SMs to save one list reload per CTA); PDGN_PULL_PARTS=1 aligns.
    static const char* parts_env = tune_env("PDGN_PULL_PARTS");
    int parts = (parts_env && parts_env[0] == '1' && b <= 148) ? 148 / b : 0;
    if (parts > nch) parts = nch;
    const int grid = parts > 0 ? b * parts : (int)(items < 148 ? items : 148);
    const size_t smem = (size_t)nstage * stage_bytes + long_bytes;
    if (dbg & 32) {   // ablation: inverse-index build only
        *launched = true;
        return PDGN_OK;
    }
#define PDGN_LAUNCH_PULL(CC_)                                                                                                          \
    do {                                                                                                                               \
        PDGN_CUDA(cudaFuncSetAttribute(pull_stream_kernel<CC_, MODE, STRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        pull_stream_kernel<CC_, MODE, STRIDE><<<grid, 1024, smem, st>>>(src, offs, pos, b, c, ntargets, rowlen, k, tp, nstage, dst, wgt, parts, longcap, dbg); \
    } while (0)
    if (CCsel == 8 && MODE != 1) PDGN_LAUNCH_PULL(MODE != 1 ? 8 : 4);
    else if (CCsel == 4) PDGN_LAUNCH_PULL(4);
    else if (CCsel == 2) PDGN_LAUNCH_PULL(2);
    else PDGN_LAUNCH_PULL(1);
#undef PDGN_LAUNCH_PULL
    PDGN_CHECK_LAUNCH();
    *launched = true;
    return PDGN_OK;
}

template <int MODE>
static int pull_stream_mode(const float* src, const int* offs, const int* pos, int b, int c, int ntargets, int rowlen, int k,
                            float* dst, cudaStream_t st, bool* launched, const float* wgt) {
    *launched = false;
    if ((rowlen & 3) != 0 || (reinterpret_cast<uintptr_t>(src) & 15) != 0 || ntargets < 1 || rowlen < 4) return PDGN_OK;
    const size_t rowb = (size_t)rowlen * 4;
    static const char* slot_env = tune_env("PDGN_PULL_SLOT");     // "0": always pack the rows (ablation)
    const bool slots = !(slot_env && slot_env[0] == '0');
    int rc = PDGN_OK;
    // smallest slot that holds a row; when the slotted layout does not fit a two-stage ring the packed one may still
    if (slots && rowb <= (size_t)PS_SLOT_A) rc = pull_stream_try<MODE, PS_SLOT_A>(src, offs, pos, b, c, ntargets, rowlen, k, dst, st, launched, wgt);
    else if (slots && rowb <= (size_t)PS_SLOT_B) rc = pull_stream_try<MODE, PS_SLOT_B>(src, offs, pos, b, c, ntargets, rowlen, k, dst, st, launched, wgt);
    else if (slots && rowb <= (size_t)PS_SLOT_C) rc = pull_stream_try<MODE, PS_SLOT_C>(src, offs, pos, b, c, ntargets, rowlen, k, dst, st, launched, wgt);
    if (rc != PDGN_OK || *launched) return rc;
    return pull_stream_try<MODE, 0>(src, offs, pos, b, c, ntargets, rowlen, k, dst, st, launched, wgt);
}

int pull_stream_launch(int mode, const float* src, const int* offs, const int* pos, int b, int c, int ntargets, int rowlen, int k,
                       float* dst, cudaStream_t st, bool* launched, const float* wgt) {
    if (mode == 0) return pull_stream_mode<0>(src, offs, pos, b, c, ntargets, rowlen, k, dst, st, launched, wgt);
    if (mode == 1) return pull_stream_mode<1>(src, offs, pos, b, c, ntargets, rowlen, k, dst, st, launched, wgt);
    return pull_stream_mode<2>(src, offs, pos, b, c, ntargets, rowlen, k, dst, st, launched, wgt);
}

}  // namespace pdgn
