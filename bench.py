#!/usr/bin/env python
"""bench.py -- headline benchmark of pdgn_b200: all-pairs Chamfer-distance matrix throughput.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[3], the configuration the metric "CD pairs/s" is quoted on; it fits one GPU):
1000 generated x 1000 reference synthetic clouds of 2048 points (S: unit sphere + N(0,0.01^2) radial noise, seeds 0/1).
One step = the whole 1000x1000 matrix (10^6 cloud pairs, 4.19e12 point-pair distances).  With N ranks the pair grid
is 2-D tiled (pdgn_b200.dist) and the per-pair scalars are all-gathered over NCCL inside the step => strong scaling.

One JSON line on rank 0:
  value          cloud pairs/s, device-resident inputs, CUDA events on the launch stream, max over ranks
  e2e            same metric through the public API (_pairwise_EMD_CD_) from pinned HOST tensors, H2D + D2H inside
  roofline       FP32-SIMT roofline of the dominant kernel (cd_gram_kernel): 6 FMA-pipe instructions per point pair
  cpu_baseline   the reference's CPU formulation (oracle.torch_ref.pairwise_cd = distChamfer loop) on a bounded sample
  knnquery       secondary metric of BASELINE.json (Mqueries/s, k=20, B=35 x 2048) measured in the same run
  reference_gpu  "reference on B200" (SURVEY.md 8d): the reference's torch Gram-form distChamfer loop and its NNDistance
                 kernel (recompiled for sm_100a) on THIS GPU, on a bounded sample of the same workload; plus the reference's
                 real _pairwise_EMD_CD_ (CD + EMD) against ours with EMD on
  cfg3           BASELINE configs[2]: loss side of one G step (6 x get_local_pair, B=35) fwd+bwd, the four feature-space kNN
                 stages fwd+bwd, and one generator step of the reference's own PointGenerator through the drop-in
  cfg5           BASELINE configs[4]: knnquery k=32 on [8,16384,3] / [35,16384,3]; with --gpus 8 also one 4096x4096 CD matrix
  eval_emd       compute_all_metrics (3 CD + 3 EMD matrices, 1000 x 1000 clouds) from host tensors, end to end
`--impl reference` times only the CPU formulation (rank 0; other ranks exit 0).
"""
import argparse
import datetime
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CLOUDS = 1000
N_PTS = 2048
POINT_PAIRS_PER_CLOUD_PAIR = N_PTS * N_PTS
FMA_PIPE_INSTR_PER_POINT_PAIR = 6   # 3 FADD + 1 FMUL + 2 FFMA
FLOP_PER_POINT_PAIR = 8
METRIC = "cd_cloud_pairs_per_s"
UNIT = "cloud-pairs/s"
WORKLOAD = "allpairs_cd_1000x1000_clouds_2048pts"


def make_clouds(seed, n=N_CLOUDS, npts=N_PTS):
    import numpy as np
    import torch
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(size=(n, npts, 3))
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    v *= 1.0 + 0.01 * rng.standard_normal(size=(n, npts, 1))
    return torch.from_numpy(v.astype(np.float32))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", os.environ.get("PDGN_BENCH_SMI_MS", "100")], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self, window=None):
        """window = (t0, t1) wall-clock seconds: only samples taken inside it are used (the sampler is started before the
        warm-up so that nvidia-smi's own start-up does not fall into the timed region)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        rows = []
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                vals = (float(f[1]), float(f[2]), float(f[3]))
            except ValueError:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
            except ValueError:
                ts = None
            flags = [name for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9])
                     if val.lower().startswith("active")]
            rows.append((ts, vals, flags))
        inside = [r for r in rows if window is not None and r[0] is not None and window[0] <= r[0] <= window[1]]
        used = inside if inside else rows  # unparsable timestamps: fall back to every sample (warm-up included)
        sm, mx, pw, reasons = [], [], [], set()
        for _, vals, flags in used:
            sm.append(vals[0]); mx.append(vals[1]); pw.append(vals[2])
            reasons.update(flags)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


_CPU_KIND = None


def _cpu_pairwise_cd():
    """(function, kind): the CD half of the reference's _pairwise_EMD_CD_ loop (evaluation_metrics.py:85-121).  Its EMD half is
    CUDA-only, so the loop is the restated one (oracle.torch_ref.pairwise_cd); when the reference tree is staged
    (baseline/_ref/PDGN) the distance function inside it is the reference's OWN distChamfer (:35-45) => kind "reference",
    otherwise the restated one, pinned bit-for-bit by tests/golden => kind "port"."""
    global _CPU_KIND
    from oracle import torch_ref
    fn, kind = torch_ref.pairwise_cd, "port"
    try:
        from oracle import ref_tree
        if ref_tree.available():
            own = ref_tree.load_reference().evaluation_metrics.distChamfer

            def fn(sample_pcs, ref_pcs, batch_size, _loop=torch_ref.pairwise_cd):
                saved = torch_ref.dist_chamfer
                torch_ref.dist_chamfer = own
                try:
                    return _loop(sample_pcs, ref_pcs, batch_size)
                finally:
                    torch_ref.dist_chamfer = saved
            kind = "reference"
    except Exception:
        fn, kind = torch_ref.pairwise_cd, "port"
    _CPU_KIND = kind
    return fn, kind


def cpu_reference_rate(n_sample, n_ref, repeats=1):
    """Reference CPU formulation on a bounded sample of the workload: n_sample x n_ref cloud pairs of 2048 points,
    batch_size 50 as in the README's test, all host threads."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    pairwise_cd, _ = _cpu_pairwise_cd()
    smp = make_clouds(0, n_sample)
    ref = make_clouds(1, n_ref)
    pairwise_cd(smp[:1], ref[: min(n_ref, 10)], 50)  # warm-up (MKL init, page-in)
    t0 = time.perf_counter()
    for _ in range(repeats):
        pairwise_cd(smp, ref, 50)
    dt = (time.perf_counter() - t0) / repeats
    return n_sample * n_ref / dt, dt


def traffic_from_profile(nc, world):
    """DRAM bytes (read+write) of one launch of the dominant kernel from the committed ncu --set full capture
    (profiles/cd_allpairs_traffic.json); only valid for the configuration that was captured (1000 clouds, 1 GPU)."""
    if nc != N_CLOUDS or world != 1:
        return None
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "cd_allpairs_traffic.json")))
        return {"bytes": t["dram_bytes_read"] + t["dram_bytes_write"], "unit": "B per launch", "source": t["source"],
                "note": "kernel is FP32-issue bound; DRAM traffic is ~0.002 % of what HBM could move in the kernel's duration"}
    except (OSError, ValueError, KeyError):
        return None


def run_reference(args, rank):
    """--impl reference: the reference's CPU path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    n_s, n_r = 8, 50  # 400 cloud pairs per step (~3.5 s on 16 cores): bounded sample of the 1000x1000 workload
    for _ in range(args.warmup):
        cpu_reference_rate(1, 10)
    rates, times = [], []
    for _ in range(args.steps):
        r, dt = cpu_reference_rate(n_s, n_r)
        rates.append(r); times.append(dt)
    value = n_s * n_r * args.steps / sum(times)
    cores = os.cpu_count() or 1
    sample = "%d x %d cloud pairs of 2048 points per step (of the 1000 x 1000 workload), batch_size 50" % (n_s, n_r)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "reference CPU formulation (torch bmm Gram form + min, distChamfer loop): the reference's own "
                   "distChamfer from the staged tree baseline/_ref/PDGN inside the CD half of its pairwise loop (kind 'reference'), or the "
                   "restatement in oracle/torch_ref.py pinned bit-exactly by tests/golden when the tree is not staged (kind 'port')"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": _CPU_KIND or "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clouds", type=int, default=N_CLOUDS, help="debug: smaller matrix (the reported workload is 1000)")
    ap.add_argument("--no-extras", action="store_true", help="skip the knnquery / gather / cpu_baseline side measurements")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner (and any NCCL_DEBUG output) on stdout while the
        # communicator is created, so file descriptor 1 points at stderr until the first collective has completed
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    from pdgn_b200 import _build
    if rank == 0:
        _build.build()
    if world > 1:
        dist.barrier()
    # the headline metric is CD cloud-pairs/s: the EMD half of _pairwise_EMD_CD_ (a next-row kernel, ~50x the CD work per
    # pair, reported separately under "emd") is switched off for the timed steps
    os.environ["PDGN_B200_SKIP_EMD"] = "1"
    from pdgn_b200 import dist as pdist
    from pdgn_b200 import evaluation_metrics as em
    from pdgn_b200 import ops

    nc = args.clouds
    smp_h = make_clouds(0, nc).pin_memory()
    ref_h = make_clouds(1, nc).pin_memory()
    smp_d, ref_d = smp_h.to(dev), ref_h.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_device():
        all_cd, _ = em._pairwise_EMD_CD_(smp_d, ref_d, 50)
        return all_cd

    out_h = torch.empty((nc, nc), dtype=torch.float32).pin_memory()

    def step_e2e():
        # pinned HOST tensors straight into the public API: it copies what this rank's tile needs (everything at N=1, its
        # rows + columns under torchrun) and returns the full matrix on the device; the copy into the pinned result buffer
        # is the D2H of the step's result
        all_cd, _ = em._pairwise_EMD_CD_(smp_h, ref_h, 50)
        out_h.copy_(all_cd, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out_h

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(step, k):
        """K steps bracketed by barrier + synchronize; CUDA events on the launch stream; max over ranks."""
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        marks = []
        for _ in range(k):
            flush.zero_()
            out = step()
            marks.append(torch.cuda.Event(enable_timing=True))
            marks[-1].record()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        per_step = [round(a.elapsed_time(b), 2) for a, b in zip([e0] + marks[:-1], marks)]
        return ms.item(), out, per_step

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        flush.zero_()  # warm-up steps are the timed steps, L2 flush included (its fill kernel is lazily loaded on first use)
        step_device()
    t_begin = time.time()
    ms_total, out, step_ms = timed(step_device, args.steps)
    clocks = sampler.stop((t_begin, time.time())) if rank == 0 else None
    for _ in range(min(args.warmup, 2)):
        flush.zero_()
        step_e2e()
    ms_e2e, out_e2e, step_ms_e2e = timed(step_e2e, args.steps)

    # the dominant kernel alone (rank's own tile, no collective): CUDA events around the C-ABI launch.  The two pack
    # kernels in the same call are ~0.01 % of it (profiles/: launch list).
    rows, cols = pdist.tile_of(rank, world, nc, nc)
    tile_pairs = (rows[1] - rows[0]) * (cols[1] - cols[0])
    sync_all()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kreps = max(1, min(args.steps, 3))
    k0.record()
    for _ in range(kreps):
        ops.cd_allpairs(smp_d, ref_d, rows=rows, cols=cols)
    k1.record()
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / kreps

    # ---- side measurements every rank takes part in (collectives inside)
    extras_all = {}
    if not args.no_extras:
        if world == 8 or os.environ.get("PDGN_BENCH_CFG5_CD"):
            extras_all["cd_4096x4096"] = bench_cfg5_cd(dev, em, dist, world, sync_all)
        extras_all["eval_emd"] = bench_eval_emd(dev, em, dist, world, sync_all, smp_h, ref_h, nc)
        os.environ["PDGN_B200_SKIP_EMD"] = "1"

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    pairs = float(nc) * nc
    value = pairs * args.steps / (ms_total * 1e-3)
    e2e_value = pairs * args.steps / (ms_e2e * 1e-3)
    props = torch.cuda.get_device_properties(dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    sm_max_mhz = float(peaks.get("sm_max_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0)
    lanes = props.multi_processor_count * 128
    peak_tflops = 2.0 * lanes * sm_max_mhz * 1e6 / 1e12           # FP32 FMA peak at max boost
    inst_rate = tile_pairs * POINT_PAIRS_PER_CLOUD_PAIR * FMA_PIPE_INSTR_PER_POINT_PAIR / (kernel_ms * 1e-3)
    achieved_tflops = 2.0 * inst_rate / 1e12                      # each FMA-pipe instruction = one FMA slot (2 FLOP)
    obs_mhz = (clocks or {}).get("sm_mhz") or sm_max_mhz
    roofline = {
        "bound": "fp32", "kernel": "cd_gram_kernel", "achieved": achieved_tflops, "peak": peak_tflops, "unit": "TFLOP/s",
        "frac": achieved_tflops / peak_tflops,
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch (bytes), or null outside the captured configuration
        "traffic": (traffic_from_profile(nc, world) or {}).get("bytes"), "traffic_detail": traffic_from_profile(nc, world),
        "kernel_ms": kernel_ms, "cloud_pairs_per_launch": tile_pairs,
        "definition": "achieved = cloud pairs x 2048^2 point pairs x 6 FMA-pipe instr x 2 FLOP-slots / kernel time (the ALGORITHMIC "
                      "work of SURVEY.md 8d: the direct-form distance); peak = SMs x 128 lanes x 2 x sm_max_mhz (MEASURED_PEAKS.json)",
        "executed": "the kernel computes |a|^2 + |b|^2 - 2a.b (the reference's own default arithmetic): 3 three-source FFMA + 1 FADD per pair "
                    "+ 1.0 min3, 5.4 SASS instr per pair; PDGN_B200_CD_EXACT=1 runs the 6-instruction direct form (bit-identical minima to "
                    "NmDistanceKernel) at ~0.70",
        "frac_of_executed_fma_pipe_instr": tile_pairs * POINT_PAIRS_PER_CLOUD_PAIR * 4 / (kernel_ms * 1e-3) / (lanes * sm_max_mhz * 1e6),
        "flop_frac": tile_pairs * POINT_PAIRS_PER_CLOUD_PAIR * FLOP_PER_POINT_PAIR / (kernel_ms * 1e-3) / 1e12 / peak_tflops,
        "frac_at_observed_clock": achieved_tflops / (2.0 * lanes * obs_mhz * 1e6 / 1e12),
        "sms": props.multi_processor_count,
    }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "step_ms": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD if nc == N_CLOUDS else "allpairs_cd_%dx%d_clouds_2048pts" % (nc, nc),
                   "clouds": [nc, nc], "points_per_cloud": N_PTS, "partition": "%dx%d rank grid, all_gather of scalars" % pdist.rank_grid(world),
                   "l2": "256 MiB memset between steps (inside the timed region, ~0.05 ms)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": world * ((rows[1] - rows[0]) + (cols[1] - cols[0])) * N_PTS * 3 * 4, "d2h_bytes_per_step": nc * nc * 4,
                "h2d_note": "summed over ranks: each rank copies only the rows + columns of its tile",
                "api": "pdgn_b200.evaluation_metrics._pairwise_EMD_CD_(pinned host tensors) (PDGN_B200_SKIP_EMD=1: CD half) + .cpu()"},
        # per step: 2 direct-form packs + 2 Gram-form packs + scale + gate + the direct-form kernel (returns at once when the gate
        # picks the Gram form) + the Gram-form kernel
        "gpu_launches": 8 * args.steps,
        "roofline": roofline, "clocks": clocks,
    }
    if not args.no_extras:
        line["knnquery"] = bench_knn(dev, lanes, sm_max_mhz)
        line["gathers"] = bench_gathers(dev, float(peaks.get("hbm_gbs") or 6650.0), "measured" if peaks.get("hbm_gbs") else "fallback")
        line["emd"] = bench_emd(dev)
        line["cfg5"] = dict(_guard(bench_cfg5_knn, dev, lanes, sm_max_mhz), **{k: v for k, v in extras_all.items() if k == "cd_4096x4096"})
        line["eval_emd"] = extras_all.get("eval_emd")
        line["cfg3"] = _guard(bench_cfg3, dev)
        if world == 1:
            line["reference_gpu"] = _guard(bench_reference_gpu, dev, value)
            n_s, n_r = 24, 50  # ~10 s of host work: a bounded sample of the 1000 x 1000 workload
            rate, dt = cpu_reference_rate(n_s, n_r)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": _CPU_KIND or "port", "seconds": dt,
                                    "sample": "%d x %d cloud pairs of 2048 points (reference Gram-form distChamfer loop, batch_size 50)" % (n_s, n_r)}
    # sanity: the e2e result equals the device-resident one
    line["e2e"]["matches_device_result"] = bool(torch.equal(out.cpu(), out_e2e))
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _time_ms(fn, reps, flush):
    import torch
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def bench_knn(dev, lanes, sm_max_mhz):
    """BASELINE config 2: knnquery k=20 on B=35 x 2048 xyz points (self query)."""
    import numpy as np
    import torch
    from pdgn_b200 import ops
    rng = np.random.default_rng(0)
    xyz = torch.from_numpy(rng.uniform(-1, 1, (35, 2048, 3)).astype(np.float32)).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ms = _time_ms(lambda: ops.knn_xyz(20, xyz), 10, flush)
    q = 35 * 2048
    inst = q * 2048 * FMA_PIPE_INSTR_PER_POINT_PAIR / (ms * 1e-3)
    return {"mqueries_per_s": q / (ms * 1e-3) / 1e6, "ms": ms, "shape": "B=35 n=m=2048 k=20",
            "fp32_issue_frac": inst / (lanes * sm_max_mhz * 1e6)}


def bench_emd(dev):
    """Next-row kernel: all-pairs approximate EMD on a 148 x 148 sample of the 1000 x 1000 workload."""
    import torch
    from pdgn_b200 import ops
    n = 148
    a, b = make_clouds(0, n).to(dev), make_clouds(1, n).to(dev)
    ops.emd_allpairs(a[:8], b[:8])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.emd_allpairs(a, b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    rate = n * n / (ms * 1e-3)
    return {"cloud_pairs_per_s": rate, "ms": ms, "sample": "148 x 148 cloud pairs of 2048 points", "seconds_per_1000x1000": 1e6 / rate}


def bench_gathers(dev, hbm_gbs, src):
    """Grouping fwd/bwd on the live shape (C=3) and the feature-sized stress shape; algorithmic bytes per SURVEY.md 8d."""
    import numpy as np
    import torch
    from pdgn_b200 import ops
    rng = np.random.default_rng(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = {"hbm_peak_gbs": hbm_gbs, "hbm_peak_source": src}
    for tag, (b, c, n, m, k) in {"c3_live": (35, 3, 2048, 2048, 20), "c256_stress": (35, 256, 1024, 1024, 10)}.items():
        feat = torch.from_numpy(rng.standard_normal((b, c, n)).astype(np.float32)).to(dev)
        idx = torch.from_numpy(rng.integers(0, n, (b, m, k)).astype(np.int32)).to(dev)
        go = torch.from_numpy(rng.standard_normal((b, c, m, k)).astype(np.float32)).to(dev)
        out = torch.empty((b, c, m, k), dtype=torch.float32, device=dev)
        grad = torch.zeros((b, c, n), dtype=torch.float32, device=dev)
        from pdgn_b200._lib import lib
        L = lib()
        st = torch.cuda.current_stream().cuda_stream
        f_ms = _time_ms(lambda: L.pdgn_group_fwd(feat.data_ptr(), idx.data_ptr(), b, c, n, m, k, out.data_ptr(), st), 10, flush)
        ws_bytes = L.pdgn_group_bwd_workspace(b, n, m, k)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        b_ms = _time_ms(lambda: L.pdgn_group_bwd_ws(go.data_ptr(), idx.data_ptr(), b, c, n, m, k, grad.data_ptr(), ws.data_ptr(), ws_bytes, st), 10, flush)
        bytes_fwd = 4.0 * (b * c * m * k + b * m * k + b * c * n)
        res[tag] = {"fwd_ms": f_ms, "fwd_gbs": bytes_fwd / (f_ms * 1e-3) / 1e9, "fwd_frac": bytes_fwd / (f_ms * 1e-3) / 1e9 / hbm_gbs,
                    "bwd_ms": b_ms, "bwd_gbs": bytes_fwd / (b_ms * 1e-3) / 1e9, "bwd_frac": bytes_fwd / (b_ms * 1e-3) / 1e9 / hbm_gbs,
                    "algorithmic_mb": bytes_fwd / 1e6}
    # three-point interpolation, 1024 -> 2048 points, C=256; bytes = 4(BCn + BCm) + 24 Bn
    b, c, m, n = 35, 256, 1024, 2048
    feat = torch.from_numpy(rng.standard_normal((b, c, m)).astype(np.float32)).to(dev)
    idx3 = torch.from_numpy(rng.integers(0, m, (b, n, 3)).astype(np.int32)).to(dev)
    w3 = torch.rand((b, n, 3), device=dev)
    out = torch.empty((b, c, n), dtype=torch.float32, device=dev)
    grad = torch.zeros((b, c, m), dtype=torch.float32, device=dev)
    ws_bytes = L.pdgn_interp_bwd_workspace(b, n, m)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    f_ms = _time_ms(lambda: L.pdgn_interp_fwd(feat.data_ptr(), idx3.data_ptr(), w3.data_ptr(), b, c, m, n, out.data_ptr(), st), 10, flush)
    b_ms = _time_ms(lambda: L.pdgn_interp_bwd_ws(out.data_ptr(), idx3.data_ptr(), w3.data_ptr(), b, c, n, m, grad.data_ptr(), ws.data_ptr(), ws_bytes, st), 10, flush)
    nbytes = 4.0 * (b * c * n + b * c * m) + 24.0 * b * n
    res["interp_c256"] = {"fwd_ms": f_ms, "fwd_gbs": nbytes / (f_ms * 1e-3) / 1e9, "fwd_frac": nbytes / (f_ms * 1e-3) / 1e9 / hbm_gbs,
                          "bwd_ms": b_ms, "bwd_gbs": nbytes / (b_ms * 1e-3) / 1e9, "bwd_frac": nbytes / (b_ms * 1e-3) / 1e9 / hbm_gbs,
                          "algorithmic_mb": nbytes / 1e6}
    # edge features of the generator's last stage (get_edge_features, PDGNet_v2.py:461-477): C=256, N=1024, k=num_k//2=10.
    # bytes = the [B,2C,N,k] tensor + x / grad_x + the int64 index
    b, c, n, k = 35, 256, 1024, 10
    x = torch.from_numpy(rng.standard_normal((b, c, n)).astype(np.float32)).to(dev)
    idx64 = torch.from_numpy(rng.integers(0, n, (b, n, k)).astype(np.int64)).to(dev)
    ee = torch.empty((b, 2 * c, n, k), dtype=torch.float32, device=dev)
    gx = torch.zeros((b, c, n), dtype=torch.float32, device=dev)
    ws_bytes = L.pdgn_edge_feat_bwd_workspace(b, n, k)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    f_ms = _time_ms(lambda: L.pdgn_edge_feat_fwd(x.data_ptr(), idx64.data_ptr(), b, c, n, k, ee.data_ptr(), st), 10, flush)
    b_ms = _time_ms(lambda: L.pdgn_edge_feat_bwd_ws(ee.data_ptr(), idx64.data_ptr(), b, c, n, k, gx.data_ptr(), ws.data_ptr(), ws_bytes, st), 10, flush)
    nbytes = 4.0 * (2 * b * c * n * k + b * c * n) + 8.0 * b * n * k
    res["edge_c256"] = {"fwd_ms": f_ms, "fwd_gbs": nbytes / (f_ms * 1e-3) / 1e9, "fwd_frac": nbytes / (f_ms * 1e-3) / 1e9 / hbm_gbs,
                        "bwd_ms": b_ms, "bwd_gbs": nbytes / (b_ms * 1e-3) / 1e9, "bwd_frac": nbytes / (b_ms * 1e-3) / 1e9 / hbm_gbs,
                        "algorithmic_mb": nbytes / 1e6}
    return res


def _guard(fn, *a):
    """A side measurement must never take the headline line down with it."""
    try:
        return fn(*a)
    except Exception as e:  # noqa: BLE001
        return {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}


def bench_cfg5_knn(dev, lanes, sm_max_mhz):
    """BASELINE configs[4], kNN half: knnquery k=32, self query, 16384-point clouds, B=8 and B=35."""
    import numpy as np
    import torch
    from pdgn_b200 import ops
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = {}
    for b in (8, 35):
        rng = np.random.default_rng(b)
        xyz = torch.from_numpy(rng.uniform(-1, 1, (b, 16384, 3)).astype(np.float32)).to(dev)
        ms = _time_ms(lambda: ops.knn_xyz(32, xyz), 5, flush)
        q = b * 16384
        res["knn_k32_B%d_n16384" % b] = {"mqueries_per_s": q / (ms * 1e-3) / 1e6, "ms": ms,
                                        "fp32_issue_frac": q * 16384.0 * FMA_PIPE_INSTR_PER_POINT_PAIR / (ms * 1e-3) / (lanes * sm_max_mhz * 1e6)}
    return res


def bench_cfg5_cd(dev, em, dist, world, sync_all):
    """BASELINE configs[4], CD half: one 4096 x 4096 all-pairs matrix of 2048-point clouds over all ranks (device-resident
    inputs, all_gather of the scalars inside; 1 warm-up + 2 timed steps; max over ranks)."""
    import torch
    try:
        n = int(os.environ.get("PDGN_BENCH_CFG5_CLOUDS", "4096"))
        a, b = make_clouds(2, n).to(dev), make_clouds(3, n).to(dev)
        em._pairwise_EMD_CD_(a, b, 50)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            out, _ = em._pairwise_EMD_CD_(a, b, 50)
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1) / 2], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = ms.item()
        return {"clouds": [n, n], "ms": ms, "cloud_pairs_per_s": float(n) * n / (ms * 1e-3), "n_gpus": world,
                "fp32_issue_frac_per_gpu": float(n) * n * POINT_PAIRS_PER_CLOUD_PAIR * FMA_PIPE_INSTR_PER_POINT_PAIR / (ms * 1e-3) /
                (world * torch.cuda.get_device_properties(dev).multi_processor_count * 128 * 1965.0e6),
                "checksum": float(out.double().sum().item())}
    except Exception as e:  # noqa: BLE001
        return {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}


def bench_eval_emd(dev, em, dist, world, sync_all, smp_h, ref_h, nc):
    """What PDGNet_v2.test() actually waits for (PDGNet_v2.py:319): compute_all_metrics = three all-pairs CD matrices AND three
    all-pairs approximate-EMD matrices, here from pinned HOST tensors with the H2D copies and a D2H of the 12 scalars inside
    the timed region (one run, no repeat: it is seconds to a minute long)."""
    import torch
    try:
        os.environ["PDGN_B200_SKIP_EMD"] = "0"
        n = int(os.environ.get("PDGN_BENCH_EVAL_CLOUDS", str(nc)))
        a, b = smp_h[:n], ref_h[:n]
        em.compute_all_metrics(a[:16], b[:16], 50)      # warm-up: lazy module loads, NCCL channels
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = em.compute_all_metrics(a, b, 50)
        host = {k: float(v.item()) for k, v in res.items()}
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        sec = ms.item() * 1e-3
        return {"clouds": [n, n], "seconds": sec, "cloud_pairs_per_s_cd_and_emd": 3.0 * n * n / sec, "n_gpus": world, "keys": len(host),
                "h2d_bytes": 2 * n * N_PTS * 3 * 4, "reference_says": "'may take about 2 hours' for the test phase (README.md:47)",
                "mmd_cd": host.get("lgan_mmd-CD"), "mmd_emd": host.get("lgan_mmd-EMD")}
    except Exception as e:  # noqa: BLE001
        return {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
    finally:
        os.environ["PDGN_B200_SKIP_EMD"] = "1"


def _ev_ms(fn, reps=5, warm=2):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def bench_cfg3(dev):
    """BASELINE configs[2] (training path, B=35): the loss side of one G step -- six get_local_pair calls = 12 kNN + 12 grouping +
    12 ChamferLoss problems (PDGNet_v2.py:232-237) -- forward+backward through the op the drop-in installs, the generator's four
    feature-space kNN stages (get_edge_features_xyz, k=10) forward+backward, and, when the reference tree is staged, one G step
    of the reference's OWN PointGenerator + get_local_pair code run through the drop-in vs over the reference stack."""
    import numpy as np
    import torch
    from pdgn_b200 import edge_features as ef
    from pdgn_b200 import local_pair
    B = 35
    rng = np.random.default_rng(0)

    def cloud(n):
        v = rng.standard_normal((B, n, 3))
        v /= np.linalg.norm(v, axis=-1, keepdims=True)
        return torch.from_numpy(np.ascontiguousarray((0.5 * v).astype(np.float32).transpose(0, 2, 1))).to(dev)

    pts = {n: cloud(n) for n in (256, 512, 1024, 2048)}
    sizes = [(256, 512), (256, 1024), (256, 2048), (512, 1024), (512, 2048), (1024, 2048)]

    def loss_side(backward):
        leaves = {n: p.clone().requires_grad_(backward) for n, p in pts.items()}
        total = 0
        for m_, n_ in sizes:
            a, b = local_pair.get_local_pair(leaves[m_], leaves[n_])
            total = total + a + b
        if backward:
            total.backward()
        return total

    def loss_side_batched(backward):
        leaves = [pts[n].clone().requires_grad_(backward) for n in (256, 512, 1024, 2048)]
        total = local_pair.shape_losses(leaves, 20).sum()
        if backward:
            total.backward()
        return total

    res = {"batch": B, "loss_side_fwd_ms": _ev_ms(lambda: loss_side_batched(False)), "loss_side_fwd_bwd_ms": _ev_ms(lambda: loss_side_batched(True)),
           "loss_side_launches": {"fwd": 6, "bwd": 5, "note": "pdgn_shape_loss_fwd/bwd: all 9 kNN / 9 statistics / 24 minimum problems of the "
                                  "step per operator launch; what the drop-in runs for the trainer's six get_local_pair calls"},
           "loss_side_per_call_fwd_ms": _ev_ms(lambda: loss_side(False)), "loss_side_per_call_fwd_bwd_ms": _ev_ms(lambda: loss_side(True))}
    stages, feat_knn = {}, {}
    for c, n in [(32, 128), (64, 256), (128, 512), (256, 1024)]:
        x = torch.randn(B, c, n, device=dev)
        pc = torch.rand(B, 3, n, device=dev) * 2 - 1

        def stage():
            xr, pr = x.clone().requires_grad_(True), pc.clone().requires_grad_(True)
            e_fea, e_xyz = ef.get_edge_features_xyz(xr, pr, 10)
            (e_fea.sum() + e_xyz.sum()).backward()

        stages["C%d_N%d" % (c, n)] = _ev_ms(stage)
        # the kNN kernel of the stage alone: exact FP32 direct distances, 2C FMA-pipe instructions per (query, candidate) pair
        from pdgn_b200 import ops
        kms = _ev_ms(lambda: ops.knn_feat(x, 10, skip=1))
        # the same op on the FP32 SIMT kernel (csrc/knn_feat.cu), which the tensor-core path (csrc/knn_feat_tc.cu: tcgen05 TF32
        # Gram filter + exact FP32 re-rank, identical indices) replaced for these shapes
        from pdgn_b200._lib import lib as _lib_
        _L = _lib_()
        _idx = torch.empty((B, n, 10), dtype=torch.int64, device=dev)
        _st = torch.cuda.current_stream().cuda_stream
        sms = _ev_ms(lambda: _L.pdgn_knn_feat(x.data_ptr(), B, c, n, 10, 1, _idx.data_ptr(), None, _st))
        feat_knn["C%d_N%d" % (c, n)] = {"ms": kms, "impl": "tcgen05 tf32 filter + exact fp32 re-rank", "simt_kernel_ms": sms,
                                        "speedup_over_simt": sms / kms,
                                        "fp32_issue_frac": B * n * n * 2.0 * c / (kms * 1e-3) / (148 * 128 * 1.965e9),
                                        "fp32_issue_frac_note": "direct-form work (2 FP32 instr per pair and channel) / time / FP32 issue peak; > 1 is possible: the pairwise work runs on the tensor cores"}
    res["edge_feature_stage_fwd_bwd_ms"] = stages
    res["feature_knn"] = feat_knn
    try:
        from oracle import ref_tree
        if not ref_tree.available():
            raise RuntimeError("reference tree not staged")
        ref = ref_tree.load_reference()
        drop = ref_tree.load_dropin()
        torch.manual_seed(0)
        g_ref = ref.model.PointGenerator(2048, 20).to(dev).train()
        g_drop = drop.PointGenerator(2048, 20).to(dev).train()
        g_drop.load_state_dict(g_ref.state_dict())
        z = torch.randn(B, 128, device=dev) * 0.2
        t_ref = ref_tree.bare_trainer(ref.model, ref.pointops.Gen_QueryAndGroupXYZ(radius=None, nsample=20, use_xyz=False), ref.chamfer_loss.ChamferLoss())
        t_drop = ref_tree.bare_trainer(drop, drop.pointops.Gen_QueryAndGroupXYZ(radius=None, nsample=20, use_xyz=False), drop.chamfer_loss.ChamferLoss())

        def g_step(gen, trainer):
            gen.zero_grad(set_to_none=True)
            p1, p2, p3, p4 = gen(z)
            sim = 0.0
            for a, b in ((p1, p2), (p1, p3), (p1, p4), (p2, p3), (p2, p4), (p3, p4)):
                mu, cov = trainer.get_local_pair(a, b)
                sim = sim + mu + cov
            (0.1 * sim).backward()

        res["generator_step_ms"] = _ev_ms(lambda: g_step(g_drop, t_drop), reps=3, warm=2)
        res["generator_step_reference_stack_ms"] = _ev_ms(lambda: g_step(g_ref, t_ref), reps=2, warm=1)
        res["generator_step_note"] = ("the reference's own PointGenerator.forward + get_local_pair x6 + backward (PDGNet_v2.py:228-255 without the "
                                      "discriminators), batch 35, same weights: through pdgn_b200.dropin vs over the reference's torch "
                                      "get_edge_features + recompiled pointops kernels + torch Gram ChamferLoss on this GPU")
    except Exception as e:  # noqa: BLE001
        res["generator_step_ms"] = None
        res["generator_step_note"] = "not measured: %s" % str(e)[:200]
    return res


def bench_reference_gpu(dev, our_pairs_per_s):
    """'Reference on B200' (SURVEY.md 8d): the reference's two ways of filling the CD matrix, on THIS GPU, on a bounded sample
    of the 1000 x 1000 workload (10 generated clouds x all 1000 references, batch_size 50 as in README's test):
      torch_gram   -- its default path: torch bmm Gram-form distChamfer inside the Python double loop (evaluation_metrics.py:85-121)
      nndistance   -- accelerated_cd=True: the NNDistance kernel (nndistance.cu:2-128) recompiled for sm_100a, same loop
    and the reference's real _pairwise_EMD_CD_ (CD + EMD) on 2 x 100 pairs next to ours with EMD on."""
    import torch
    from oracle import ref_kernels, torch_ref
    if not ref_kernels.available():
        return {"unavailable": "oracle/_ref/libpdgn_ref.so not built"}
    n_s, n_r = 10, N_CLOUDS
    smp, ref = make_clouds(0, n_s).to(dev), make_clouds(1, n_r).to(dev)
    res = {"sample": "%d x %d cloud pairs of 2048 points, batch_size 50" % (n_s, n_r)}

    def wall(fn, reps):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    dist_fn = torch_ref.dist_chamfer
    kind = "port"
    try:
        from oracle import ref_tree
        if ref_tree.available():
            rmod = ref_tree.load_reference().evaluation_metrics
            dist_fn, kind = rmod.distChamfer, "reference"
    except Exception:  # noqa: BLE001
        rmod = None

    def loop(dfn):
        rows = []
        for s in range(n_s):
            parts = []
            for r0 in range(0, n_r, 50):
                rb = ref[r0:r0 + 50]
                rep = smp[s].view(1, -1, 3).expand(rb.size(0), -1, -1).contiguous()
                dl, dr = dfn(rep, rb)
                parts.append((dl.mean(dim=1) + dr.mean(dim=1)).view(1, -1))
            rows.append(torch.cat(parts, dim=1))
        return torch.cat(rows, dim=0)

    def nnd(a, b):
        d1, _, d2, _ = ref_kernels.nndistance(a, b)
        return d1, d2

    t = wall(lambda: loop(dist_fn), 2)
    res["torch_gram"] = {"cloud_pairs_per_s": n_s * n_r / t, "seconds": t, "kind": kind, "ours_over_it": our_pairs_per_s / (n_s * n_r / t)}
    t = wall(lambda: loop(nnd), 2)
    res["nndistance"] = {"cloud_pairs_per_s": n_s * n_r / t, "seconds": t, "kind": "reference", "ours_over_it": our_pairs_per_s / (n_s * n_r / t)}
    if kind == "reference":
        from pdgn_b200 import ops
        a, b = smp[:2].contiguous(), ref[:100].contiguous()
        t = wall(lambda: rmod._pairwise_EMD_CD_(a, b, 50, accelerated_cd=False), 1)
        ours = make_clouds(0, 148).to(dev), make_clouds(1, 148).to(dev)
        t_o = wall(lambda: (ops.cd_allpairs(*ours), ops.emd_allpairs(*ours)), 1)
        res["pairwise_emd_cd"] = {"reference_cloud_pairs_per_s": 200 / t, "reference_sample": "2 x 100 pairs (the function PDGNet_v2.test() "
                                  "reaches through compute_all_metrics, CD + EMD)", "ours_cloud_pairs_per_s": 148 * 148 / t_o,
                                  "ours_sample": "148 x 148 pairs, CD + EMD", "ours_over_it": (148 * 148 / t_o) / (200 / t)}
    return res


if __name__ == "__main__":
    main()
