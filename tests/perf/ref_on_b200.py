"""'Reference on B200' column: time the reference's own CUDA kernels, recompiled unmodified for
sm_100a (oracle/_ref), next to the pdgn_b200 kernels on the same shapes.  Lives under tests/ because it executes oracle/_ref (test infrastructure); not collected by pytest.
Usage (GPU box): python tests/perf/ref_on_b200.py > gpurun_out/ref_on_b200.txt"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref_kernels as rk  # noqa: E402
from pdgn_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3  # ms (the reference wrappers synchronise: NULL-stream kernels)


def row(name, ref_ms, our_ms, note=""):
    print("%-46s reference %10.3f ms   pdgn_b200 %9.4f ms   speed-up %8.1fx  %s" % (name, ref_ms, our_ms, ref_ms / our_ms, note))


xyz = torch.from_numpy(rng.uniform(-1, 1, (35, 2048, 3)).astype(np.float32)).to(dev)
row("knnquery k=20, B=35, n=m=2048", timeit(lambda: rk.knnquery(20, xyz, xyz), 3), timeit(lambda: ops.knn_xyz(20, xyz)))
unk = torch.from_numpy(rng.uniform(-1, 1, (35, 2048, 3)).astype(np.float32)).to(dev)
kn = torch.from_numpy(rng.uniform(-1, 1, (35, 1024, 3)).astype(np.float32)).to(dev)
row("nearestneighbor (3-NN) 2048 vs 1024, B=35", timeit(lambda: rk.nn3(unk, kn), 3), timeit(lambda: ops.nn3(unk, kn)))
for tag, (b, c, n, m, k) in {"C=3 live": (35, 3, 2048, 2048, 20), "C=256 stress": (35, 256, 1024, 1024, 10)}.items():
    feat = torch.randn(b, c, n, device=dev)
    idx = torch.randint(0, n, (b, m, k), device=dev, dtype=torch.int32)
    go = torch.randn(b, c, m, k, device=dev)
    row("grouping fwd %s" % tag, timeit(lambda: rk.group_fwd(feat, idx)), timeit(lambda: ops.group_fwd(feat, idx)))
    row("grouping bwd %s" % tag, timeit(lambda: rk.group_bwd(go, idx, n)), timeit(lambda: ops.group_bwd(go, idx, n)))
feat = torch.randn(35, 256, 1024, device=dev)
idx3 = torch.randint(0, 1024, (35, 2048, 3), device=dev, dtype=torch.int32)
w3 = torch.rand(35, 2048, 3, device=dev)
go = torch.randn(35, 256, 2048, device=dev)
row("interpolation fwd C=256 1024->2048", timeit(lambda: rk.interp_fwd(feat, idx3, w3)), timeit(lambda: ops.interp_fwd(feat, idx3, w3)))
row("interpolation bwd C=256 1024->2048", timeit(lambda: rk.interp_bwd(go, idx3, w3, 1024)), timeit(lambda: ops.interp_bwd(go, idx3, w3, 1024)))
a = torch.rand(50, 2048, 3, device=dev)
b_ = torch.rand(50, 2048, 3, device=dev)
row("NNDistance (both directions) 50 x 2048^2", timeit(lambda: rk.nndistance(a, b_)), timeit(lambda: ops.chamfer_min(a, b_)))
# all-pairs CD through the reference's accelerated_cd loop shape: 1 sample cloud expanded against 50 refs per call
smp = torch.rand(20, 2048, 3, device=dev)


def ref_pairwise():
    for s in range(smp.size(0)):
        rep = smp[s].view(1, -1, 3).expand(50, -1, -1).contiguous()
        d1, _, d2, _ = rk.nndistance(rep, b_)
        (d1.mean(1) + d2.mean(1))


row("all-pairs CD 20 x 50 clouds (NNDistance loop)", timeit(ref_pairwise, 2), timeit(lambda: ops.cd_allpairs(smp, b_)),
    "(reference default path is slower still: torch bmm Gram form)")
rep = smp[0].view(1, -1, 3).expand(16, -1, -1).contiguous()
row("approx EMD 16 pairs x 2048^2 (ApproxMatch+MatchCost)", timeit(lambda: rk.match_cost(rep, b_[:16].contiguous()), 2),
    timeit(lambda: ops.emd_allpairs(smp[:1], b_[:16].contiguous())), "(one CTA per pair: our kernel is built for >= 300 pairs)")
