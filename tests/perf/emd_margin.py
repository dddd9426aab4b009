"""Max relative error of the all-pairs EMD kernel against the CPU oracle and the recompiled reference kernels on a few
shapes (tests/perf: executes oracle/).  PDGN_LIB=<other .so> measures another build."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pdgn_b200 import _lib, ops
if os.environ.get("PDGN_LIB"):
    _lib.SO_PATH = os.environ["PDGN_LIB"]
from oracle import cpu as ocpu
from oracle import ref_kernels as rk
rng = np.random.default_rng(11)
def sph(k, n):
    v = rng.standard_normal((k, n, 3)); v /= np.linalg.norm(v, axis=-1, keepdims=True); return v.astype(np.float32)
def uni(k, n):
    return rng.uniform(-1, 1, (k, n, 3)).astype(np.float32)
def blob(k, n):
    return (rng.standard_normal((k, n, 3)) * rng.uniform(0.05, 0.6, (k, 1, 3))).astype(np.float32)
worst_o = worst_r = 0.0
for maker in (sph, uni, blob):
    for (na, nb, n) in [(4, 4, 2048), (4, 4, 1024), (4, 4, 512)]:
        A, B = maker(na, n), maker(nb, n)
        out = ops.emd_allpairs(torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()).cpu().numpy()
        ref = (ocpu.emd_cost(np.repeat(A, nb, axis=0), np.tile(B, (na, 1, 1))) / np.float32(n)).reshape(na, nb)
        eo = float(np.max(np.abs(out - ref) / np.abs(ref)))
        er = 0.0
        for s_ in range(na):
            rep = torch.from_numpy(A[s_:s_ + 1]).cuda().expand(nb, -1, -1).contiguous()
            r = (rk.match_cost(rep, torch.from_numpy(B).cuda()) / float(n)).cpu().numpy()
            er = max(er, float(np.max(np.abs(out[s_] - r) / np.abs(r))))
        worst_o, worst_r = max(worst_o, eo), max(worst_r, er)
        print("%-5s n=%4d  vs oracle %.2e   vs reference kernels %.2e" % (maker.__name__, n, eo, er), flush=True)
print("worst: vs oracle %.2e, vs reference kernels %.2e (test tolerance 2e-4)" % (worst_o, worst_r))
