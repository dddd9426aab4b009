"""All-pairs CD: accuracy of the Gram-form kernel against the FP64 truth / the direct-form oracle, and timings of both forms
(PDGN_B200_CD_EXACT=1 selects the direct form).  tools/ only.  Usage: python tools/cd_check.py [clouds]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import clouds_sphere, clouds_uniform
from oracle import cpu as ocpu
from pdgn_b200 import ops
dev = torch.device("cuda:0")
mode = "exact" if os.environ.get("PDGN_B200_CD_EXACT") == "1" else "default"
rng = np.random.default_rng(0)

def truth64(A, B):
    out = np.zeros((len(A), len(B)))
    for i, a in enumerate(A.astype(np.float64)):
        for j, b in enumerate(B.astype(np.float64)):
            d = ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1)
            out[i, j] = d.min(1).mean() + d.min(0).mean()
    return out

def plane(rng, n, npts, _):
    v = rng.uniform(-1, 1, (n, npts, 3)); v[..., 2] *= 1e-3
    return v.astype(np.float32)

for name, maker, npts in [("sphere", clouds_sphere, 2048), ("uniform cube", clouds_uniform, 2048), ("thin plane", plane, 2048),
                          ("sphere 512", clouds_sphere, 512), ("half-scale sphere", lambda r, n, p, d: 0.5 * clouds_sphere(r, n, p, d), 2048)]:
    A, B = maker(rng, 6, npts, 3), maker(rng, 5, npts, 3)
    t = truth64(A, B)
    ours = ops.cd_allpairs(torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)).cpu().numpy().astype(np.float64)
    orc = ocpu.cd_allpairs(A, B).astype(np.float64)
    print("%-18s mode=%-7s max rel err vs FP64 truth: ours %.2e   direct-form oracle %.2e   ours vs oracle %.2e" % (
        name, mode, np.abs(ours / t - 1).max(), np.abs(orc / t - 1).max(), np.abs(ours / orc - 1).max()), flush=True)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 592
A = torch.from_numpy(clouds_sphere(rng, n, 2048, 3)).to(dev)
B = torch.from_numpy(clouds_sphere(rng, n, 2048, 3)).to(dev)
ops.cd_allpairs(A, B); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.cd_allpairs(A, B); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("mode=%s %dx%d clouds of 2048: %.2f ms = %.4g cloud-pairs/s = %.3f of the 6-instr issue roofline" % (
    mode, n, n, ms, n * n / ms * 1e3, n * n * 2048.0 * 2048 * 6 / (ms * 1e-3) / (148 * 128 * 1.965e9)))
e0.record(); ops.cd_allpairs(A, A); e1.record(); torch.cuda.synchronize()
print("mode=%s same-set (symmetric) %dx%d: %.2f ms" % (mode, n, n, e0.elapsed_time(e1)))
