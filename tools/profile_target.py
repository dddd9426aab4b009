"""Small driver for ncu captures (tools/; not part of the product): runs each hot kernel a few times on its
BASELINE shape so that `ncu -k regex:<kernel>` can pick it up.  Usage: python tools/profile_target.py [what ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdgn_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
what = sys.argv[1:] or ["cd", "knn", "group", "feat", "chamfer", "nn3", "interp"]


def sphere(n, npts):
    v = rng.standard_normal((n, npts, 3))
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    return torch.from_numpy(v.astype(np.float32)).to(dev)


reps = int(os.environ.get("REPS", "3"))
for _ in range(reps):
    if "cd" in what:
        n = int(os.environ.get("CD_CLOUDS", "296"))
        A, B = sphere(n, 2048), sphere(n, 2048)
        ops.cd_allpairs(A, B)
    if "knn" in what:
        xyz = torch.from_numpy(rng.uniform(-1, 1, (35, 2048, 3)).astype(np.float32)).to(dev)
        ops.knn_xyz(20, xyz)
    if "nn3" in what:
        unk = torch.from_numpy(rng.uniform(-1, 1, (35, 2048, 3)).astype(np.float32)).to(dev)
        kn = torch.from_numpy(rng.uniform(-1, 1, (35, 1024, 3)).astype(np.float32)).to(dev)
        ops.nn3(unk, kn)
    if "group" in what:
        for (b, c, n, m, k) in [(35, 3, 2048, 2048, 20), (35, 256, 1024, 1024, 10)]:
            feat = torch.randn(b, c, n, device=dev)
            idx = torch.randint(0, n, (b, m, k), device=dev, dtype=torch.int32)
            out = ops.group_fwd(feat, idx)
            ops.group_bwd(out, idx, n)
    if "interp" in what:
        feat = torch.randn(35, 256, 1024, device=dev)
        idx = torch.randint(0, 1024, (35, 2048, 3), device=dev, dtype=torch.int32)
        w = torch.rand(35, 2048, 3, device=dev)
        out = ops.interp_fwd(feat, idx, w)
        ops.interp_bwd(out, idx, w, 1024)
    if "feat" in what:
        x = torch.randn(35, 256, 1024, device=dev)
        idx = ops.knn_feat(x, 10)
        ee = ops.edge_feat_fwd(x, idx)
        ops.edge_feat_bwd(ee, idx, 256)
    if "chamfer" in what:
        a = torch.rand(35, 1024, 3, device=dev)
        b_ = torch.rand(35, 1024, 3, device=dev)
        ops.chamfer_min(a, b_)
torch.cuda.synchronize()
print("done")
