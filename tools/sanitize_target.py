"""Small-shape pass over every kernel for compute-sanitizer (memcheck / racecheck / initcheck); tools/ only.
Usage: compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdgn_b200 import edge_features as ef  # noqa: E402
from pdgn_b200 import evaluation_metrics as em  # noqa: E402
from pdgn_b200 import local_pair, ops, pointops  # noqa: E402
from pdgn_b200.chamfer_loss import ChamferLoss  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(0)
R = lambda *s: (torch.rand(*s, generator=g) * 2 - 1).to(dev)

for n, m, k in [(300, 77, 20), (2048, 130, 20), (2500, 64, 20), (700, 40, 32), (50, 9, 8), (333, 33, 3), (4100, 33, 50)]:
    ops.knn_xyz(k, R(2, n, 3), R(2, m, 3), return_dist=True)
ops.nn3(R(2, 100, 3), R(2, 37, 3))
for b, c, n, m, k in [(2, 3, 64, 50, 5), (2, 16, 256, 128, 8), (1, 8, 1500, 300, 4), (2, 5, 33, 7, 3)]:
    f = R(b, c, n).requires_grad_(True)
    idx = torch.randint(0, n, (b, m, k), generator=g, dtype=torch.int32).to(dev)
    pointops.grouping(f, idx).sum().backward()
    i3 = torch.randint(0, n, (b, m, 3), generator=g, dtype=torch.int32).to(dev)
    w = torch.rand(b, m, 3, generator=g).to(dev)
    f2 = R(b, c, n).requires_grad_(True)
    pointops.interpolation(f2, i3, w).sum().backward()
for d in (3, 9):
    p, q = R(2, 130, d).requires_grad_(True), R(2, 70, d).requires_grad_(True)
    ChamferLoss()(p, q).backward()
A, B = R(5, 300, 3), R(4, 300, 3)
ops.cd_allpairs(A, B)
ops.cd_allpairs(A, A)
ops.cd_allpairs(R(2, 2500, 3), R(3, 2500, 3))
ops.cd_allpairs_host(A.cpu(), B.cpu())
ops.emd_allpairs(A, B)
ops.emd_allpairs(R(2, 100, 3), R(2, 260, 3))
x = torch.randn(2, 24, 200, generator=g).to(dev).requires_grad_(True)
pc = R(2, 3, 200).requires_grad_(True)
e1, e2 = ef.get_edge_features_xyz(x, pc, 10)
(e1.sum() + e2.sum()).backward()
p1, p2 = R(2, 3, 128).requires_grad_(True), R(2, 3, 300).requires_grad_(True)
a, b_ = local_pair.get_local_pair(p1, p2)
(a + b_).backward()
em.jsd_between_point_cloud_sets(R(3, 256, 3) * 0.5, R(3, 256, 3) * 0.3)
torch.cuda.synchronize()
print("sanitize target done")
