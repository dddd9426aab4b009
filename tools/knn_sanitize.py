"""Sliced / unsliced kNN select kernels on small shapes for compute-sanitizer (tools/ only)."""
import sys
import torch
sys.path.insert(0, ".")
from pdgn_b200 import ops
g = torch.Generator().manual_seed(1)
dev = torch.device("cuda:0")
R = lambda *s: (torch.rand(*s, generator=g) * 2 - 1).to(dev)
for (b, n, m, k) in [(3, 256, 256, 20), (2, 2048, 256, 20), (2, 1001, 130, 20), (4, 501, 600, 16), (2, 2039, 1100, 13), (1, 300, 40, 3), (2, 70, 50, 20), (40, 2048, 128, 20), (70, 512, 300, 8)]:
    ops.knn_xyz(k, R(b, n, 3), R(b, m, 3), return_dist=True)
ops.nn3(R(3, 700, 3), R(3, 333, 3))
torch.cuda.synchronize()
print("knn sanitize target done")
