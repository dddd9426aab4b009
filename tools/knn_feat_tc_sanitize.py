"""Small tensor-core feature-kNN problems for compute-sanitizer (tools/ only)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import cpu as ocpu
from pdgn_b200 import ops
rng = np.random.default_rng(3)
for (b, c, n, k) in [(2, 32, 128, 10), (1, 64, 256, 10), (1, 40, 384, 19)]:
    x = rng.standard_normal((b, c, n)).astype(np.float32)
    idx, d2 = ops.knn_feat(torch.from_numpy(x).cuda(), k, skip=1, return_dist=True)
    ri, rd = ocpu.knn_feat(x, k, skip=1)
    assert np.array_equal(idx.cpu().numpy(), ri) and np.array_equal(d2.cpu().numpy(), rd)
print("sanitize run ok")
