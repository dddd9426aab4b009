"""One edge-feature backward and one grouping backward on the C=256 shape for ncu (tools/ only)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from pdgn_b200 import ops
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
b, c, n, k = 35, 256, 1024, 10
idx = torch.from_numpy(rng.integers(0, n, (b, n, k)).astype(np.int32)).to(dev)
gee = torch.randn(b, 2 * c, n, k, device=dev)
go = torch.randn(b, c, n, k, device=dev)
for _ in range(2):
    ops.edge_feat_bwd(gee, idx.long(), c)
    ops.group_bwd(go, idx, n)
torch.cuda.synchronize()
w3 = torch.rand(b, 2 * n, 3, device=dev)
idx3 = torch.from_numpy(rng.integers(0, n, (b, 2 * n, 3)).astype(np.int32)).to(dev)
go3 = torch.randn(b, c, 2 * n, device=dev)
for _ in range(2):
    ops.interp_bwd(go3, idx3, w3, n)
torch.cuda.synchronize()
