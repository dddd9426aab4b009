"""ncu / timing target: the batched shape-preserving loss of one G step (B=35, levels 256/512/1024/2048).  tools/ only."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from pdgn_b200 import local_pair
rng = np.random.default_rng(0)
def cloud(n):
    v = rng.standard_normal((35, n, 3)); v /= np.linalg.norm(v, axis=-1, keepdims=True)
    return torch.from_numpy(np.ascontiguousarray((0.5 * v).astype(np.float32).transpose(0, 2, 1))).cuda()
pts = [cloud(n) for n in (256, 512, 1024, 2048)]
for _ in range(3):
    leaves = [p.clone().requires_grad_(True) for p in pts]
    local_pair.shape_losses(leaves, 20).sum().backward()
torch.cuda.synchronize()
