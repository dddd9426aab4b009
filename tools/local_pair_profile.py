"""One fused get_local_pair forward+backward per training shape, for an ncu launch list (tools/ only)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from pdgn_b200 import local_pair
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
def cloud(n):
    v = rng.standard_normal((35, 3, n)); v /= np.linalg.norm(v, axis=1, keepdims=True)
    return torch.from_numpy(v.astype(np.float32)).to(dev).requires_grad_(True)
for rep in range(2):
    for (m, n) in [(256, 2048), (512, 2048), (1024, 2048), (256, 512), (512, 1024), (1024, 2048)]:
        p1, p2 = cloud(m), cloud(n)
        a, b = local_pair.get_local_pair(p1, p2)
        (a + b).backward()
torch.cuda.synchronize()
