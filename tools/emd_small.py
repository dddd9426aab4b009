"""One small all-pairs EMD launch (24x24 clouds of 2048 points on the unit sphere) for ncu / compute-sanitizer; tools/ only."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from pdgn_b200 import ops
k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rng = np.random.default_rng(0)
v = rng.standard_normal((2 * k, 2048, 3)); v /= np.linalg.norm(v, axis=-1, keepdims=True)
P = torch.from_numpy(v.astype(np.float32)).cuda()
out = ops.emd_allpairs(P[:k].contiguous(), P[k:].contiguous())
torch.cuda.synchronize()
print(float(out.mean()))
