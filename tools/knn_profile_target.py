"""ncu target: knnquery on the cfg2 shape (B=35, n=m=2048, k=20, uniform cube).  tools/ only."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from pdgn_b200 import ops
rng = np.random.default_rng(0)
k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
xyz = torch.from_numpy(rng.uniform(-1, 1, (35, 2048, 3)).astype(np.float32)).cuda()
for _ in range(3):
    ops.knn_xyz(k, xyz)
torch.cuda.synchronize()
