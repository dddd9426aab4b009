import sys
import torch
sys.path.insert(0, ".")
from pdgn_b200 import ops
x = torch.randn(35, 256, 1024, device="cuda")
for _ in range(2):
    ops.knn_feat(x, 10, skip=1)
torch.cuda.synchronize()
