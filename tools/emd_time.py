"""Time the all-pairs approximate-EMD kernel (tools/; not part of the product)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from pdgn_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148
rng = np.random.default_rng(0)
def sph(k):
    v = rng.standard_normal((k, 2048, 3)); v /= np.linalg.norm(v, axis=-1, keepdims=True)
    return torch.from_numpy(v.astype(np.float32)).cuda()
A, B = sph(n), sph(n)
ops.emd_allpairs(A, B); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = ops.emd_allpairs(A, B); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
rate = n * n / ms * 1e3
print("EMD all-pairs %dx%dx2048: %.1f ms -> %.0f cloud-pairs/s -> 1000x1000 in %.1f s" % (n, n, ms, rate, 1e6 / rate))
