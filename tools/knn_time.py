"""knnquery timings on the BASELINE shapes (tools/ only).  PDGN_KNN_G32=1 selects the 32-group bound."""
import os, sys
import torch
sys.path.insert(0, ".")
from pdgn_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, reps=20):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps
print("G32" if os.environ.get("PDGN_KNN_G32") else "G64", end=": ")
for (b, n, k) in [(35, 2048, 20), (35, 1024, 20), (35, 512, 20), (35, 256, 20), (35, 2048, 16), (35, 2048, 10), (4, 16384, 32)]:
    xyz = (torch.rand(b, n, 3, generator=g) * 2 - 1).to(dev)
    ms = t(lambda: ops.knn_xyz(k, xyz))
    print("b%d n%d k%d %.4f ms |" % (b, n, k, ms), end=" ")
print()
for (b, n, m, k) in [(35, 2048, 256, 20), (35, 2048, 512, 20), (35, 2048, 1024, 20), (35, 512, 256, 20)]:
    xyz = (torch.rand(b, n, 3, generator=g) * 2 - 1).to(dev)
    q = (torch.rand(b, m, 3, generator=g) * 2 - 1).to(dev)
    print("b%d n%d m%d k%d %.4f ms |" % (b, n, m, k, t(lambda: ops.knn_xyz(k, xyz, q))), end=" ")
print()
for (b, n, m) in [(35, 2048, 1024), (35, 1024, 512), (35, 512, 256)]:
    unk = (torch.rand(b, n, 3, generator=g) * 2 - 1).to(dev)
    kn = (torch.rand(b, m, 3, generator=g) * 2 - 1).to(dev)
    print("nn3 b%d unknown%d known%d %.4f ms |" % (b, n, m, t(lambda: ops.nn3(unk, kn))), end=" ")
print()
