"""Round-2 kernels on small shapes for compute-sanitizer (tools/ only): knn_gram_kernel (forced; all subgroup sizes, ragged
n / m, duplicates -> prune + cooperative exact path), the problem-descriptor launches behind pdgn_shape_loss_fwd/bwd (incl. the
per-problem kNN fallback), the paired EMD entry, interpolation backward on the streaming pull kernel, PDGN_B200_VERIFY.
Usage: PDGN_B200_TUNE=1 PDGN_KNN_IMPL=gram compute-sanitizer --tool memcheck python tools/sanitize_r02.py"""
import sys
import torch
sys.path.insert(0, ".")
from pdgn_b200 import local_pair, ops, pointops
g = torch.Generator().manual_seed(2)
dev = torch.device("cuda:0")
R = lambda *s: (torch.rand(*s, generator=g) * 2 - 1).to(dev)
for (b, n, m, k) in [(2, 2048, 600, 20), (1, 2039, 513, 16), (2, 1001, 130, 20), (2, 501, 40, 20), (1, 257, 257, 1), (1, 1500, 33, 7)]:
    ops.knn_xyz(k, R(b, n, 3), R(b, m, 3), return_dist=True)
d = R(1, 1024, 3)
d[:, 512:] = d[:, :512]
d[:, :200] = d[:, :1]
ops.knn_xyz(20, d.contiguous(), R(1, 70, 3), return_dist=True)          # duplicates: prune loop + exact path
for npts in [(256, 512, 1024, 2048), (300, 700), (64, 300, 512)]:
    leaves = [(R(2, 3, n) * 0.5).requires_grad_(True) for n in npts]
    local_pair.shape_losses(leaves, 20).sum().backward()
ops.emd_paired(R(3, 300, 3), R(3, 300, 3))
f = R(2, 16, 300).requires_grad_(True)
i3 = torch.randint(0, 300, (2, 640, 3), generator=g, dtype=torch.int32).to(dev)
w = torch.rand(2, 640, 3, generator=g).to(dev)
pointops.interpolation(f, i3, w).sum().backward()                       # streaming pull, MODE 2
torch.cuda.synchronize()
print("sanitize r02 target done")
