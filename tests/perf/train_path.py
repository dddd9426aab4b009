"""BASELINE config 3 (training path) timing: the generator's four feature-space kNN stages and the 12 loss-side
(kNN -> grouping -> mean/cov -> ChamferLoss) problems of one G step (models/PDGNet_v2.py:232-237, :864-877), forward+backward,
through pdgn_b200's public modules, next to the reference's own torch formulations run on the same GPU (oracle.torch_ref; the
reference's pointops glue does not build on torch 2.x, so its knn+group leg is driven through oracle/_ref).
Lives under tests/ because it executes oracle/ (test infrastructure); not collected by pytest.
Usage (GPU box): python tests/perf/train_path.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref_kernels as rk  # noqa: E402
from oracle import torch_ref as tref  # noqa: E402
from pdgn_b200 import edge_features as ef  # noqa: E402
from pdgn_b200 import pointops  # noqa: E402
from pdgn_b200.chamfer_loss import ChamferLoss  # noqa: E402

dev = torch.device("cuda:0")
B = 35
rng = np.random.default_rng(0)


def ev_time(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def cloud(n):
    v = rng.standard_normal((B, n, 3))
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    return torch.from_numpy(v.astype(np.float32)).to(dev)


def mean_cov(points):  # compute_mean_covariance, PDGNet_v2.py:127-134 (dense torch, not on the replaced path)
    mu = points.mean(dim=-1, keepdim=True)
    tmp = points - mu
    return mu, torch.bmm(tmp, tmp.transpose(1, 2)) / points.size(2)


def local_pair(group, chamfer, pt1, pt2):
    """get_local_pair (PDGNet_v2.py:136-155) with the grouping module / chamfer loss passed in."""
    b, _, m1 = pt1.size()
    new_xyz = pt1.transpose(1, 2).contiguous()
    g1 = group(pt1.transpose(1, 2).contiguous(), new_xyz).transpose(1, 2).contiguous().view(-1, 3, 20)
    g2 = group(pt2.transpose(1, 2).contiguous(), new_xyz).transpose(1, 2).contiguous().view(-1, 3, 20)
    mu1, var1 = mean_cov(g1)
    mu2, var2 = mean_cov(g2)
    return chamfer(mu1.view(b, -1, 3), mu2.view(b, -1, 3)) / float(m1) + chamfer(var1.view(b, -1, 9), var2.view(b, -1, 9)) / float(m1)


class RefGroup(torch.nn.Module):
    """Gen_QueryAndGroupXYZ through the reference's own recompiled kernels (forward only; its autograd glue is unbuildable)."""
    def forward(self, xyz, new_xyz):
        idx, _ = rk.knnquery(20, xyz, new_xyz)
        return rk.group_fwd(xyz.transpose(1, 2).contiguous(), idx)


class RefChamfer(torch.nn.Module):
    def forward(self, preds, gts):
        return tref.chamfer_loss(preds, gts)


sizes = [(256, 512), (256, 1024), (256, 2048), (512, 1024), (512, 2048), (1024, 2048)]
pts = {n: cloud(n).transpose(1, 2).contiguous() for n in (256, 512, 1024, 2048)}  # [B,3,n] like the generator outputs


def loss_side(group, chamfer, backward):
    leaves = {n: p.clone().requires_grad_(backward) for n, p in pts.items()}
    total = 0
    for m_, n_ in sizes:
        total = total + local_pair(group, chamfer, leaves[m_], leaves[n_])
    if backward:
        total.backward()
    return total


ours_group = pointops.Gen_QueryAndGroupXYZ(radius=None, nsample=20, use_xyz=False)
t_ours = ev_time(lambda: loss_side(ours_group, ChamferLoss(), True))
t_ours_f = ev_time(lambda: loss_side(ours_group, ChamferLoss(), False))
t_ref_f = ev_time(lambda: loss_side(RefGroup(), RefChamfer(), False), reps=2, warm=1)
from pdgn_b200 import local_pair as fused  # noqa: E402


def loss_side_fused(backward, fn=None):
    fn = fn or fused.get_local_pair
    leaves = {n: p.clone().requires_grad_(backward) for n, p in pts.items()}
    total = 0
    for m_, n_ in sizes:
        a, b = fn(leaves[m_], leaves[n_])
        total = total + a + b
    if backward:
        total.backward()
    return total


t_fused = ev_time(lambda: loss_side_fused(True))
t_fused_f = ev_time(lambda: loss_side_fused(False))
t_fops = ev_time(lambda: loss_side_fused(True, fused.get_local_pair_ops))
t_fops_f = ev_time(lambda: loss_side_fused(False, fused.get_local_pair_ops))
print("loss side, 6 x get_local_pair (12 kNN+group, 12 Chamfer), B=35:")
print("  pdgn_b200 get_local_pair, one C call per direction  fwd+bwd %8.3f ms   fwd %8.3f ms" % (t_fused, t_fused_f))
print("  pdgn_b200 get_local_pair, fused ops from Python     fwd+bwd %8.3f ms   fwd %8.3f ms" % (t_fops, t_fops_f))
print("  pdgn_b200 op-by-op composition  fwd+bwd %8.3f ms   fwd %8.3f ms" % (t_ours, t_ours_f))
print("  reference kernels + torch Gram Chamfer on this GPU, fwd only %8.3f ms" % t_ref_f)

print("generator feature-space kNN stages (get_edge_features_xyz fwd+bwd), B=35, k=10:")
for c, n in [(32, 128), (64, 256), (128, 512), (256, 1024)]:
    x = torch.randn(B, c, n, device=dev)
    pc = torch.rand(B, 3, n, device=dev) * 2 - 1

    def ours():
        xr, pr = x.clone().requires_grad_(True), pc.clone().requires_grad_(True)
        e_fea, e_xyz = ef.get_edge_features_xyz(xr, pr, 10)
        (e_fea.sum() + e_xyz.sum()).backward()

    def ref():
        xr, pr = x.clone().requires_grad_(True), pc.clone().requires_grad_(True)
        e_fea, e_xyz, _ = tref.get_edge_features_xyz(xr, pr, 10)
        (e_fea.sum() + e_xyz.sum()).backward()

    print("  C=%3d N=%4d   pdgn_b200 %8.3f ms   reference torch formulation on this GPU %8.3f ms" % (c, n, ev_time(ours), ev_time(ref, reps=2, warm=1)))
