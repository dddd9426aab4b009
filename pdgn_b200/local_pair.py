"""Fused loss-side of PDGN's shape-preserving loss: get_local_pair (models/PDGNet_v2.py:136-155) in ~12 launches
instead of ~45 (next row SURVEY.md 8f-2).

    like_mu12, like_var12 = get_local_pair(pt1 [B,3,M], pt2 [B,3,N])

Same value and gradients as the reference composition
    Gen_QueryAndGroupXYZ(k=20) on (pt1, pt1) and (pt2, pt1) -> compute_mean_covariance -> ChamferLoss(mu)/M, ChamferLoss(var)/M
but the [B,3,M,k] grouped tensors, their transposes and the dense mean/repeat/bmm chain never exist: kNN indices feed
pdgn_local_stats_fwd directly.  `pdgn_b200.dropin.install()` rebinds PDGNet_v2.get_local_pair to this function.
"""
import torch
from torch.autograd import Function

from . import ops
from .chamfer_loss import chamfer_min


class _LocalStats(Function):
    """(mu [B,M,3], cov [B,M,9]) of the k nearest neighbours (in xyz [B,n,3]) of every query; differentiable w.r.t. xyz."""

    @staticmethod
    def forward(ctx, xyz, queries, k):
        xyz = xyz.contiguous()
        idx = ops.knn_xyz(k, xyz, queries.contiguous())
        mu, cov = ops.local_stats_fwd(xyz, idx)
        ctx.save_for_backward(xyz, idx, mu)
        return mu, cov

    @staticmethod
    def backward(ctx, gmu, gcov):
        xyz, idx, mu = ctx.saved_tensors
        return ops.local_stats_bwd(xyz, idx, mu, gmu.contiguous(), gcov.contiguous()), None, None


local_stats = _LocalStats.apply


def get_local_pair(pt1, pt2, nsample=20):
    """pt1 [B,3,M], pt2 [B,3,N] -> (like_mu12, like_var12), 0-d tensors (PDGNet_v2.py:136-155)."""
    m = pt1.size(2)
    p1 = pt1.transpose(1, 2).contiguous()           # [B,M,3]; also the query set (new_xyz)
    p2 = pt2.transpose(1, 2).contiguous()
    queries = p1.detach()                           # knnquery has no gradient (pointops.py:431-432)
    mu1, var1 = local_stats(p1, queries, nsample)
    mu2, var2 = local_stats(p2, queries, nsample)
    # ChamferLoss(preds, gts) = sum of both directional minima (chamfer_loss.py:13-20)
    a, b = chamfer_min(mu2, mu1)
    c, d = chamfer_min(var2, var1)
    return (a.sum() + b.sum()) / float(m), (c.sum() + d.sum()) / float(m)
