"""Fused loss-side of PDGN's shape-preserving loss: get_local_pair (models/PDGNet_v2.py:136-155) in ~12 launches
instead of ~45, enqueued by ONE C call per direction (next row SURVEY.md 8f-2).

    like_mu12, like_var12 = get_local_pair(pt1 [B,3,M], pt2 [B,3,N])

Same value and gradients as the reference composition
    Gen_QueryAndGroupXYZ(k=20) on (pt1, pt1) and (pt2, pt1) -> compute_mean_covariance -> ChamferLoss(mu)/M, ChamferLoss(var)/M
but the [B,3,M,k] grouped tensors, their transposes and the dense mean/repeat/bmm chain never exist: kNN indices feed
pdgn_local_stats_fwd directly.  `pdgn_b200.dropin.install()` rebinds PDGNet_v2.get_local_pair to this function.
"""
import os

import torch
from torch.autograd import Function

from . import ops
from ._lib import check, lib
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


class _LocalPairCall(Function):
    """The whole chain behind one C call per direction (csrc/local_pair.cu): no torch op between the kernels."""

    @staticmethod
    def forward(ctx, pt1, pt2, k):
        pt1, pt2 = pt1.contiguous(), pt2.contiguous()
        b, _, m = pt1.shape
        n = pt2.shape[2]
        L = lib()
        ws_bytes = L.pdgn_local_pair_workspace(b, m, n, k)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=pt1.device)
        out = torch.empty((2,), dtype=torch.float32, device=pt1.device)
        with torch.cuda.device(pt1.device):
            check(L.pdgn_local_pair_fwd(pt1.data_ptr(), pt2.data_ptr(), b, m, n, k, out.data_ptr(), ws.data_ptr(), ws_bytes,
                                        torch.cuda.current_stream(pt1.device).cuda_stream), "pdgn_local_pair_fwd")
        ctx.ws, ctx.dims = ws, (b, m, n, k)
        return out.unbind(0)

    @staticmethod
    def backward(ctx, g_mu, g_var):
        b, m, n, k = ctx.dims
        ws = ctx.ws
        gout = torch.stack([g_mu, g_var]).to(torch.float32).contiguous()
        gp1 = torch.zeros((b, 3, m), dtype=torch.float32, device=ws.device)
        gp2 = torch.zeros((b, 3, n), dtype=torch.float32, device=ws.device)
        with torch.cuda.device(ws.device):
            check(lib().pdgn_local_pair_bwd(b, m, n, k, gout.data_ptr(), gp1.data_ptr(), gp2.data_ptr(), ws.data_ptr(), ws.numel(),
                                            torch.cuda.current_stream(ws.device).cuda_stream), "pdgn_local_pair_bwd")
        return gp1, gp2, None


def get_local_pair(pt1, pt2, nsample=20):
    """pt1 [B,3,M], pt2 [B,3,N] -> (like_mu12, like_var12), 0-d tensors (PDGNet_v2.py:136-155).

    One C call forward, one backward (PDGN_B200_LOCAL_PAIR=ops selects the op-by-op composition below, same kernels)."""
    if os.environ.get("PDGN_B200_LOCAL_PAIR", "call") != "ops" and nsample <= 64:
        if pt1.dtype != torch.float32 or pt2.dtype != torch.float32 or not pt1.is_cuda or not pt2.is_cuda:
            raise TypeError("get_local_pair takes CUDA float32 tensors [B,3,M] / [B,3,N]")
        return _LocalPairCall.apply(pt1, pt2, int(nsample))
    return get_local_pair_ops(pt1, pt2, nsample)


def get_local_pair_ops(pt1, pt2, nsample=20):
    """The same chain composed from the individual ops (autograd through _LocalStats and chamfer_min)."""
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
