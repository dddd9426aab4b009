"""Fused loss-side of PDGN's shape-preserving loss: get_local_pair (models/PDGNet_v2.py:136-155) in ~12 launches
instead of ~45, enqueued by ONE C call per direction (next row SURVEY.md 8f-2).

    like_mu12, like_var12 = get_local_pair(pt1 [B,3,M], pt2 [B,3,N])

Same value and gradients as the reference composition
    Gen_QueryAndGroupXYZ(k=20) on (pt1, pt1) and (pt2, pt1) -> compute_mean_covariance -> ChamferLoss(mu)/M, ChamferLoss(var)/M
but the [B,3,M,k] grouped tensors, their transposes and the dense mean/repeat/bmm chain never exist: kNN indices feed
pdgn_local_stats_fwd directly.  `pdgn_b200.dropin.install()` rebinds PDGNet_v2.get_local_pair to this function.
"""
import ctypes
import os
import weakref

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


class _ShapeLosses(Function):
    """All level pairs of one generator step behind one C call per direction (csrc/shape_loss.cu): every operator runs all
    its problems of the step from a descriptor table -- 6 launches forward, 4 + a memset backward."""

    @staticmethod
    def forward(ctx, k, *pts):
        pts = [p.contiguous() for p in pts]
        levels = len(pts)
        b = pts[0].shape[0]
        npts = (ctypes.c_int * levels)(*[int(p.shape[2]) for p in pts])
        ptrs = (ctypes.c_void_p * levels)(*[p.data_ptr() for p in pts])
        dev = pts[0].device
        L = lib()
        ws_bytes = L.pdgn_shape_loss_workspace(b, levels, npts, k)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        pairs = levels * (levels - 1) // 2
        out = torch.empty((2 * pairs,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(L.pdgn_shape_loss_fwd(ptrs, b, levels, npts, k, out.data_ptr(), ws.data_ptr(), ws_bytes,
                                        torch.cuda.current_stream(dev).cuda_stream), "pdgn_shape_loss_fwd")
        ctx.ws, ctx.dims, ctx.shapes = ws, (b, levels, k), [tuple(p.shape) for p in pts]
        ctx.need = [p.requires_grad for p in pts]
        return out

    @staticmethod
    def backward(ctx, gout):
        b, levels, k = ctx.dims
        ws = ctx.ws
        dev = ws.device
        gout = gout.to(torch.float32).contiguous()
        grads = [torch.zeros(shape, dtype=torch.float32, device=dev) if need else None for shape, need in zip(ctx.shapes, ctx.need)]
        npts = (ctypes.c_int * levels)(*[s[2] for s in ctx.shapes])
        gptrs = (ctypes.c_void_p * levels)(*[g.data_ptr() if g is not None else None for g in grads])
        with torch.cuda.device(dev):
            check(lib().pdgn_shape_loss_bwd(b, levels, npts, k, gout.data_ptr(), gptrs, ws.data_ptr(), ws.numel(),
                                            torch.cuda.current_stream(dev).cuda_stream), "pdgn_shape_loss_bwd")
        return (None, *grads)


def shape_losses(points, nsample=20):
    """points = the generator's outputs [p1, p2, p3, p4] ([B,3,N_l] each, 2..4 levels) -> float32 tensor [2 * pairs]:
    (like_mu, like_var) of get_local_pair(p_a, p_c) for the level pairs (a < c) in the order PDGNet_v2.train evaluates them
    (PDGNet_v2.py:232-237): (1,2) (1,3) (1,4) (2,3) (2,4) (3,4).  Differentiable w.r.t. every level."""
    points = list(points)
    if not 2 <= len(points) <= 4:
        raise ValueError("shape_losses takes 2 to 4 levels")
    for p in points:
        if not (isinstance(p, torch.Tensor) and p.is_cuda and p.dtype == torch.float32 and p.dim() == 3 and p.shape[1] == 3):
            raise TypeError("shape_losses takes CUDA float32 tensors [B,3,N]")
    return _ShapeLosses.apply(int(nsample), *points)


# ---- speculative batching for the unmodified trainer -------------------------------------------------------------------
# PDGNet_v2.train calls get_local_pair six times on the four outputs of ONE generator forward.  pdgn_b200.dropin wraps
# PointGenerator.forward to note its outputs here; the first get_local_pair call whose arguments are two of them evaluates
# ALL level pairs at once (shape_losses) and the other five calls are answered from that result.  Nothing is assumed about
# the call order, and arguments that are not noted generator outputs take the per-call path.
_noted = None


def note_generator_outputs(outputs):
    """Called by the drop-in's PointGenerator.forward wrapper with the tuple of level outputs."""
    global _noted
    outs = [o for o in outputs if isinstance(o, torch.Tensor)]
    ok = 2 <= len(outs) <= 4 and all(o.is_cuda and o.dtype == torch.float32 and o.dim() == 3 and o.shape[1] == 3 for o in outs)
    _noted = {"refs": [weakref.ref(o) for o in outs], "versions": [o._version for o in outs], "result": None} if ok else None


def _noted_pair(pt1, pt2, nsample):
    st = _noted
    if st is None or os.environ.get("PDGN_B200_LOCAL_PAIR", "call") == "each":
        return None
    outs = [r() for r in st["refs"]]
    if any(o is None for o in outs) or any(o._version != v for o, v in zip(outs, st["versions"])):
        return None
    ia = next((i for i, o in enumerate(outs) if o is pt1), None)
    ic = next((i for i, o in enumerate(outs) if o is pt2), None)
    if ia is None or ic is None or ia >= ic:
        return None
    if st["result"] is None or st["result"][0] != nsample:
        st["result"] = (nsample, shape_losses(outs, nsample))
    levels = len(outs)
    p = sum(levels - 1 - a for a in range(ia)) + (ic - ia - 1)
    res = st["result"][1]
    return res[2 * p], res[2 * p + 1]


def get_local_pair(pt1, pt2, nsample=20):
    """pt1 [B,3,M], pt2 [B,3,N] -> (like_mu12, like_var12), 0-d tensors (PDGNet_v2.py:136-155).

    One C call forward, one backward (PDGN_B200_LOCAL_PAIR=ops selects the op-by-op composition below, same kernels).  When
    both arguments are outputs of the generator forward the drop-in last noted, all level pairs of the step are evaluated in
    one batched call and this call returns its two entries (PDGN_B200_LOCAL_PAIR=each disables that)."""
    if nsample <= 64:
        hit = _noted_pair(pt1, pt2, int(nsample))
        if hit is not None:
            return hit
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
