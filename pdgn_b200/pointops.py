"""Mirror of the hot-path part of lib/pointops/functions/pointops.py (reference file:line cited per symbol).

Same names, argument order, shapes, dtypes and gradient behaviour as the reference's autograd Functions, so
`from lib.pointops.functions import pointops` can be pointed here (pdgn_b200.dropin) and models/PDGNet_v2.py runs
unchanged.  Differences, all deliberate: tensors are allocated with torch.empty(device=...) instead of the
removed torch.cuda.FloatTensor constructors; every launch goes to the current stream (the reference's grouping /
interpolation kernels use the legacy NULL stream, grouping_cuda_kernel.cu:85); errors raise instead of exit(-1).
"""
from typing import Tuple

import torch
import torch.nn as nn
from torch.autograd import Function

from . import ops


class KNNQuery(Function):
    """pointops.py:408-434 -> knnquery_cuda_kernel.cu:6-50.  idx int32 (b, m, nsample); no gradient."""

    @staticmethod
    def forward(ctx, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor = None) -> torch.Tensor:
        if new_xyz is None:
            new_xyz = xyz
        assert xyz.is_contiguous()
        assert new_xyz.is_contiguous()
        idx = ops.knn_xyz(nsample, xyz, new_xyz)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None


knnquery = KNNQuery.apply


class KNNQueryNaive(Function):
    """pointops.py:368-405: the authors' pure-torch statement of knnquery (first nsample of a sort).
    Served by the same CUDA kernel; the (d2, index) order makes the unstable-sort ambiguity explicit."""

    @staticmethod
    def forward(ctx, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor = None) -> torch.Tensor:
        if new_xyz is None:
            new_xyz = xyz
        idx = ops.knn_xyz(nsample, xyz.contiguous(), new_xyz.contiguous())
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None


knnquery_naive = KNNQueryNaive.apply


class KNNQueryExclude(Function):
    """pointops.py:437-474: ranks 1..nsample of the sort (rank 0 dropped, whatever it is)."""

    @staticmethod
    def forward(ctx, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor = None) -> torch.Tensor:
        if new_xyz is None:
            new_xyz = xyz
        idx = ops.knn_xyz(nsample + 1, xyz.contiguous(), new_xyz.contiguous())[:, :, 1:].contiguous()
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None


knnquery_exclude = KNNQueryExclude.apply


class NearestNeighbor(Function):
    """pointops.py:61-83 -> nearestneighbor_cuda_kernel_fast (interpolation_cuda_kernel.cu:134-176).
    Returns (sqrt(dist2) (b, n, 3), idx int32 (b, n, 3)); no gradient."""

    @staticmethod
    def forward(ctx, unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        assert unknown.is_contiguous()
        assert known.is_contiguous()
        dist2, idx = ops.nn3(unknown, known)
        dist = torch.sqrt(dist2)
        ctx.mark_non_differentiable(dist, idx)
        return dist, idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


nearestneighbor = NearestNeighbor.apply


class Interpolation(Function):
    """pointops.py:86-119 -> interpolation_{forward_fast,backward} (interpolation_cuda_kernel.cu:181-195, :90-114)."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        assert features.is_contiguous()
        assert idx.is_contiguous()
        assert weight.is_contiguous()
        ctx.interpolation_for_backward = (idx, weight, features.size(2))
        return ops.interp_fwd(features, idx, weight)

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        idx, weight, m = ctx.interpolation_for_backward
        return ops.interp_bwd(grad_out.contiguous(), idx, weight, m), None, None


interpolation = Interpolation.apply


class Grouping(Function):
    """pointops.py:122-151 -> grouping_forward_fast / grouping_backward (grouping_cuda_kernel.cu:60-75, :28-46)."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        assert features.is_contiguous()
        assert idx.is_contiguous()
        ctx.for_backwards = (idx, features.size(2))
        return ops.group_fwd(features, idx)

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        idx, n = ctx.for_backwards
        return ops.group_bwd(grad_out.contiguous(), idx, n), None


grouping = Grouping.apply


def _no_ballquery(*_a, **_k):
    raise NotImplementedError("ballquery (radius search) is outside the PDGN hot path: PDGN always passes radius=None "
                              "(models/PDGNet_v2.py:115); see DESIGN.md 'out of scope'")


ballquery = _no_ballquery


class Gen_QueryAndGroupXYZ(nn.Module):
    """pointops.py:670-703: knnquery -> transpose -> grouping; returns grouped ABSOLUTE xyz (b, 3, m, nsample)."""

    def __init__(self, radius=None, nsample=32, use_xyz=True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor = None) -> torch.Tensor:
        if new_xyz is None:
            new_xyz = xyz
        if self.radius is not None:
            idx = ballquery(self.radius, self.nsample, xyz, new_xyz)
        else:
            idx = knnquery(self.nsample, xyz, new_xyz)
        xyz_trans = xyz.transpose(1, 2).contiguous()
        return grouping(xyz_trans, idx)


class QueryAndGroup(nn.Module):
    """pointops.py:526-569: kNN (radius=None) grouping of xyz (centred on new_xyz) and features."""

    def __init__(self, radius=None, nsample=32, use_xyz=True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz, new_xyz=None, features=None, idx=None):
        if new_xyz is None:
            new_xyz = xyz
        if idx is None:
            idx = ballquery(self.radius, self.nsample, xyz, new_xyz) if self.radius is not None else knnquery(self.nsample, xyz, new_xyz)
        xyz_trans = xyz.transpose(1, 2).contiguous()
        grouped_xyz = grouping(xyz_trans, idx)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is not None:
            grouped_features = grouping(features, idx)
            return torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
        return grouped_xyz


class GroupAll(nn.Module):
    """pointops.py:753-777 (pure tensor reshaping, no kernel)."""

    def __init__(self, use_xyz: bool = True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is not None:
            grouped_features = features.unsqueeze(2)
            return torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        return grouped_xyz
