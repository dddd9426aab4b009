"""Mirror of get_edge_features / get_edge_features_xyz (models/PDGNet_v2.py:439-477, :479-528).

Same signatures and outputs ([B,2C,N,k] = cat(central, neighbour - central), plus [B,6,N,k] for xyz), same
differentiability (gradients flow to x and pc through the gather; indices are constants).  The [B,N,N] Gram
matrix, the full torch.sort, the Python loop of 2*B index_select calls and the repeat+cat copies are replaced by
two kernels: pdgn_knn_feat (exact FP32 feature-space kNN, ranks 1..k) and pdgn_edge_feat_fwd (gather + centring
written once), with pdgn_edge_feat_bwd as the backward.
"""
import torch
from torch.autograd import Function

from . import ops


class _EdgeFeat(Function):
    @staticmethod
    def forward(ctx, x, idx):
        x = x.contiguous()
        ctx.save_for_backward(idx)
        ctx.c = x.size(1)
        return ops.edge_feat_fwd(x, idx)

    @staticmethod
    def backward(ctx, grad_ee):
        (idx,) = ctx.saved_tensors
        return ops.edge_feat_bwd(grad_ee.contiguous(), idx, ctx.c), None


edge_feat = _EdgeFeat.apply


def feature_knn(x, k):
    """idx int64 [B,N,k]: ranks 1..k of the ascending (d2, index) order in feature space (rank 0 dropped)."""
    with torch.no_grad():
        return ops.knn_feat(x.detach().contiguous().float(), k, skip=1)


def get_edge_features(x, k, num=-1):
    """x [B,C,N] -> [B,2C,N,k]  (PDGNet_v2.py:439-477)."""
    idx = feature_knn(x, k)
    return edge_feat(x.float(), idx)


def get_edge_features_xyz(x, pc, k, num=-1):
    """x [B,C,N], pc [B,3,N] -> (e_fea [B,2C,N,k], e_xyz [B,6,N,k])  (PDGNet_v2.py:479-528); kNN in feature space."""
    idx = feature_knn(x, k)
    return edge_feat(x.float(), idx), edge_feat(pc.float(), idx)
