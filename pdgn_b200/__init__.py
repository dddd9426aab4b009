"""pdgn_b200 -- B200 (sm_100a) implementation of PDGN's nearest-neighbour / distance hot path.

Host side is Python/PyTorch (device memory, streams, autograd, torch.distributed); every op runs in
libpdgn_b200.so (hand-written CUDA, C ABI in include/pdgn_b200.h).  Modules mirror the reference's:

  pdgn_b200.pointops            lib/pointops/functions/pointops.py   (knnquery, grouping, nearestneighbor, ...)
  pdgn_b200.chamfer_loss        utils/chamfer_loss.py                (ChamferLoss)
  pdgn_b200.evaluation_metrics  evaluation/evaluation_metrics.py     (distChamfer, _pairwise_EMD_CD_, compute_all_metrics, ...)
  pdgn_b200.edge_features       models/PDGNet_v2.py:439-528          (get_edge_features, get_edge_features_xyz)
  pdgn_b200.dist                multi-GPU 2-D tiling of the all-pairs CD matrix (torch.distributed / NCCL)
  pdgn_b200.dropin              installs the modules above under the reference's import names
"""
from ._lib import PdgnError, lib  # noqa: F401

__version__ = "0.1.0"
