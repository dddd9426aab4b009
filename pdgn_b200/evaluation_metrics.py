"""Mirror of the CD half of evaluation/evaluation_metrics.py (reference file:line cited per symbol).

Same names, arguments and result keys as the reference, so `from evaluation.evaluation_metrics import *` can be
pointed here (pdgn_b200.dropin).  What changes is where the work happens:
  * _pairwise_EMD_CD_ (:85-121): the Python double loop of distChamfer calls (20 000 iterations per 1000x1000
    matrix) becomes ONE launch of the all-pairs kernel (csrc/cd_allpairs.cu); `batch_size` is accepted and ignored.
    Under torch.distributed the matrix is 2-D tiled over the ranks (pdgn_b200.dist).
  * distChamfer / distChamferCUDA (:35-45, :22-23): the paired min-distance kernel (csrc/chamfer.cu).
  * lgan_mmd_cov, knn, compute_all_metrics (:125-200): unchanged torch reductions on the [N,N] matrices.
  * emd_approx / match_cost (:26-31, match_cost.py) and the EMD half of _pairwise_EMD_CD_: the all-pairs approximate-EMD
    kernel (csrc/emd.cu), forward only -- the reference uses EMD only in evaluation.
Set PDGN_B200_SKIP_EMD=1 to skip the EMD matrices (they cost ~35x the CD ones): all_emd is then None and the *-EMD keys
are omitted.
"""
import os
import warnings

import numpy as np
import torch

from . import ops
from .chamfer_loss import chamfer_min


def distChamferCUDA(x, y):
    """evaluation_metrics.py:22-23 -> nn_distance (nn_distance.py:6-41): (dist1 [B,Nx], dist2 [B,Ny]), differentiable."""
    return chamfer_min(x, y)


def distChamfer(a, b):
    """evaluation_metrics.py:35-45.  Returns (P.min(1)[0], P.min(2)[0]) = (per-b-point min over a, per-a-point min over b)."""
    m_ab, m_ba = chamfer_min(a, b)
    return m_ba, m_ab


def match_cost(seta, setb):
    """match_cost.py:6-44, forward only: per-pair approximate-EMD matching cost [B] of seta [B,n,3] vs setb [B,m,3]."""
    seta = seta.detach().contiguous().float()
    setb = setb.detach().contiguous().float()
    n = seta.size(1)
    out = torch.empty((seta.size(0),), dtype=torch.float32, device=seta.device)
    for i in range(seta.size(0)):  # paired form = diagonal of the all-pairs problem, one 1x1 tile per pair
        out[i:i + 1] = ops.emd_allpairs(seta, setb, rows=(i, i + 1), cols=(i, i + 1)).view(1) * float(n)
    return out


def emd_approx(sample, ref):
    """evaluation_metrics.py:26-31: match_cost / N."""
    B, N, N_ref = sample.size(0), sample.size(1), ref.size(1)
    assert N == N_ref, "Not sure what would EMD do in this case"
    return match_cost(sample, ref) / float(N)


def _skip_emd():
    return os.environ.get("PDGN_B200_SKIP_EMD", "0") not in ("", "0")


def EMD_CD(sample_pcs, ref_pcs, batch_size, accelerated_cd=False, reduced=True):
    """evaluation_metrics.py:48-82, CD part: paired (not all-pairs) Chamfer distance."""
    N_sample, N_ref = sample_pcs.shape[0], ref_pcs.shape[0]
    assert N_sample == N_ref, "REF:%d SMP:%d" % (N_ref, N_sample)
    cd_lst = []
    for b_start in range(0, N_sample, batch_size):
        b_end = min(N_sample, b_start + batch_size)
        dl, dr = distChamfer(sample_pcs[b_start:b_end].contiguous(), ref_pcs[b_start:b_end].contiguous())
        cd_lst.append(dl.mean(dim=1) + dr.mean(dim=1))
    cd = torch.cat(cd_lst).mean() if reduced else torch.cat(cd_lst)
    if _skip_emd():
        return {"MMD-CD": cd}
    emd = emd_approx(sample_pcs.contiguous(), ref_pcs.contiguous())
    return {"MMD-CD": cd, "MMD-EMD": emd.mean() if reduced else emd}


def _pairwise_EMD_CD_(sample_pcs, ref_pcs, batch_size=None, accelerated_cd=True):
    """evaluation_metrics.py:85-121.  Returns (all_cd, all_emd), both [N_sample, N_ref] (all_emd None if skipped)."""
    from . import dist
    sample_pcs = sample_pcs.contiguous().float()
    ref_pcs = ref_pcs.contiguous().float()
    skip = _skip_emd()
    if dist.is_distributed():
        return dist.pairwise_cd(sample_pcs, ref_pcs), (None if skip else dist.pairwise_emd(sample_pcs, ref_pcs))
    return ops.cd_allpairs(sample_pcs, ref_pcs), (None if skip else ops.emd_allpairs(sample_pcs, ref_pcs))


def knn(Mxx, Mxy, Myy, k, sqrt=False):
    """evaluation_metrics.py:125-154: k-NN two-sample test on the [n0+n1, n0+n1] block matrix [[Mxx, Mxy], [Mxy^T, Myy]]
    with the diagonal excluded; a point is predicted "x" when at least k/2 of its k nearest others are x.  Same reductions
    and key set as the reference (tp/fp/fn/tn, precision, recall, acc_t, acc_f, acc), run on the GPU."""
    n0, n1 = Mxx.size(0), Myy.size(0)
    n = n0 + n1
    is_x = Mxx.new_zeros(n)
    is_x[:n0] = 1.0
    M = Mxx.new_empty((n, n))
    M[:n0, :n0] = Mxx
    M[:n0, n0:] = Mxy
    M[n0:, :n0] = Mxy.t()
    M[n0:, n0:] = Myy
    if sqrt:
        M = M.abs().sqrt()
    M.diagonal().fill_(float("inf"))                     # a cloud is never its own neighbour
    nearest = M.topk(k, dim=0, largest=False).indices    # [k, n]
    votes = is_x[nearest].sum(dim=0)
    pred = (votes >= float(k) / 2).to(Mxx.dtype)
    not_x, not_pred = 1 - is_x, 1 - pred
    s = {"tp": (pred * is_x).sum(), "fp": (pred * not_x).sum(), "fn": (not_pred * is_x).sum(), "tn": (not_pred * not_x).sum()}
    s["precision"] = s["tp"] / (s["tp"] + s["fp"] + 1e-10)
    s["recall"] = s["tp"] / (s["tp"] + s["fn"] + 1e-10)
    s["acc_t"] = s["tp"] / (s["tp"] + s["fn"] + 1e-10)
    s["acc_f"] = s["tn"] / (s["tn"] + s["fp"] + 1e-10)
    s["acc"] = (is_x == pred).to(Mxx.dtype).mean()
    return s


def lgan_mmd_cov(all_dist):
    """evaluation_metrics.py:157-169 on an [N_sample, N_ref] matrix: MMD = mean over refs of the distance to the closest
    sample, COV = fraction of refs that are the closest ref of some sample, MMD-smp = mean over samples of their closest ref."""
    n_ref = all_dist.size(1)
    closest_ref_val, closest_ref = all_dist.min(dim=1)
    covered = torch.unique(closest_ref).numel()
    return {
        "lgan_mmd": all_dist.min(dim=0).values.mean(),
        "lgan_cov": torch.tensor(float(covered) / float(n_ref)).to(all_dist),
        "lgan_mmd_smp": closest_ref_val.mean(),
    }


def compute_all_metrics(sample_pcs, ref_pcs, batch_size=None, accelerated_cd=False):
    """evaluation_metrics.py:172-200: MMD / COV / 1-NNA from the three all-pairs matrices (rs, rr, ss)."""
    results = {}
    M_rs_cd, M_rs_emd = _pairwise_EMD_CD_(sample_pcs, ref_pcs, batch_size, accelerated_cd=accelerated_cd)
    results.update({"%s-CD" % k: v for k, v in lgan_mmd_cov(M_rs_cd.t()).items()})
    if M_rs_emd is not None:
        results.update({"%s-EMD" % k: v for k, v in lgan_mmd_cov(M_rs_emd.t()).items()})
    M_rr_cd, M_rr_emd = _pairwise_EMD_CD_(ref_pcs, ref_pcs, batch_size, accelerated_cd=accelerated_cd)
    M_ss_cd, M_ss_emd = _pairwise_EMD_CD_(sample_pcs, sample_pcs, batch_size, accelerated_cd=accelerated_cd)
    one_nn_cd_res = knn(M_rr_cd, M_rs_cd, M_ss_cd, 1, sqrt=False)
    results.update({"1-NN-CD-%s" % k: v for k, v in one_nn_cd_res.items() if "acc" in k})
    if M_rs_emd is not None:
        one_nn_emd_res = knn(M_rr_emd, M_rs_emd, M_ss_emd, 1, sqrt=False)
        results.update({"1-NN-EMD-%s" % k: v for k, v in one_nn_emd_res.items() if "acc" in k})
    else:
        warnings.warn("pdgn_b200: PDGN_B200_SKIP_EMD is set; only the -CD keys are returned", stacklevel=2)
    return results


#######################################################
# JSD (evaluation_metrics.py:206-321) -- next row 8f-4
#######################################################
def unit_cube_grid_point_cloud(resolution, clip_sphere=False):
    """evaluation_metrics.py:206-224: centres of a resolution^3 grid in the unit cube (numpy, as the reference returns)."""
    spacing = 1.0 / float(resolution - 1)
    ax = (np.arange(resolution, dtype=np.float32) * np.float32(spacing) - np.float32(0.5)).astype(np.float32)
    grid = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), axis=-1).astype(np.float32)
    if clip_sphere:
        grid = grid.reshape(-1, 3)
        grid = grid[np.linalg.norm(grid, axis=1) <= 0.5]
    return grid, spacing


def _nearest_grid_index(points, resolution, in_sphere):
    """Index (into the possibly sphere-clipped grid list) of the nearest grid centre for every point [P,3] on the GPU.
    The nearest centre of the FULL regular grid is a rounding; it is also the nearest ALLOWED centre whenever it lies inside
    the sphere (true for all but a handful of boundary points), and those few go through the exact 1-NN kernel."""
    dev = points.device
    spacing = 1.0 / float(resolution - 1)
    cell = torch.clamp(torch.round((points + 0.5) / spacing), 0, resolution - 1).long()
    flat = (cell[:, 0] * resolution + cell[:, 1]) * resolution + cell[:, 2]
    if not in_sphere:
        return flat, resolution ** 3
    grid_np, _ = unit_cube_grid_point_cloud(resolution, True)
    full_np, _ = unit_cube_grid_point_cloud(resolution, False)
    keep = np.linalg.norm(full_np.reshape(-1, 3), axis=1) <= 0.5
    lut = torch.full((resolution ** 3,), -1, dtype=torch.long, device=dev)
    lut[torch.from_numpy(np.nonzero(keep)[0]).to(dev)] = torch.arange(int(keep.sum()), device=dev)
    out = lut[flat]
    miss = out < 0
    if bool(miss.any()):
        grid = torch.from_numpy(grid_np).to(dev).unsqueeze(0).contiguous()
        q = points[miss].unsqueeze(0).contiguous()
        out[miss] = ops.knn_xyz(1, grid, q).view(-1).long()
    return out, int(keep.sum())


def entropy_of_occupancy_grid(pclouds, grid_resolution, in_sphere=False, verbose=False):
    """evaluation_metrics.py:241-280 on the GPU: (mean cell entropy, per-cell point counters [numpy float64])."""
    pcs = torch.as_tensor(pclouds, dtype=torch.float32)
    if not pcs.is_cuda:
        pcs = pcs.cuda()
    n_clouds, n_pts, _ = pcs.shape
    epsilon = 10e-4
    bound = 0.5 + epsilon
    if verbose and (abs(pcs.max().item()) > bound or abs(pcs.min().item()) > bound):
        warnings.warn("Point-clouds are not in unit cube.")
    if verbose and in_sphere and pcs.pow(2).sum(dim=2).sqrt().max().item() > bound:
        warnings.warn("Point-clouds are not in unit sphere.")
    idx, n_cells = _nearest_grid_index(pcs.reshape(-1, 3).contiguous(), grid_resolution, in_sphere)
    counters = torch.bincount(idx, minlength=n_cells).double()
    cloud_id = torch.arange(n_clouds, device=pcs.device).repeat_interleave(n_pts)
    per_cloud = torch.unique(cloud_id * n_cells + idx)
    bern = torch.bincount(per_cloud % n_cells, minlength=n_cells).double()
    p = bern[bern > 0] / float(n_clouds)
    q = 1.0 - p
    ent = -(p * torch.log(p)) - torch.where(q > 0, q * torch.log(torch.clamp(q, min=1e-300)), torch.zeros_like(q))
    return float(ent.sum().item()) / float(n_cells), counters.cpu().numpy()


def jensen_shannon_divergence(P, Q):
    """evaluation_metrics.py:283-302 (base-2 entropies; tiny vectors, host side)."""
    P = np.asarray(P, dtype=np.float64)
    Q = np.asarray(Q, dtype=np.float64)
    if np.any(P < 0) or np.any(Q < 0):
        raise ValueError("Negative values.")
    if len(P) != len(Q):
        raise ValueError("Non equal size.")
    P_ = P / np.sum(P)
    Q_ = Q / np.sum(Q)

    def _entropy2(v):
        v = v[v > 0]
        return float(-(v * np.log2(v)).sum())

    res = _entropy2((P_ + Q_) / 2.0) - (_entropy2(P_) + _entropy2(Q_)) / 2.0
    res2 = _jsdiv(P_, Q_)
    if not np.allclose(res, res2, atol=10e-5, rtol=0):
        warnings.warn("Numerical values of two JSD methods don't agree.")
    return res


def _jsdiv(P, Q):
    """evaluation_metrics.py:305-321."""
    def _kldiv(A, B):
        idx = np.logical_and(A > 0, B > 0)
        a, b = A[idx], B[idx]
        return float(np.sum(a * np.log2(a / b)))

    P_ = P / np.sum(P)
    Q_ = Q / np.sum(Q)
    M = 0.5 * (P_ + Q_)
    return 0.5 * (_kldiv(P_, M) + _kldiv(Q_, M))


def jsd_between_point_cloud_sets(sample_pcs, ref_pcs, resolution=28):
    """evaluation_metrics.py:227-238: JSD between the occupancy-grid distributions of two cloud sets (numpy or torch in)."""
    in_unit_sphere = True
    sample_grid_var = entropy_of_occupancy_grid(sample_pcs, resolution, in_unit_sphere)[1]
    ref_grid_var = entropy_of_occupancy_grid(ref_pcs, resolution, in_unit_sphere)[1]
    return jensen_shannon_divergence(sample_grid_var, ref_grid_var)
