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
are omitted.  Limit: the EMD kernel handles clouds of at most EMD_MAX_POINTS = 2048 points (every PDGN configuration); larger
clouds raise ValueError unless EMD is skipped.  The CD kernels take up to 16384 points per cloud.
"""
import os
import warnings

import numpy as np
import torch

from . import ops
from .chamfer_loss import chamfer_min


EMD_MAX_POINTS = 2048


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
    _check_emd_size(seta.size(1), setb.size(1))
    return ops.emd_paired(seta, setb) * float(seta.size(1))      # one launch, one CTA per pair


def _check_emd_size(n, m):
    if n > EMD_MAX_POINTS or m > EMD_MAX_POINTS:
        raise ValueError("pdgn_b200: the approximate-EMD kernel keeps both clouds on chip and handles at most %d points per "
                         "cloud (got %d and %d); set PDGN_B200_SKIP_EMD=1 to compute the CD metrics only" % (EMD_MAX_POINTS, n, m))


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
    """evaluation_metrics.py:85-121.  Returns (all_cd, all_emd), both [N_sample, N_ref] (all_emd None if skipped).
    CUDA inputs as in the reference's call (PDGNet_v2.py:319); HOST inputs are also accepted (the end-to-end path): the
    copies to the current CUDA device then happen here -- under torch.distributed only this rank's rows / columns -- and the
    matrices come back on that device."""
    from . import dist
    sample_pcs = sample_pcs.contiguous().float()
    ref_pcs = ref_pcs.contiguous().float() if ref_pcs is not sample_pcs else sample_pcs
    skip = _skip_emd()
    if not skip:
        _check_emd_size(sample_pcs.size(1), ref_pcs.size(1))
    if not sample_pcs.is_cuda and (not skip or not dist.is_distributed()):
        dev = torch.device("cuda", torch.cuda.current_device())
        same = ref_pcs is sample_pcs
        sample_pcs = sample_pcs.to(dev, non_blocking=True)
        ref_pcs = sample_pcs if same else ref_pcs.to(dev, non_blocking=True)
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
def _grid_axis(resolution):
    """The per-axis cell centres exactly as the reference stores them: i * spacing - 0.5 in double, rounded to float32
    (evaluation_metrics.py:211-218)."""
    spacing = 1.0 / float(resolution - 1)
    return np.array([i * spacing - 0.5 for i in range(resolution)], dtype=np.float64).astype(np.float32)


def unit_cube_grid_point_cloud(resolution, clip_sphere=False):
    """evaluation_metrics.py:206-224: centres of a resolution^3 grid in the unit cube (numpy, as the reference returns)."""
    spacing = 1.0 / float(resolution - 1)
    ax = _grid_axis(resolution)
    grid = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), axis=-1).astype(np.float32)
    if clip_sphere:
        grid = grid.reshape(-1, 3)
        grid = grid[np.linalg.norm(grid, axis=1) <= 0.5]
    return grid, spacing


_JSD_RERANK = 8


def _nearest_grid_index(points, resolution, in_sphere):
    """Index (into the possibly sphere-clipped grid list) of the nearest grid centre for every point [P,3], reproducing the
    reference's sklearn NearestNeighbors(n_neighbors=1) answer (float64 Euclidean distances to the float32 cell centres,
    evaluation_metrics.py:262-266) EXACTLY rather than to FP32 rounding:
      * the squared distance to a regular grid is separable, so the nearest centre of the FULL grid is the per-axis nearest;
        it is found in float64 among the rounded cell and its two neighbours (float32 differences are exact in float64);
      * when that centre lies inside the sphere it is also the nearest ALLOWED one; the few boundary points for which it
        does not go through the exact FP32 kNN kernel for the %d best allowed centres, re-ranked in float64.""" % _JSD_RERANK
    dev = points.device
    spacing = 1.0 / float(resolution - 1)
    axis = torch.from_numpy(_grid_axis(resolution).astype(np.float64)).to(dev)
    p64 = points.double()
    cell = torch.clamp(torch.round((p64 + 0.5) / spacing), 0, resolution - 1).long()
    best = cell
    best_d = (p64 - axis[cell]).abs()
    for off in (-1, 1):
        cand = torch.clamp(cell + off, 0, resolution - 1)
        d = (p64 - axis[cand]).abs()
        take = (d < best_d) | ((d == best_d) & (cand < best))
        best = torch.where(take, cand, best)
        best_d = torch.where(take, d, best_d)
    flat = (best[:, 0] * resolution + best[:, 1]) * resolution + best[:, 2]
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
        grid = torch.from_numpy(grid_np).to(dev)
        q = points[miss].contiguous()
        k = min(_JSD_RERANK, grid.shape[0])
        cand = ops.knn_xyz(k, grid.unsqueeze(0).contiguous(), q.unsqueeze(0).contiguous()).view(-1, k).long()
        d2 = (q.double().unsqueeze(1) - grid.double()[cand]).pow(2).sum(dim=2)            # [Q, k] float64
        dmin = d2.min(dim=1, keepdim=True).values
        pick = torch.where(d2 == dmin, cand, torch.full_like(cand, grid.shape[0])).min(dim=1).values   # lowest index among equals
        out[miss] = pick
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
