"""Functional wrappers: torch tensors in, torch tensors out, straight onto the C ABI (include/pdgn_b200.h).

No autograd here (see pointops.py / chamfer_loss.py / edge_features.py for the Functions) and no fallbacks:
inputs must be CUDA FP32/int32/int64 contiguous tensors, as the reference asserts (pointops.py:421-422).
Every call is enqueued on torch's current stream of the tensors' device.
"""
import torch

from ._lib import check, lib


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _req(t, name, dtype=torch.float32):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError("%s must be a CUDA tensor (libpdgn_b200 has no CPU path)" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t


def knn_xyz(k, xyz, new_xyz=None, return_dist=False):
    """idx int32 [b,m,k] (and dist2 [b,m,k]) of the k nearest xyz[b] points for every new_xyz[b] point."""
    if new_xyz is None:
        new_xyz = xyz
    _req(xyz, "xyz"); _req(new_xyz, "new_xyz")
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = torch.empty((b, m, k), dtype=torch.int32, device=xyz.device)
    dist2 = torch.empty((b, m, k), dtype=torch.float32, device=xyz.device) if return_dist else None
    with torch.cuda.device(xyz.device):
        check(lib().pdgn_knn_xyz(xyz.data_ptr(), new_xyz.data_ptr(), b, n, m, int(k), idx.data_ptr(),
                                 dist2.data_ptr() if return_dist else None, _stream(xyz)), "pdgn_knn_xyz")
    return (idx, dist2) if return_dist else idx


def nn3(unknown, known):
    """(dist2 [b,n,3] SQUARED, idx int32 [b,n,3])."""
    _req(unknown, "unknown"); _req(known, "known")
    b, n, _ = unknown.shape
    m = known.shape[1]
    dist2 = torch.empty((b, n, 3), dtype=torch.float32, device=unknown.device)
    idx = torch.empty((b, n, 3), dtype=torch.int32, device=unknown.device)
    with torch.cuda.device(unknown.device):
        check(lib().pdgn_nn3(unknown.data_ptr(), known.data_ptr(), b, n, m, dist2.data_ptr(), idx.data_ptr(), _stream(unknown)), "pdgn_nn3")
    return dist2, idx


def group_fwd(features, idx):
    _req(features, "features"); _req(idx, "idx", torch.int32)
    b, c, n = features.shape
    _, m, k = idx.shape
    out = torch.empty((b, c, m, k), dtype=torch.float32, device=features.device)
    with torch.cuda.device(features.device):
        check(lib().pdgn_group_fwd(features.data_ptr(), idx.data_ptr(), b, c, n, m, k, out.data_ptr(), _stream(features)), "pdgn_group_fwd")
    return out


def group_bwd(grad_out, idx, n):
    _req(grad_out, "grad_out"); _req(idx, "idx", torch.int32)
    b, c, m, k = grad_out.shape
    grad = torch.zeros((b, c, n), dtype=torch.float32, device=grad_out.device)
    L = lib()
    ws_bytes = L.pdgn_group_bwd_workspace(b, n, m, k)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        check(L.pdgn_group_bwd_ws(grad_out.data_ptr(), idx.data_ptr(), b, c, n, m, k, grad.data_ptr(), ws.data_ptr(), ws_bytes,
                                  _stream(grad_out)), "pdgn_group_bwd_ws")
    return grad


def interp_fwd(features, idx, weight):
    _req(features, "features"); _req(idx, "idx", torch.int32); _req(weight, "weight")
    b, c, m = features.shape
    n = idx.shape[1]
    out = torch.empty((b, c, n), dtype=torch.float32, device=features.device)
    with torch.cuda.device(features.device):
        check(lib().pdgn_interp_fwd(features.data_ptr(), idx.data_ptr(), weight.data_ptr(), b, c, m, n, out.data_ptr(), _stream(features)), "pdgn_interp_fwd")
    return out


def interp_bwd(grad_out, idx, weight, m):
    _req(grad_out, "grad_out"); _req(idx, "idx", torch.int32); _req(weight, "weight")
    b, c, n = grad_out.shape
    grad = torch.zeros((b, c, m), dtype=torch.float32, device=grad_out.device)
    L = lib()
    ws_bytes = L.pdgn_interp_bwd_workspace(b, n, m)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        check(L.pdgn_interp_bwd_ws(grad_out.data_ptr(), idx.data_ptr(), weight.data_ptr(), b, c, n, m, grad.data_ptr(), ws.data_ptr(),
                                   ws_bytes, _stream(grad_out)), "pdgn_interp_bwd_ws")
    return grad


def chamfer_min(x, y, want_arg=True):
    """x [b,nx,d], y [b,ny,d] -> (min_xy [b,nx], arg_xy int32|None, min_yx [b,ny], arg_yx int32|None)."""
    _req(x, "x"); _req(y, "y")
    b, nx, d = x.shape
    ny = y.shape[1]
    dev = x.device
    mxy = torch.empty((b, nx), dtype=torch.float32, device=dev)
    myx = torch.empty((b, ny), dtype=torch.float32, device=dev)
    axy = torch.empty((b, nx), dtype=torch.int32, device=dev) if want_arg else None
    ayx = torch.empty((b, ny), dtype=torch.int32, device=dev) if want_arg else None
    with torch.cuda.device(dev):
        check(lib().pdgn_chamfer_min(x.data_ptr(), y.data_ptr(), b, nx, ny, d, mxy.data_ptr(), axy.data_ptr() if want_arg else None,
                                     myx.data_ptr(), ayx.data_ptr() if want_arg else None, _stream(x)), "pdgn_chamfer_min")
    return mxy, axy, myx, ayx


def chamfer_bwd(x, y, w_xy, arg_xy, w_yx, arg_yx):
    """Gradients (grad_x, grad_y) of sum(w_xy*min_xy) + sum(w_yx*min_yx)."""
    _req(x, "x"); _req(y, "y")
    b, nx, d = x.shape
    ny = y.shape[1]
    gx = torch.zeros_like(x)
    gy = torch.zeros_like(y)
    with torch.cuda.device(x.device):
        check(lib().pdgn_chamfer_bwd(x.data_ptr(), y.data_ptr(), b, nx, ny, d,
                                     _req(w_xy, "w_xy").data_ptr(), _req(arg_xy, "arg_xy", torch.int32).data_ptr(),
                                     _req(w_yx, "w_yx").data_ptr(), _req(arg_yx, "arg_yx", torch.int32).data_ptr(),
                                     gx.data_ptr(), gy.data_ptr(), _stream(x)), "pdgn_chamfer_bwd")
    return gx, gy


def cd_allpairs(A, B, rows=None, cols=None, out=None):
    """Chamfer matrix tile: A [na,npts,3], B [nb,npts,3] -> [rows1-rows0, cols1-cols0] (default: everything)."""
    _req(A, "A"); _req(B, "B")
    na, npts, _ = A.shape
    nb = B.shape[0]
    if B.shape[1] != npts:
        raise ValueError("all-pairs CD needs equal point counts (the reference's distChamfer does too)")
    r0, r1 = rows if rows is not None else (0, na)
    c0, c1 = cols if cols is not None else (0, nb)
    if out is None:
        out = torch.empty((r1 - r0, c1 - c0), dtype=torch.float32, device=A.device)
    L = lib()
    ws_bytes = L.pdgn_cd_allpairs_workspace(r1 - r0, c1 - c0, npts)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=A.device)
    with torch.cuda.device(A.device):
        check(L.pdgn_cd_allpairs(A.data_ptr(), B.data_ptr(), na, nb, npts, r0, r1, c0, c1, out.data_ptr(), out.stride(0),
                                 ws.data_ptr(), ws_bytes, _stream(A)), "pdgn_cd_allpairs")
    return out


def emd_allpairs(A, B, rows=None, cols=None, out=None):
    """Approximate-EMD matrix tile: A [na,n,3], B [nb,m,3] -> [rows, cols] of match_cost / n (no gradient)."""
    _req(A, "A"); _req(B, "B")
    na, n, _ = A.shape
    nb, m, _ = B.shape
    r0, r1 = rows if rows is not None else (0, na)
    c0, c1 = cols if cols is not None else (0, nb)
    if out is None:
        out = torch.empty((r1 - r0, c1 - c0), dtype=torch.float32, device=A.device)
    L = lib()
    ws_bytes = L.pdgn_emd_allpairs_workspace(r1 - r0, c1 - c0, n, m)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=A.device)
    with torch.cuda.device(A.device):
        check(L.pdgn_emd_allpairs(A.data_ptr(), B.data_ptr(), na, nb, n, m, r0, r1, c0, c1, out.data_ptr(), out.stride(0),
                                  ws.data_ptr(), ws_bytes, _stream(A)), "pdgn_emd_allpairs")
    return out


def emd_paired(A, B):
    """Approximate EMD of A[i] vs B[i]: A [b,n,3], B [b,m,3] -> [b] of match_cost / n (no gradient), one launch."""
    _req(A, "A"); _req(B, "B")
    b, n, _ = A.shape
    if B.shape[0] != b:
        raise ValueError("paired EMD needs equal batch sizes")
    m = B.shape[1]
    out = torch.empty((b,), dtype=torch.float32, device=A.device)
    L = lib()
    ws_bytes = L.pdgn_emd_paired_workspace(b, n, m)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=A.device)
    with torch.cuda.device(A.device):
        check(L.pdgn_emd_paired(A.data_ptr(), B.data_ptr(), b, n, m, out.data_ptr(), ws.data_ptr(), ws_bytes, _stream(A)), "pdgn_emd_paired")
    return out


def cd_allpairs_host(A, B, rows=None, cols=None):
    """Same, for CPU (ideally pinned) tensors: H2D + kernel + D2H inside the C call; returns a CPU tensor."""
    if A.is_cuda or B.is_cuda:
        raise TypeError("cd_allpairs_host takes host tensors")
    A = A.contiguous().float()
    B = B.contiguous().float()
    na, npts, _ = A.shape
    nb = B.shape[0]
    r0, r1 = rows if rows is not None else (0, na)
    c0, c1 = cols if cols is not None else (0, nb)
    out = torch.empty((r1 - r0, c1 - c0), dtype=torch.float32, pin_memory=True)
    check(lib().pdgn_cd_allpairs_host(A.data_ptr(), B.data_ptr(), na, nb, npts, r0, r1, c0, c1, out.data_ptr(), out.stride(0),
                                      torch.cuda.current_stream().cuda_stream), "pdgn_cd_allpairs_host")
    return out


def knn_feat(x, k, skip=1, return_dist=False):
    """x [b,c,n] -> idx int64 [b,n,k]: ranks skip..skip+k-1 of the (d2, index) order in feature space."""
    _req(x, "x")
    b, c, n = x.shape
    idx = torch.empty((b, n, k), dtype=torch.int64, device=x.device)
    dist2 = torch.empty((b, n, k), dtype=torch.float32, device=x.device) if return_dist else None
    L = lib()
    ws_bytes = L.pdgn_knn_feat_workspace(b, c, n)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        check(L.pdgn_knn_feat_ws(x.data_ptr(), b, c, n, int(k), int(skip), idx.data_ptr(),
                                 dist2.data_ptr() if return_dist else None, ws.data_ptr(), ws_bytes, _stream(x)), "pdgn_knn_feat_ws")
    return (idx, dist2) if return_dist else idx


def edge_feat_fwd(x, idx):
    _req(x, "x"); _req(idx, "idx", torch.int64)
    b, c, n = x.shape
    k = idx.shape[2]
    ee = torch.empty((b, 2 * c, n, k), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().pdgn_edge_feat_fwd(x.data_ptr(), idx.data_ptr(), b, c, n, k, ee.data_ptr(), _stream(x)), "pdgn_edge_feat_fwd")
    return ee


def edge_feat_bwd(grad_ee, idx, c):
    _req(grad_ee, "grad_ee"); _req(idx, "idx", torch.int64)
    b, _, n, k = grad_ee.shape
    gx = torch.zeros((b, c, n), dtype=torch.float32, device=grad_ee.device)
    L = lib()
    ws_bytes = L.pdgn_edge_feat_bwd_workspace(b, n, k)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=grad_ee.device)
    with torch.cuda.device(grad_ee.device):
        check(L.pdgn_edge_feat_bwd_ws(grad_ee.data_ptr(), idx.data_ptr(), b, c, n, k, gx.data_ptr(), ws.data_ptr(), ws_bytes,
                                      _stream(grad_ee)), "pdgn_edge_feat_bwd_ws")
    return gx


def local_stats_fwd(xyz, idx):
    """xyz [b,n,3], idx int32 [b,m,k] -> (mu [b,m,3], cov [b,m,9]) of every query's k neighbours."""
    _req(xyz, "xyz"); _req(idx, "idx", torch.int32)
    b, n, _ = xyz.shape
    _, m, k = idx.shape
    mu = torch.empty((b, m, 3), dtype=torch.float32, device=xyz.device)
    cov = torch.empty((b, m, 9), dtype=torch.float32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        check(lib().pdgn_local_stats_fwd(xyz.data_ptr(), idx.data_ptr(), b, n, m, k, mu.data_ptr(), cov.data_ptr(), _stream(xyz)), "pdgn_local_stats_fwd")
    return mu, cov


def local_stats_bwd(xyz, idx, mu, grad_mu, grad_cov):
    _req(xyz, "xyz"); _req(idx, "idx", torch.int32); _req(mu, "mu"); _req(grad_mu, "grad_mu"); _req(grad_cov, "grad_cov")
    b, n, _ = xyz.shape
    _, m, k = idx.shape
    gx = torch.zeros_like(xyz)
    with torch.cuda.device(xyz.device):
        check(lib().pdgn_local_stats_bwd(xyz.data_ptr(), idx.data_ptr(), mu.data_ptr(), grad_mu.data_ptr(), grad_cov.data_ptr(), b, n, m, k,
                                         gx.data_ptr(), _stream(xyz)), "pdgn_local_stats_bwd")
    return gx
