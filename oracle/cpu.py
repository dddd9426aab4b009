"""ctypes binding of liboracle.so (oracle/pdgn_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force=False):
    """Compile pdgn_oracle.c with gcc (seconds)."""
    src = os.path.join(_HERE, "pdgn_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", _SO, src, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _out(shape, dtype):
    a = np.empty(shape, dtype=dtype)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def knn_xyz(xyz, new_xyz, k):
    """(idx int32 [b,m,k], dist2 f32 [b,m,k]) -- knnquery_cuda_kernel.cu:6-50."""
    xyz, px = _f(xyz)
    new_xyz, pq = _f(new_xyz)
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    assert k <= 256
    idx, pi = _out((b, m, k), np.int32)
    d2, pd = _out((b, m, k), np.float32)
    lib().oracle_knn_xyz(px, pq, b, n, m, k, pi, pd)
    return idx, d2


def nn3(unknown, known):
    """(dist2 f32 [b,n,3] SQUARED, idx int32 [b,n,3]) -- interpolation_cuda_kernel.cu:134-176."""
    unknown, pu = _f(unknown)
    known, pk = _f(known)
    b, n, _ = unknown.shape
    m = known.shape[1]
    d2, pd = _out((b, n, 3), np.float32)
    idx, pi = _out((b, n, 3), np.int32)
    lib().oracle_nn3(pu, pk, b, n, m, pd, pi)
    return d2, idx


def group_fwd(points, idx):
    points, pp = _f(points)
    idx, pi = _i(idx)
    b, c, n = points.shape
    _, m, k = idx.shape
    out, po = _out((b, c, m, k), np.float32)
    lib().oracle_group_fwd(pp, pi, b, c, n, m, k, po)
    return out


def group_bwd(grad_out, idx, n):
    grad_out, pg = _f(grad_out)
    idx, pi = _i(idx)
    b, c, m, k = grad_out.shape
    gp = np.zeros((b, c, n), dtype=np.float32)
    lib().oracle_group_bwd(pg, pi, b, c, n, m, k, gp.ctypes.data_as(ctypes.c_void_p))
    return gp


def interp_fwd(points, idx, weight):
    points, pp = _f(points)
    idx, pi = _i(idx)
    weight, pw = _f(weight)
    b, c, m = points.shape
    n = idx.shape[1]
    out, po = _out((b, c, n), np.float32)
    lib().oracle_interp_fwd(pp, pi, pw, b, c, m, n, po)
    return out


def interp_bwd(grad_out, idx, weight, m):
    grad_out, pg = _f(grad_out)
    idx, pi = _i(idx)
    weight, pw = _f(weight)
    b, c, n = grad_out.shape
    gp = np.zeros((b, c, m), dtype=np.float32)
    lib().oracle_interp_bwd(pg, pi, pw, b, c, n, m, gp.ctypes.data_as(ctypes.c_void_p))
    return gp


def nn_min(x, y):
    """Directional min squared distance + argmin, x [b,nx,D] against y [b,ny,D]."""
    x, px = _f(x)
    y, py = _f(y)
    b, nx, D = x.shape
    ny = y.shape[1]
    mind, pm = _out((b, nx), np.float32)
    arg, pa = _out((b, nx), np.int32)
    lib().oracle_nn_min(px, py, b, nx, ny, D, pm, pa)
    return mind, arg


def nndistance(xyz1, xyz2):
    """(dist1, idx1, dist2, idx2) -- nndistance.cu:2-128."""
    d1, i1 = nn_min(xyz1, xyz2)
    d2, i2 = nn_min(xyz2, xyz1)
    return d1, i1, d2, i2


def cd_allpairs(A, B):
    """[na, nb] Chamfer matrix (mean+mean of direct-difference min squared distances)."""
    A, pa = _f(A)
    B, pb = _f(B)
    na, npts, _ = A.shape
    nb = B.shape[0]
    assert B.shape[1] == npts
    out, po = _out((na, nb), np.float32)
    lib().oracle_cd_allpairs(pa, pb, na, nb, npts, po)
    return out


def knn_feat(x, k, skip=1):
    """(idx int64 [b,n,k], dist2 f32) for x [b,c,n]: exact FP32 direct distances, (d2, idx) order."""
    x, px = _f(x)
    b, c, n = x.shape
    assert skip + k <= n
    idx, pi = _out((b, n, k), np.int64)
    d2, pd = _out((b, n, k), np.float32)
    lib().oracle_knn_feat(px, b, c, n, k, skip, pi, pd)
    return idx, d2


def emd_cost(xyz1, xyz2):
    """Approximate-EMD matching cost per pair, xyz1 [b,n,3] vs xyz2 [b,m,3] -> [b] (approxmatch.cu + matchcost)."""
    xyz1, p1 = _f(xyz1)
    xyz2, p2 = _f(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    out, po = _out((b,), np.float32)
    lib().oracle_emd_cost(p1, p2, b, n, m, po)
    return out


def emd_allpairs(A, B):
    """all_emd[s, r] = match_cost(A_s, B_r) / npts (evaluation_metrics.py:26-31, :110)."""
    A = np.ascontiguousarray(A, dtype=np.float32)
    B = np.ascontiguousarray(B, dtype=np.float32)
    na, nb, npts = A.shape[0], B.shape[0], A.shape[1]
    x1 = np.repeat(A, nb, axis=0)
    x2 = np.tile(B, (na, 1, 1))
    return (emd_cost(x1, x2) / np.float32(npts)).reshape(na, nb)
