"""ctypes binding of oracle/_ref/libpdgn_ref.so: the REFERENCE's own CUDA kernels (lib/pointops/src/*/..._kernel.cu,
evaluation/pytorch_structural_losses/src/nndistance.cu) compiled unmodified for sm_100a by oracle/Makefile.
TEST INFRASTRUCTURE ONLY; needs a GPU.  Note the reference launches grouping / interpolation / 3-NN on the legacy
NULL stream (grouping_cuda_kernel.cu:85): callers synchronise around these calls."""
import ctypes
import os

import torch

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libpdgn_ref.so")
_lib = None


def available():
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def knnquery(k, xyz, new_xyz):
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = torch.zeros((b, m, k), dtype=torch.int32, device=xyz.device)
    d2 = torch.zeros((b, m, k), dtype=torch.float32, device=xyz.device)
    torch.cuda.synchronize()
    lib().knnquery_cuda_launcher(b, n, m, k, _p(xyz), _p(new_xyz), _p(idx), _p(d2), ctypes.c_void_p(0))
    torch.cuda.synchronize()
    return idx, d2


def nn3(unknown, known):
    b, n, _ = unknown.shape
    m = known.shape[1]
    d2 = torch.zeros((b, n, 3), dtype=torch.float32, device=unknown.device)
    idx = torch.zeros((b, n, 3), dtype=torch.int32, device=unknown.device)
    torch.cuda.synchronize()
    lib().nearestneighbor_cuda_launcher_fast(b, n, m, _p(unknown), _p(known), _p(d2), _p(idx))
    torch.cuda.synchronize()
    return d2, idx


def group_fwd(points, idx):
    b, c, n = points.shape
    _, m, k = idx.shape
    out = torch.zeros((b, c, m, k), dtype=torch.float32, device=points.device)
    torch.cuda.synchronize()
    lib().grouping_forward_cuda_launcher_fast(b, c, n, m, k, _p(points), _p(idx), _p(out))
    torch.cuda.synchronize()
    return out


def group_bwd(grad_out, idx, n):
    b, c, m, k = grad_out.shape
    g = torch.zeros((b, c, n), dtype=torch.float32, device=grad_out.device)
    torch.cuda.synchronize()
    lib().grouping_backward_cuda_launcher(b, c, n, m, k, _p(grad_out), _p(idx), _p(g))
    torch.cuda.synchronize()
    return g


def interp_fwd(points, idx, weight):
    b, c, m = points.shape
    n = idx.shape[1]
    out = torch.zeros((b, c, n), dtype=torch.float32, device=points.device)
    torch.cuda.synchronize()
    lib().interpolation_forward_cuda_launcher_fast(b, c, m, n, _p(points), _p(idx), _p(weight), _p(out))
    torch.cuda.synchronize()
    return out


def interp_bwd(grad_out, idx, weight, m):
    b, c, n = grad_out.shape
    g = torch.zeros((b, c, m), dtype=torch.float32, device=grad_out.device)
    torch.cuda.synchronize()
    lib().interpolation_backward_cuda_launcher(b, c, n, m, _p(grad_out), _p(idx), _p(weight), _p(g))
    torch.cuda.synchronize()
    return g


def nndistance(xyz1, xyz2):
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dev = xyz1.device
    d1 = torch.zeros((b, n), dtype=torch.float32, device=dev)
    i1 = torch.zeros((b, n), dtype=torch.int32, device=dev)
    d2 = torch.zeros((b, m), dtype=torch.float32, device=dev)
    i2 = torch.zeros((b, m), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    rc = lib().ref_nndistance(b, n, _p(xyz1), m, _p(xyz2), _p(d1), _p(i1), _p(d2), _p(i2), ctypes.c_void_p(0))
    torch.cuda.synchronize()
    assert rc == 0, rc
    return d1, i1, d2, i2


def match_cost(xyz1, xyz2):
    """[b] approximate-EMD matching cost through the reference's approxmatch + matchcost kernels."""
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dev = xyz1.device
    match = torch.zeros((b, m, n), dtype=torch.float32, device=dev)
    temp = torch.zeros((b, (n + m) * 2), dtype=torch.float32, device=dev)
    out = torch.zeros((b,), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    rc = lib().ref_match_cost(b, n, m, _p(xyz1), _p(xyz2), _p(match), _p(temp), _p(out), ctypes.c_void_p(0))
    torch.cuda.synchronize()
    assert rc == 0, rc
    return out
