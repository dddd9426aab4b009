"""ctypes loader for libpdgn_b200.so -- the only native dependency of the package.

There is deliberately no fallback: if the library is missing or an entry point fails, the caller gets an
exception (the reference calls exit(-1) from inside its launchers, knnquery_cuda_kernel.cu:66-70; we raise).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libpdgn_b200.so")
_lib = None

_P = ctypes.c_void_p
_I = ctypes.c_int
_LL = ctypes.c_longlong
_SZ = ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/pdgn_b200.h
SIGNATURES = {
    "pdgn_abi_version": (_I, []),
    "pdgn_error_string": (ctypes.c_char_p, [_I]),
    "pdgn_knn_xyz": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "pdgn_nn3": (_I, [_P, _P, _I, _I, _I, _P, _P, _P]),
    "pdgn_group_fwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "pdgn_group_bwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "pdgn_group_bwd_workspace": (_SZ, [_I, _I, _I, _I]),
    "pdgn_group_bwd_ws": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _SZ, _P]),
    "pdgn_interp_fwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "pdgn_interp_bwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "pdgn_interp_bwd_workspace": (_SZ, [_I, _I, _I]),
    "pdgn_interp_bwd_ws": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _SZ, _P]),
    "pdgn_chamfer_min": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "pdgn_chamfer_bwd": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "pdgn_cd_allpairs_workspace": (_SZ, [_I, _I, _I]),
    "pdgn_cd_allpairs": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _LL, _P, _SZ, _P]),
    "pdgn_cd_allpairs_host": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _LL, _P]),
    "pdgn_emd_allpairs_workspace": (_SZ, [_I, _I, _I, _I]),
    "pdgn_emd_allpairs": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _LL, _P, _SZ, _P]),
    "pdgn_emd_paired_workspace": (_SZ, [_I, _I, _I]),
    "pdgn_emd_paired": (_I, [_P, _P, _I, _I, _I, _P, _P, _SZ, _P]),
    "pdgn_local_stats_fwd": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "pdgn_local_stats_bwd": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "pdgn_local_pair_workspace": (_SZ, [_I, _I, _I, _I]),
    "pdgn_local_pair_fwd": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _SZ, _P]),
    "pdgn_local_pair_bwd": (_I, [_I, _I, _I, _I, _P, _P, _P, _P, _SZ, _P]),
    "pdgn_shape_loss_workspace": (_SZ, [_I, _I, _P, _I]),
    "pdgn_shape_loss_fwd": (_I, [_P, _I, _I, _P, _I, _P, _P, _SZ, _P]),
    "pdgn_shape_loss_bwd": (_I, [_I, _I, _P, _I, _P, _P, _P, _SZ, _P]),
    "pdgn_knn_feat": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _P]),
    "pdgn_knn_feat_workspace": (_SZ, [_I, _I, _I]),
    "pdgn_knn_feat_ws": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _P, _SZ, _P]),
    "pdgn_edge_feat_fwd": (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "pdgn_edge_feat_bwd": (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "pdgn_edge_feat_bwd_workspace": (_SZ, [_I, _I, _I]),
    "pdgn_edge_feat_bwd_ws": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _SZ, _P]),
}


class PdgnError(RuntimeError):
    pass


def lib():
    """Load the shared library once; raise loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise PdgnError(
                "libpdgn_b200.so not found at %s -- build it with `python -m pdgn_b200._build` "
                "(there is no CPU or PyTorch fallback for these ops)" % SO_PATH)
        L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here means header and library disagree
            fn.restype = res
            fn.argtypes = args
        if L.pdgn_abi_version() != 1:
            raise PdgnError("libpdgn_b200.so ABI version mismatch")
        _lib = L
    return _lib


def check(code, what):
    if code != 0:
        msg = lib().pdgn_error_string(code)
        raise PdgnError("%s failed: %s (code %d)" % (what, msg.decode() if msg else "?", code))
