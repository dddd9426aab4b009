"""TEST INFRASTRUCTURE ONLY: run the reference's UNMODIFIED Python tree on the GPU box.

`make -C oracle refpy` stages fpthink/PDGN's .py files into baseline/_ref/PDGN (git-ignored, travels with gpurun; the
same place a `pip install --target baseline/_ref` would have put them had the reference been installable).  This module
loads that tree in two flavours, in ONE process and without the two touching each other:

  * load_reference()  -- "the reference on a B200": the reference's own Python (lib/pointops/functions/pointops.py,
    utils/chamfer_loss.py, evaluation/evaluation_metrics.py, models/PDGNet_v2.py) over the reference's own CUDA kernels
    recompiled for sm_100a (oracle/_ref/libpdgn_ref.so).  The two pybind extensions that do not build on torch 2.11
    (`pointops_cuda`: THC/THC.h is gone; `StructuralLossesBackend`) are replaced by ctypes shims with the SAME function
    names and argument lists (pointops_api.cpp:16-39, pybind/bind.cpp:10-16) over the same launchers.
  * load_dropin()     -- the same tree after pdgn_b200.dropin.install(): `import models.PDGNet_v2` as main.py does.

h5py (absent from this image, used only by the dataset readers) is stubbed with an empty module in both flavours.
"""
import contextlib
import ctypes
import importlib
import importlib.util
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TREE = os.path.join(ROOT, "baseline", "_ref", "PDGN")


def available():
    from . import ref_kernels
    return os.path.isfile(os.path.join(TREE, "models", "PDGNet_v2.py")) and ref_kernels.available()


def tree_available():
    return os.path.isfile(os.path.join(TREE, "models", "PDGNet_v2.py"))


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _cur(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def legacy_constructors():
    """The reference allocates with torch.cuda.IntTensor(b, m, k) / FloatTensor / LongTensor (pointops.py:425-426 ...).
    Where torch still accepts that spelling nothing is touched; otherwise equivalent factories are installed."""
    try:
        torch.cuda.IntTensor(1, 1)
        torch.cuda.FloatTensor(1, 1)
        return
    except Exception:
        pass

    def mk(dtype):
        def ctor(*sizes):
            return torch.empty(*sizes, dtype=dtype, device="cuda")
        return ctor
    torch.cuda.IntTensor = mk(torch.int32)
    torch.cuda.FloatTensor = mk(torch.float32)
    torch.cuda.LongTensor = mk(torch.int64)


def make_reference_pointops_cuda():
    """`pointops_cuda` (pointops_api.cpp:16-39) over the reference's own launchers.  The reference's glue passes the
    current stream to knnquery only; the other launchers use the legacy NULL stream (grouping_cuda_kernel.cu:85), which
    synchronises with torch's default stream by CUDA's legacy-stream rule."""
    from . import ref_kernels
    L = ref_kernels.lib()
    m = types.ModuleType("pointops_cuda")

    def knnquery_cuda(b, n, m_, nsample, xyz, new_xyz, idx, dist2):
        L.knnquery_cuda_launcher(b, n, m_, nsample, _p(xyz), _p(new_xyz), _p(idx), _p(dist2), _cur(xyz))

    def grouping_forward_cuda(b, c, n, m_, nsample, points, idx, out):
        L.grouping_forward_cuda_launcher_fast(b, c, n, m_, nsample, _p(points), _p(idx), _p(out))

    def grouping_backward_cuda(b, c, n, m_, nsample, grad_out, idx, grad_points):
        L.grouping_backward_cuda_launcher(b, c, n, m_, nsample, _p(grad_out), _p(idx), _p(grad_points))

    def nearestneighbor_cuda(b, n, m_, unknown, known, dist2, idx):
        L.nearestneighbor_cuda_launcher_fast(b, n, m_, _p(unknown), _p(known), _p(dist2), _p(idx))

    def interpolation_forward_cuda(b, c, m_, n, points, idx, weight, out):
        L.interpolation_forward_cuda_launcher_fast(b, c, m_, n, _p(points), _p(idx), _p(weight), _p(out))

    def interpolation_backward_cuda(b, c, n, m_, grad_out, idx, weight, grad_points):
        L.interpolation_backward_cuda_launcher(b, c, n, m_, _p(grad_out), _p(idx), _p(weight), _p(grad_points))

    for f in (knnquery_cuda, grouping_forward_cuda, grouping_backward_cuda, nearestneighbor_cuda,
              interpolation_forward_cuda, interpolation_backward_cuda):
        setattr(m, f.__name__, f)
    return m


def make_reference_structural_backend():
    """`StructuralLossesBackend` (pybind/bind.cpp:10-16; structural_loss.cpp:20-140): callee allocates and returns."""
    from . import ref_kernels
    L = ref_kernels.lib()
    m = types.ModuleType("StructuralLossesBackend")

    def _new(shape, like, dtype=torch.float32):
        return torch.empty(shape, dtype=dtype, device=like.device)

    def NNDistance(set_d, set_q):
        b, n, _ = set_d.shape
        q = set_q.shape[1]
        d1, i1 = _new((b, n), set_d), _new((b, n), set_d, torch.int32)
        d2, i2 = _new((b, q), set_d), _new((b, q), set_d, torch.int32)
        rc = L.ref_nndistance(b, n, _p(set_d), q, _p(set_q), _p(d1), _p(i1), _p(d2), _p(i2), _cur(set_d))
        assert rc == 0, rc
        return [d1, i1, d2, i2]

    def NNDistanceGrad(set_d, set_q, idx1, idx2, g1, g2):
        b, n, _ = set_d.shape
        q = set_q.shape[1]
        gx1, gx2 = _new((b, n, 3), set_d), _new((b, q, 3), set_d)
        g1, g2 = g1.contiguous(), g2.contiguous()
        # nndistance.cu:150-151 zero-fills with cudaMemset on the NULL stream: order it against torch's stream
        torch.cuda.current_stream(set_d.device).synchronize()
        rc = L.ref_nndistancegrad(b, n, _p(set_d), q, _p(set_q), _p(g1), _p(idx1), _p(g2), _p(idx2), _p(gx1), _p(gx2), _cur(set_d))
        assert rc == 0, rc
        torch.cuda.current_stream(set_d.device).synchronize()
        return [gx1, gx2]

    def ApproxMatch(set_d, set_q):
        b, n, _ = set_d.shape
        q = set_q.shape[1]
        match, temp = _new((b, q, n), set_d), _new((b, (q + n) * 2), set_d)
        rc = L.ref_approxmatch(b, n, q, _p(set_d), _p(set_q), _p(match), _p(temp), _cur(set_d))
        assert rc == 0, rc
        return [match, temp]

    def MatchCost(set_d, set_q, match):
        b, n, _ = set_d.shape
        q = set_q.shape[1]
        out = _new((b,), set_d)
        rc = L.ref_matchcost(b, n, q, _p(set_d), _p(set_q), _p(match), _p(out), _cur(set_d))
        assert rc == 0, rc
        return out

    def MatchCostGrad(set_d, set_q, match):
        b, n, _ = set_d.shape
        q = set_q.shape[1]
        g1, g2 = _new((b, n, 3), set_d), _new((b, q, 3), set_d)
        rc = L.ref_matchcostgrad(b, n, q, _p(set_d), _p(set_q), _p(match), _p(g1), _p(g2), _cur(set_d))
        assert rc == 0, rc
        return [g1, g2]

    for f in (NNDistance, NNDistanceGrad, ApproxMatch, MatchCost, MatchCostGrad):
        setattr(m, f.__name__, f)
    return m


@contextlib.contextmanager
def _swapped(mods):
    """Temporarily put `mods` (name -> module) into sys.modules; restore the previous entries afterwards."""
    missing = object()
    saved = {k: sys.modules.get(k, missing) for k in mods}
    sys.modules.update(mods)
    try:
        yield
    finally:
        for k, v in saved.items():
            if v is missing:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def _load(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(TREE, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _pkg(name):
    m = types.ModuleType(name)
    m.__path__ = []
    return m


def _paths():
    for p in (TREE, os.path.join(TREE, "utils")):
        if p not in sys.path:
            sys.path.append(p)


_ref = None


def load_reference():
    """Namespace with the reference's modules over the reference's kernels: .pointops, .chamfer_loss,
    .evaluation_metrics, .nn_distance, .match_cost, .model (models/PDGNet_v2.py)."""
    global _ref
    if _ref is not None:
        return _ref
    legacy_constructors()
    _paths()
    ns = types.SimpleNamespace()
    h5 = sys.modules.get("h5py") or types.ModuleType("h5py")
    backend = make_reference_structural_backend()
    with _swapped({"pointops_cuda": make_reference_pointops_cuda()}):
        ns.pointops = _load("_pdgn_ref.pointops", "lib/pointops/functions/pointops.py")
    ns.chamfer_loss = _load("_pdgn_ref.chamfer_loss", "utils/chamfer_loss.py")
    with _swapped({"metrics": _pkg("metrics"), "metrics.pytorch_structural_losses": _pkg("metrics.pytorch_structural_losses"),
                   "metrics.pytorch_structural_losses.StructuralLossesBackend": backend}):
        ns.nn_distance = _load("_pdgn_ref.nn_distance", "evaluation/pytorch_structural_losses/nn_distance.py")
        ns.match_cost = _load("_pdgn_ref.match_cost", "evaluation/pytorch_structural_losses/match_cost.py")
    # `make` in evaluation/pytorch_structural_losses copies the package to evaluation/StructuralLosses (Makefile:74-75)
    sl = _pkg("evaluation.StructuralLosses")
    with _swapped({"evaluation": _pkg("evaluation"), "evaluation.StructuralLosses": sl,
                   "evaluation.StructuralLosses.match_cost": ns.match_cost, "evaluation.StructuralLosses.nn_distance": ns.nn_distance}):
        ns.evaluation_metrics = _load("_pdgn_ref.evaluation_metrics", "evaluation/evaluation_metrics.py")
    lib_pkgs = {n: _pkg(n) for n in ("lib", "lib.pointops", "lib.pointops.functions")}
    lib_pkgs["lib.pointops.functions"].pointops = ns.pointops
    ev = _pkg("evaluation")
    ev.evaluation_metrics = ns.evaluation_metrics
    swap = dict(lib_pkgs)
    swap.update({"lib.pointops.functions.pointops": ns.pointops, "evaluation": ev,
                 "evaluation.evaluation_metrics": ns.evaluation_metrics, "h5py": h5})
    import utils as ref_utils  # the staged tree's utils package (utils.misc, utils.data are plain Python)
    saved_cl = sys.modules.get("utils.chamfer_loss"), getattr(ref_utils, "chamfer_loss", None)
    swap["utils.chamfer_loss"] = ns.chamfer_loss
    ref_utils.chamfer_loss = ns.chamfer_loss
    try:
        with _swapped(swap):
            ns.model = _load("_pdgn_ref.PDGNet_v2", "models/PDGNet_v2.py")
    finally:
        if saved_cl[1] is None:
            if hasattr(ref_utils, "chamfer_loss"):
                delattr(ref_utils, "chamfer_loss")
        else:
            ref_utils.chamfer_loss = saved_cl[1]
    _ref = ns
    return ns


def load_dropin():
    """models.PDGNet_v2 of the staged tree, imported the way main.py imports it, after pdgn_b200.dropin.install()."""
    _paths()
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    from pdgn_b200 import dropin
    dropin.install()
    return importlib.import_module("models.PDGNet_v2")


def bare_trainer(model_module, group, chamfer):
    """A PDGNet_v2 trainer object without its __init__ (which opens datasets and checkpoints): only what
    get_local_pair / compute_mean_covariance touch (PDGNet_v2.py:60,115,127-155)."""
    t = object.__new__(model_module.PDGNet_v2)
    t.group = group
    t.chamfer_loss = chamfer
    return t
