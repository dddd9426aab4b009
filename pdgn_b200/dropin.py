"""Drop-in installation: serve the reference's import names from this package so fpthink/PDGN's main.py and
models/PDGNet_v2.py run unchanged.

The reference reaches the hot path through four imports:
    models/PDGNet_v2.py:17   from lib.pointops.functions import pointops
    models/PDGNet_v2.py:22   from utils import chamfer_loss
    models/PDGNet_v2.py:21   from evaluation.evaluation_metrics import *
    lib/pointops/functions/pointops.py:7   import pointops_cuda          (only if the reference's own pointops is imported)
and through two module-level functions of models/PDGNet_v2.py (get_edge_features, get_edge_features_xyz :439-528).

install() pre-seeds sys.modules for the first group (the reference's `utils`, `lib`, `evaluation` packages stay
importable for everything else: utils.provider, utils.misc ...) and registers a post-import hook that rebinds the
two functions on models.PDGNet_v2 / models.PDGNet.  `python -m pdgn_b200.dropin /path/to/PDGN/main.py --phase test ...`
does install() and then runs the reference's main.py with runpy.
"""
import importlib
import importlib.abc
import importlib.util
import runpy
import sys
import types

from . import chamfer_loss, edge_features, evaluation_metrics, local_pair, pointops

_REBIND = ("models.PDGNet_v2", "models.PDGNet")


def make_pointops_cuda():
    """A `pointops_cuda` module with the reference's pybind signatures (pointops_api.cpp:16-39) for the six hot
    entry points, writing into caller-allocated tensors exactly as the reference's glue does."""
    import torch
    from ._lib import check, lib

    def st(t):
        return torch.cuda.current_stream(t.device).cuda_stream

    m = types.ModuleType("pointops_cuda")
    m.__doc__ = "pdgn_b200 replacement for the reference's pointops_cuda extension (hot-path subset)"

    def knnquery_cuda(b, n, m_, nsample, xyz, new_xyz, idx, dist2):
        check(lib().pdgn_knn_xyz(xyz.data_ptr(), new_xyz.data_ptr(), b, n, m_, nsample, idx.data_ptr(), dist2.data_ptr(), st(xyz)), "knnquery_cuda")

    def grouping_forward_cuda(b, c, n, m_, nsample, points, idx, out):
        check(lib().pdgn_group_fwd(points.data_ptr(), idx.data_ptr(), b, c, n, m_, nsample, out.data_ptr(), st(points)), "grouping_forward_cuda")

    def grouping_backward_cuda(b, c, n, m_, nsample, grad_out, idx, grad_points):
        check(lib().pdgn_group_bwd(grad_out.data_ptr(), idx.data_ptr(), b, c, n, m_, nsample, grad_points.data_ptr(), st(grad_out)), "grouping_backward_cuda")

    def nearestneighbor_cuda(b, n, m_, unknown, known, dist2, idx):
        check(lib().pdgn_nn3(unknown.data_ptr(), known.data_ptr(), b, n, m_, dist2.data_ptr(), idx.data_ptr(), st(unknown)), "nearestneighbor_cuda")

    def interpolation_forward_cuda(b, c, m_, n, points, idx, weight, out):
        check(lib().pdgn_interp_fwd(points.data_ptr(), idx.data_ptr(), weight.data_ptr(), b, c, m_, n, out.data_ptr(), st(points)), "interpolation_forward_cuda")

    def interpolation_backward_cuda(b, c, n, m_, grad_out, idx, weight, grad_points):
        check(lib().pdgn_interp_bwd(grad_out.data_ptr(), idx.data_ptr(), weight.data_ptr(), b, c, n, m_, grad_points.data_ptr(), st(grad_out)), "interpolation_backward_cuda")

    for f in (knnquery_cuda, grouping_forward_cuda, grouping_backward_cuda, nearestneighbor_cuda,
              interpolation_forward_cuda, interpolation_backward_cuda):
        setattr(m, f.__name__, f)
    return m


class _RebindFinder(importlib.abc.MetaPathFinder):
    """After models.PDGNet_v2 (or PDGNet) is executed, point its get_edge_features{,_xyz} at the CUDA versions."""

    def find_spec(self, name, path, target=None):
        if name not in _REBIND:
            return None
        sys.meta_path.remove(self)
        try:
            spec = importlib.util.find_spec(name)
        finally:
            sys.meta_path.insert(0, self)
        if spec is None or spec.loader is None:
            return spec
        inner = spec.loader
        # keep the module's real loader (get_source / get_code / is_package stay available to inspect, linecache,
        # traceback, coverage) and only extend its exec_module on this one instance
        run = inner.exec_module

        def exec_module(module, _run=run):
            _run(module)
            rebind(module)

        inner.exec_module = exec_module
        return spec


def rebind(module):
    module.get_edge_features = edge_features.get_edge_features
    module.get_edge_features_xyz = edge_features.get_edge_features_xyz
    # the trainer's get_local_pair (PDGNet_v2.py:136-155) -> the fused op; nsample is hard-wired to 20 there (:115, :144-145)
    for obj in list(vars(module).values()):
        if isinstance(obj, type) and "get_local_pair" in vars(obj) and "_reference_get_local_pair" not in vars(obj):
            obj._reference_get_local_pair = obj.get_local_pair      # the composed path stays reachable (tests, A/B timing)
            obj.get_local_pair = lambda self, pt1, pt2: local_pair.get_local_pair(pt1, pt2, 20)
    # the generator's forward notes its outputs, so that the six get_local_pair calls of a G step (PDGNet_v2.py:232-237) are
    # answered by ONE batched evaluation (pdgn_b200.local_pair.shape_losses); values and gradients are those of the six calls
    gen = vars(module).get("PointGenerator")
    if isinstance(gen, type) and "_reference_forward" not in vars(gen):
        gen._reference_forward = gen.forward

        def forward(self, *a, **kw):
            outs = gen._reference_forward(self, *a, **kw)
            if isinstance(outs, (tuple, list)):
                local_pair.note_generator_outputs(outs)
            return outs

        gen.forward = forward
    return module


_installed = False


def install():
    """Idempotent.  Call before importing the reference's models package."""
    global _installed
    if _installed:
        return
    sys.modules["lib.pointops.functions.pointops"] = pointops
    sys.modules["utils.chamfer_loss"] = chamfer_loss
    sys.modules["evaluation.evaluation_metrics"] = evaluation_metrics
    sys.modules.setdefault("pointops_cuda", make_pointops_cuda())
    # `from lib.pointops.functions import pointops` needs the parent packages; the reference ships them without
    # __init__.py (namespace packages), so only create parents that cannot be found on sys.path.
    for parent, child, mod in (("lib.pointops.functions", "pointops", pointops), ("utils", "chamfer_loss", chamfer_loss),
                               ("evaluation", "evaluation_metrics", evaluation_metrics)):
        try:
            pkg = importlib.import_module(parent)
        except ImportError:
            parts = parent.split(".")
            for i in range(1, len(parts) + 1):
                name = ".".join(parts[:i])
                if name not in sys.modules:
                    ns = types.ModuleType(name)
                    ns.__path__ = []
                    sys.modules[name] = ns
                    if i > 1:
                        setattr(sys.modules[".".join(parts[:i - 1])], parts[i - 1], ns)
            pkg = sys.modules[parent]
        setattr(pkg, child, mod)
    sys.meta_path.insert(0, _RebindFinder())
    _installed = True


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m pdgn_b200.dropin /path/to/PDGN/main.py [main.py args]")
    script = argv[0]
    import os
    sys.path.insert(0, os.path.dirname(os.path.abspath(script)))
    install()
    sys.argv = argv
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
