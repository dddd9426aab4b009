"""CPU: the C-ABI library loads, exports everything include/pdgn_b200.h declares, and the product path has no
fallback (no oracle import, loud failure without the library / on CPU tensors)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from pdgn_b200 import _build
    return _build.build()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "pdgn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pdgn_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = header_symbols()
    for name in ["pdgn_knn_xyz", "pdgn_nn3", "pdgn_group_fwd", "pdgn_group_bwd", "pdgn_interp_fwd", "pdgn_interp_bwd",
                 "pdgn_chamfer_min", "pdgn_chamfer_bwd", "pdgn_cd_allpairs", "pdgn_cd_allpairs_workspace",
                 "pdgn_cd_allpairs_host", "pdgn_knn_feat", "pdgn_edge_feat_fwd", "pdgn_edge_feat_bwd", "pdgn_group_bwd_ws",
                 "pdgn_interp_bwd_ws", "pdgn_edge_feat_bwd_ws", "pdgn_knn_feat_ws", "pdgn_knn_feat_workspace"]:
        assert name in syms


def test_library_exports_every_declared_symbol(built):
    L = ctypes.CDLL(built)
    for name in header_symbols():
        assert hasattr(L, name), name
    from pdgn_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_symbols()
    assert _lib.lib().pdgn_abi_version() == 1
    assert b"bad argument" in _lib.lib().pdgn_error_string(-1)


def test_library_is_compiled_for_sm100a_with_tma(built):
    out = subprocess.run(["cuobjdump", "-lelf", built], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    full = subprocess.run(["cuobjdump", "-sass", built], capture_output=True, text=True).stdout
    parts = [p for p in full.split("Function : ") if "cd_allpairs_kernel" in p.split("\n")[0]]
    assert parts, "all-pairs kernel not found in the library"
    sass = parts[0]
    assert "UBLKCP" in sass       # TMA bulk copy feeds the candidate tiles
    assert "FMNMX3" in sass       # 3-input min (sm_100)
    assert "REDUX" in sass        # warp-level column-min reduction


def test_argument_errors_return_codes_not_crashes(built):
    from pdgn_b200 import _lib
    L = _lib.lib()
    assert L.pdgn_knn_xyz(None, None, 1, 1, 1, 1, None, None, None) == -1
    assert L.pdgn_cd_allpairs(None, None, 1, 1, 1, 0, 1, 0, 1, None, 1, None, 0, None) == -1
    assert L.pdgn_cd_allpairs_workspace(1000, 1000, 2048) >= 2000 * 3 * 2048 * 4
    with pytest.raises(_lib.PdgnError):
        _lib.check(-2, "x")


def test_product_code_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "pdgn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and "torch_ref" not in text, f


def test_ops_refuse_cpu_tensors(built):
    from pdgn_b200 import ops
    x = torch.zeros(1, 8, 3)
    with pytest.raises(TypeError):
        ops.knn_xyz(2, x)
    with pytest.raises(TypeError):
        ops.cd_allpairs(x, x)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from pdgn_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "SO_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.PdgnError, match="no CPU or PyTorch fallback"):
        _lib.lib()


def test_modules_mirror_the_reference_names(built):
    from pdgn_b200 import chamfer_loss, edge_features, evaluation_metrics, pointops
    for name in ["knnquery", "grouping", "nearestneighbor", "interpolation", "KNNQuery", "Grouping", "NearestNeighbor",
                 "Interpolation", "Gen_QueryAndGroupXYZ", "QueryAndGroup", "GroupAll", "knnquery_naive", "knnquery_exclude"]:
        assert hasattr(pointops, name), name
    assert hasattr(chamfer_loss, "ChamferLoss")
    for name in ["distChamfer", "distChamferCUDA", "emd_approx", "EMD_CD", "_pairwise_EMD_CD_", "knn", "lgan_mmd_cov",
                 "compute_all_metrics", "unit_cube_grid_point_cloud", "jsd_between_point_cloud_sets", "entropy_of_occupancy_grid",
                 "jensen_shannon_divergence"]:  # every def of the reference's evaluation_metrics.py
        assert hasattr(evaluation_metrics, name), name
    for name in ["get_edge_features", "get_edge_features_xyz"]:
        assert hasattr(edge_features, name), name


def test_workspace_queries_need_no_gpu(built):
    """The *_workspace entry points are pure host arithmetic (callable without a device): positive for sane shapes, growing with
    the problem, 0 for nonsense; pdgn_knn_feat_workspace covers the tensor-core path's centred tiled copy (channels padded to 32),
    the transposed copy, partial norms, 64-entry candidate lists and counters."""
    from pdgn_b200 import _lib
    L = _lib.lib()
    small, big = L.pdgn_knn_feat_workspace(2, 40, 640), L.pdgn_knn_feat_workspace(35, 256, 1024)
    assert 0 < small < big
    n_pad = 2 * 64 * 640 * 4                      # centred copy: 40 channels padded to 64
    assert small >= n_pad + 2 * 40 * 640 * 4 + 2 * 640 * 64 * 4
    assert L.pdgn_knn_feat_workspace(-1, 8, 128) == 0
    assert L.pdgn_group_bwd_workspace(35, 1024, 1024, 10) > 35 * 1024 * 10 * 4
    assert L.pdgn_interp_bwd_workspace(35, 2048, 1024) > 35 * 2048 * 3 * 4
    assert L.pdgn_edge_feat_bwd_workspace(35, 1024, 10) > 35 * 1024 * 10 * 4
