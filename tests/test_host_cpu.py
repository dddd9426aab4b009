"""CPU: host-side logic of the package that involves no kernel -- the MMD/COV/1-NNA reductions and the JSD arithmetic of
pdgn_b200.evaluation_metrics -- against the values the reference's own code produced (tests/golden)."""
import numpy as np
import pytest
import torch

from oracle import torch_ref as tref


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def test_mmd_cov_and_1nna_reductions_match_reference(golden):
    from pdgn_b200 import evaluation_metrics as em
    g = golden("evaluation_metrics")
    M_rs, M_rr, M_ss = T(g["all_cd"]), T(g["m_rr"]), T(g["m_ss"])
    mm = em.lgan_mmd_cov(M_rs.t())
    for key in [k for k in g.files if k.startswith("mmdcov:")]:
        assert mm[key[len("mmdcov:"):]].item() == pytest.approx(float(g[key]), rel=1e-7), key
    nn1 = em.knn(M_rr, M_rs, M_ss, 1, sqrt=False)
    for key in [k for k in g.files if k.startswith("knn:")]:
        assert nn1[key[len("knn:"):]].item() == pytest.approx(float(g[key]), rel=1e-7, abs=1e-12), key


@pytest.mark.parametrize("n0,n1,k,sqrt", [(7, 9, 3, True), (20, 20, 5, False), (4, 4, 1, True), (1, 6, 2, False)])
def test_knn_two_sample_test_equals_restated_reference(n0, n1, k, sqrt):
    from pdgn_b200 import evaluation_metrics as em
    rng = np.random.default_rng(n0 * 10 + n1)
    X = T(rng.random((n0, n0)).astype(np.float32))
    Y = T(rng.random((n1, n1)).astype(np.float32))
    XY = T(rng.random((n0, n1)).astype(np.float32))
    a, b = em.knn(X, XY, Y, k, sqrt=sqrt), tref.one_nn_accuracy(X, XY, Y, k, sqrt=sqrt)
    assert sorted(a) == sorted(b)
    for key in b:
        assert torch.equal(a[key], b[key]), key


def test_jsd_arithmetic_matches_reference(golden):
    from pdgn_b200 import evaluation_metrics as em
    g = golden("evaluation_metrics")
    jsd = em.jensen_shannon_divergence(g["jsd_counters_smp"], g["jsd_counters_ref"])
    assert jsd == pytest.approx(float(g["jsd_value"]), rel=1e-9)
    grid, spacing = em.unit_cube_grid_point_cloud(28, True)
    assert grid.shape == (len(g["jsd_counters_smp"]), 3) and grid.dtype == np.float32
    full, _ = em.unit_cube_grid_point_cloud(5, False)
    assert full.shape == (5, 5, 5, 3) and full[4, 0, 2].tolist() == [0.5, -0.5, 0.0]
    with pytest.raises(ValueError):
        em.jensen_shannon_divergence(np.array([1.0, -1.0]), np.array([1.0, 1.0]))


def test_jsd_nearest_cell_is_exactly_sklearns(golden, monkeypatch):
    """The cell assignment (analytic per-axis nearest in float64 + float64 re-rank of the sphere-boundary points) against
    sklearn's NearestNeighbors on the reference's float32 grid -- what evaluation_metrics.py:262-266 runs -- with the kNN
    kernel replaced by the CPU oracle (host logic only; the GPU test runs the real kernel against the golden counters)."""
    from sklearn.neighbors import NearestNeighbors
    from oracle import cpu as ocpu
    from pdgn_b200 import evaluation_metrics as em

    def oracle_knn(k, xyz, new_xyz):
        return torch.from_numpy(ocpu.knn_xyz(xyz.numpy(), new_xyz.numpy(), k)[0])

    monkeypatch.setattr(em.ops, "knn_xyz", oracle_knn)
    g = golden("evaluation_metrics")
    grid, _ = em.unit_cube_grid_point_cloud(28, True)
    nn = NearestNeighbors(n_neighbors=1).fit(grid)
    rng = np.random.default_rng(11)
    on_sphere = rng.standard_normal((20000, 3))
    on_sphere = (0.5 * on_sphere / np.linalg.norm(on_sphere, axis=1, keepdims=True)).astype(np.float32)
    for pts in (g["jsd_smp"].reshape(-1, 3), rng.uniform(-0.5, 0.5, (20000, 3)).astype(np.float32), on_sphere):
        idx, n_cells = em._nearest_grid_index(T(pts), 28, True)
        assert n_cells == len(grid)
        assert np.array_equal(idx.numpy(), nn.kneighbors(pts)[1][:, 0])
    idx, _ = em._nearest_grid_index(T(g["jsd_ref"].reshape(-1, 3)), 28, True)
    assert np.array_equal(np.bincount(idx.numpy(), minlength=len(grid)).astype(np.float64), g["jsd_counters_ref"])
