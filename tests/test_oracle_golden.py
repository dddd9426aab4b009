"""CPU: pin the oracle against vectors produced by the reference's own Python code (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import cpu as ocpu
from oracle import torch_ref as tref


@pytest.fixture(autouse=True)
def _one_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)  # the vectors were generated single-threaded (MKL bmm order depends on it)
    yield
    torch.set_num_threads(n)


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


@pytest.mark.parametrize("tag", ["d3", "d9", "d3sq"])
def test_chamfer_loss_restatement_matches_reference(golden, tag):
    g = golden("chamfer_loss")
    preds = T(g[tag + "_preds"]).requires_grad_(True)
    gts = T(g[tag + "_gts"]).requires_grad_(True)
    loss = tref.chamfer_loss(preds, gts)
    loss.backward()
    np.testing.assert_allclose(loss.detach().numpy(), g[tag + "_loss"], rtol=1e-6)
    np.testing.assert_allclose(preds.grad.numpy(), g[tag + "_gpreds"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gts.grad.numpy(), g[tag + "_ggts"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("tag,D", [("d3", 3), ("d9", 9), ("d3sq", 3)])
def test_c_oracle_chamfer_matches_reference_value(golden, tag, D):
    """Direct-difference C oracle vs the reference's Gram-form loss: 1e-5 relative on the scalar (north star)."""
    g = golden("chamfer_loss")
    preds, gts = g[tag + "_preds"], g[tag + "_gts"]
    m_gp, _ = ocpu.nn_min(gts, preds)
    m_pg, _ = ocpu.nn_min(preds, gts)
    loss = m_gp.astype(np.float64).sum() + m_pg.astype(np.float64).sum()
    assert abs(loss - float(g[tag + "_loss"])) <= 1e-5 * abs(float(g[tag + "_loss"]))


def test_dist_chamfer_restatement_and_c_oracle(golden):
    g = golden("evaluation_metrics")
    dl, dr = tref.dist_chamfer(T(g["a"]), T(g["b"]))
    np.testing.assert_array_equal(dl.numpy(), g["dl"])
    np.testing.assert_array_equal(dr.numpy(), g["dr"])
    # C oracle (direct form): per-point minima within Gram-form rounding of the reference
    m_ab, _ = ocpu.nn_min(g["a"], g["b"])
    m_ba, _ = ocpu.nn_min(g["b"], g["a"])
    np.testing.assert_allclose(m_ba, g["dl"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(m_ab, g["dr"], rtol=1e-4, atol=2e-6)


def test_pairwise_cd_restatement_and_c_oracle(golden):
    g = golden("evaluation_metrics")
    smp, ref = T(g["smp"]), T(g["ref"])
    np.testing.assert_array_equal(tref.pairwise_cd(smp, ref, 4).numpy(), g["all_cd"])
    np.testing.assert_array_equal(tref.pairwise_cd(ref, ref, 4).numpy(), g["m_rr"])
    np.testing.assert_allclose(ocpu.cd_allpairs(g["smp"], g["ref"]), g["all_cd"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(ocpu.cd_allpairs(g["smp"], g["smp"]), g["m_ss"], rtol=1e-5, atol=2e-7)


def test_metrics_restatement(golden):
    g = golden("evaluation_metrics")
    M_rs, M_rr, M_ss = T(g["all_cd"]), T(g["m_rr"]), T(g["m_ss"])
    res = tref.compute_all_metrics_cd(None, None, None, matrices=(M_rs, M_rr, M_ss))
    for key in [k for k in g.files if k.startswith("metric:")]:
        assert res[key[len("metric:"):]].item() == pytest.approx(float(g[key]), rel=1e-7, abs=1e-12), key
    mm = tref.lgan_mmd_cov(M_rs.t())
    for key in [k for k in g.files if k.startswith("mmdcov:")]:
        assert mm[key[len("mmdcov:"):]].item() == pytest.approx(float(g[key]), rel=1e-7), key
    nn1 = tref.one_nn_accuracy(M_rr, M_rs, M_ss, 1)
    for key in [k for k in g.files if k.startswith("knn:")]:
        assert nn1[key[len("knn:"):]].item() == pytest.approx(float(g[key]), rel=1e-7, abs=1e-12), key
    # the full pipeline from clouds, through the restated pairwise stage
    res2 = tref.compute_all_metrics_cd(T(g["smp"]), T(g["ref"]), 4)
    for key in [k for k in g.files if k.startswith("metric:")]:
        assert res2[key[len("metric:"):]].item() == pytest.approx(float(g[key]), rel=1e-6, abs=1e-12), key


def test_edge_features_restatement(golden):
    g = golden("edge_features")
    x, pc = T(g["x"]), T(g["pc"])
    ee, idx = tref.get_edge_features(x, 10)
    np.testing.assert_array_equal(idx.numpy(), g["idx"])
    np.testing.assert_array_equal(ee.numpy(), g["ee"])
    e_fea, e_xyz, idx2 = tref.get_edge_features_xyz(x, pc, 10)
    np.testing.assert_array_equal(idx2.numpy(), g["idx_xyz"])
    np.testing.assert_array_equal(e_fea.numpy(), g["e_fea"])
    np.testing.assert_array_equal(e_xyz.numpy(), g["e_xyz"])


def _tie_tolerant_equal(idx_a, idx_b, dist_of, rel=1e-5):
    """Index lists must be identical except where the candidates' distances differ by less than `rel`."""
    bad = np.argwhere(idx_a != idx_b)
    for pos in bad:
        pos = tuple(pos)
        da, db = dist_of(pos, idx_a[pos]), dist_of(pos, idx_b[pos])
        assert abs(da - db) <= rel * max(abs(da), abs(db), 1e-12), (pos, idx_a[pos], idx_b[pos], da, db)
    return len(bad)


def test_c_oracle_knn_feat_matches_reference_indices(golden):
    """The exact-FP32 (d2, index) contract vs the reference's Gram + torch.sort pick (tie tolerant)."""
    g = golden("edge_features")
    x = g["x"]
    idx, _ = ocpu.knn_feat(x, 10, skip=1)
    x64 = x.astype(np.float64)

    def dist_of(pos, j):
        b, i, _ = pos
        return float(((x64[b, :, i] - x64[b, :, j]) ** 2).sum())

    nbad = _tie_tolerant_equal(idx, g["idx"], dist_of)
    assert nbad <= idx.size // 100


def test_c_oracle_knn_xyz_matches_reference_naive(golden):
    """fmaf-chain insertion oracle vs the authors' KNNQueryNaive / KNNQueryExclude (pointops.py:368-474)."""
    g = golden("knn_naive")
    xyz, new_xyz = g["xyz"], g["new_xyz"]
    x64 = xyz.astype(np.float64)

    def mk(q64):
        def dist_of(pos, j):
            b, i, _ = pos
            return float(((q64[b, i] - x64[b, j]) ** 2).sum())
        return dist_of

    idx, d2 = ocpu.knn_xyz(xyz, new_xyz, 20)
    assert _tie_tolerant_equal(idx, g["idx_naive"], mk(new_xyz.astype(np.float64))) <= 2
    assert np.all(np.diff(d2, axis=-1) >= 0)
    idx_self, _ = ocpu.knn_xyz(xyz, xyz, 9)
    assert _tie_tolerant_equal(idx_self[:, :, :8], g["idx_self"], mk(x64)) <= 2
    assert _tie_tolerant_equal(idx_self[:, :, 1:9], g["idx_excl"], mk(x64)) <= 2
    np.testing.assert_array_equal(tref.knnquery_naive(20, T(xyz), T(new_xyz)).numpy(), g["idx_naive"])


def test_c_oracle_knn_edge_cases():
    rng = np.random.default_rng(3)
    xyz = rng.uniform(-1, 1, (2, 5, 3)).astype(np.float32)
    idx, d2 = ocpu.knn_xyz(xyz, xyz, 8)  # n < k: trailing idx 0 / dist +inf (knnquery_cuda_kernel.cu:23-26)
    assert np.all(idx[:, :, 5:] == 0) and np.all(np.isinf(d2[:, :, 5:]))
    assert np.all(idx[:, :, 0] == np.arange(5)[None, :])
    dup = np.repeat(xyz[:, :1], 6, axis=1)  # all points identical: ties resolve to ascending index
    idx, d2 = ocpu.knn_xyz(dup, dup, 4)
    assert np.all(idx == np.arange(4)[None, None, :]) and np.all(d2 == 0)
    d3, i3 = ocpu.nn3(dup[:, :2], dup)
    assert np.all(i3 == np.arange(3)[None, None, :])
    bad = xyz.copy()
    bad[0, 2] = np.nan  # NaN distances are never selected
    idx, _ = ocpu.knn_xyz(bad, xyz, 5)
    assert not np.any(idx[0, :, :4] == 2)


def test_c_oracle_gathers():
    rng = np.random.default_rng(4)
    pts = rng.standard_normal((2, 5, 17)).astype(np.float32)
    idx = rng.integers(0, 17, (2, 9, 4)).astype(np.int32)
    out = ocpu.group_fwd(pts, idx)
    ref = np.stack([pts[b][:, idx[b]] for b in range(2)])
    np.testing.assert_array_equal(out, ref)
    g = rng.standard_normal(out.shape).astype(np.float32)
    gp = ocpu.group_bwd(g, idx, 17)
    exp = np.zeros((2, 5, 17))
    for b in range(2):
        for j in range(9):
            for s in range(4):
                exp[b, :, idx[b, j, s]] += g[b, :, j, s]
    np.testing.assert_allclose(gp, exp, rtol=1e-6, atol=1e-6)
    idx3 = rng.integers(0, 17, (2, 11, 3)).astype(np.int32)
    w = rng.uniform(0, 1, (2, 11, 3)).astype(np.float32)
    o = ocpu.interp_fwd(pts, idx3, w)
    exp = np.stack([(pts[b][:, idx3[b]].astype(np.float64) * w[b][None]).sum(-1) for b in range(2)])
    np.testing.assert_allclose(o, exp, rtol=1e-6, atol=1e-6)
    go = rng.standard_normal(o.shape).astype(np.float32)
    gi = ocpu.interp_bwd(go, idx3, w, 17)
    exp = np.zeros((2, 5, 17))
    for b in range(2):
        for j in range(11):
            for t in range(3):
                exp[b, :, idx3[b, j, t]] += go[b, :, j].astype(np.float64) * w[b, j, t]
    np.testing.assert_allclose(gi, exp, rtol=1e-5, atol=1e-6)
