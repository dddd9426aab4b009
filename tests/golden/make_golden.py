"""Generate tests/golden/*.npz by running the REFERENCE's own Python code (fpthink/PDGN, mounted read-only
at /root/reference) on CPU in the build container.  The reference cannot travel to the GPU box, so the
vectors are committed; re-run with `python tests/golden/make_golden.py` to regenerate (bit-stable on the
same torch build).

How each piece of the reference is reached:
  * utils/chamfer_loss.py imports as is.
  * evaluation/evaluation_metrics.py cannot be imported (needs the un-built StructuralLosses extension):
    distChamfer, _pairwise_EMD_CD_, knn, lgan_mmd_cov, compute_all_metrics are ast-extracted and exec'd
    unmodified; `emd_approx` is stubbed to zeros (EMD is out of scope), so only the -CD keys are kept.
  * models/PDGNet_v2.py cannot be imported (h5py, pointops_cuda): get_edge_features{,_xyz} are
    ast-extracted.  torch.sort is wrapped to also record the indices the reference picked.
  * lib/pointops/functions/pointops.py imports with a stub `pointops_cuda` module; KNNQueryNaive /
    KNNQueryExclude are the authors' own pure-torch statement of knnquery.
"""
import ast
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("PDGN_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def extract(path, names, ns):
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing
    return ns


def clouds(gen, *shape, kind="U"):
    if kind == "U":
        return (torch.rand(*shape, generator=gen) * 2 - 1).float()
    v = torch.randn(*shape, generator=gen)
    v = v / v.norm(dim=-1, keepdim=True)
    return (v * (1 + 0.01 * torch.randn(*shape[:-1], 1, generator=gen))).float()


def main():
    torch.set_num_threads(1)  # summation order of MKL bmm is thread-count dependent; pin it
    g = torch.Generator().manual_seed(0)
    sys.path.insert(0, REF)

    # ---- ChamferLoss (utils/chamfer_loss.py:7-38) ---------------------------------------------------
    from utils.chamfer_loss import ChamferLoss
    cl = ChamferLoss()
    cl.use_cuda = False
    out = {}
    for tag, (B, Np, Ng, D) in {"d3": (4, 128, 96, 3), "d9": (3, 64, 64, 9), "d3sq": (2, 256, 256, 3)}.items():
        preds = clouds(g, B, Np, D).requires_grad_(True)
        gts = clouds(g, B, Ng, D).requires_grad_(True)
        loss = cl(preds, gts)
        loss.backward()
        out.update({f"{tag}_preds": preds.detach().numpy(), f"{tag}_gts": gts.detach().numpy(),
                    f"{tag}_loss": loss.detach().numpy(), f"{tag}_gpreds": preds.grad.numpy(),
                    f"{tag}_ggts": gts.grad.numpy()})
    np.savez_compressed(os.path.join(OUT, "chamfer_loss.npz"), **out)

    # ---- evaluation_metrics (distChamfer, _pairwise_EMD_CD_, lgan_mmd_cov, knn, compute_all_metrics) ---
    ns = {"torch": torch, "np": np}
    extract(os.path.join(REF, "evaluation/evaluation_metrics.py"),
            ["distChamfer", "_pairwise_EMD_CD_", "knn", "lgan_mmd_cov", "compute_all_metrics"], ns)
    ns["emd_approx"] = lambda s, r: torch.zeros(s.size(0))
    ns["distChamferCUDA"] = None
    a = clouds(g, 3, 256, 3, kind="S")
    b = clouds(g, 3, 256, 3, kind="S")
    dl, dr = ns["distChamfer"](a, b)
    smp = clouds(g, 6, 128, 3, kind="S")
    ref = clouds(g, 6, 128, 3, kind="S")
    all_cd, _ = ns["_pairwise_EMD_CD_"](smp, ref, 4, accelerated_cd=False)
    m_rr, _ = ns["_pairwise_EMD_CD_"](ref, ref, 4, accelerated_cd=False)
    m_ss, _ = ns["_pairwise_EMD_CD_"](smp, smp, 4, accelerated_cd=False)
    res = ns["compute_all_metrics"](smp, ref, 4, accelerated_cd=False)
    mm = ns["lgan_mmd_cov"](all_cd.t())
    nn1 = ns["knn"](m_rr, all_cd, m_ss, 1, sqrt=False)
    ev = {"a": a.numpy(), "b": b.numpy(), "dl": dl.numpy(), "dr": dr.numpy(), "smp": smp.numpy(), "ref": ref.numpy(),
          "all_cd": all_cd.numpy(), "m_rr": m_rr.numpy(), "m_ss": m_ss.numpy()}
    for k, v in res.items():
        if k.endswith("-CD") or "-CD-" in k:
            ev["metric:" + k] = np.asarray(v.item(), dtype=np.float64)
    for k, v in mm.items():
        ev["mmdcov:" + k] = np.asarray(v.item(), dtype=np.float64)
    for k, v in nn1.items():
        ev["knn:" + k] = np.asarray(v.item(), dtype=np.float64)
    # JSD (evaluation_metrics.py:206-321): numpy + sklearn + scipy on the reference side
    from numpy.linalg import norm
    from scipy.stats import entropy
    from sklearn.neighbors import NearestNeighbors
    import warnings
    ns3 = {"np": np, "norm": norm, "entropy": entropy, "NearestNeighbors": NearestNeighbors, "warnings": warnings}
    extract(os.path.join(REF, "evaluation/evaluation_metrics.py"),
            ["unit_cube_grid_point_cloud", "jsd_between_point_cloud_sets", "entropy_of_occupancy_grid",
             "jensen_shannon_divergence", "_jsdiv"], ns3)
    jsmp = (clouds(g, 12, 512, 3, kind="S") * 0.5).numpy()      # inside the radius-0.5 sphere the grid is clipped to
    jref = (clouds(g, 10, 512, 3, kind="U") * 0.28).numpy()
    ent_s, cnt_s = ns3["entropy_of_occupancy_grid"](jsmp, 28, True)
    ent_r, cnt_r = ns3["entropy_of_occupancy_grid"](jref, 28, True)
    ev.update({"jsd_smp": jsmp, "jsd_ref": jref, "jsd_value": np.asarray(ns3["jsd_between_point_cloud_sets"](jsmp, jref), dtype=np.float64),
               "jsd_entropy_smp": np.asarray(ent_s, dtype=np.float64), "jsd_entropy_ref": np.asarray(ent_r, dtype=np.float64),
               "jsd_counters_smp": cnt_s, "jsd_counters_ref": cnt_r})
    np.savez_compressed(os.path.join(OUT, "evaluation_metrics.npz"), **ev)

    # ---- get_edge_features{,_xyz} (models/PDGNet_v2.py:439-528) -------------------------------------
    ns2 = {"torch": torch}
    extract(os.path.join(REF, "models/PDGNet_v2.py"), ["get_edge_features", "get_edge_features_xyz"], ns2)
    picked = []
    real_sort = torch.sort

    class _T:  # proxy so that the extracted functions see torch.sort recording its indices
        def __getattr__(self, name):
            if name == "sort":
                def rec(*a, **kw):
                    r = real_sort(*a, **kw)
                    picked.append(r[1])
                    return r
                return rec
            return getattr(torch, name)
    ns2["torch"] = _T()
    x = torch.randn(2, 16, 64, generator=g)
    pc = clouds(g, 2, 3, 64)
    ee = ns2["get_edge_features"](x, 10)
    idx_a = picked[-1][:, :, 1:11]
    e_fea, e_xyz = ns2["get_edge_features_xyz"](x, pc, 10)
    idx_b = picked[-1][:, :, 1:11]
    np.savez_compressed(os.path.join(OUT, "edge_features.npz"), x=x.numpy(), pc=pc.numpy(), ee=ee.numpy(),
                        idx=idx_a.numpy(), e_fea=e_fea.numpy(), e_xyz=e_xyz.numpy(), idx_xyz=idx_b.numpy())

    # ---- KNNQueryNaive / KNNQueryExclude (lib/pointops/functions/pointops.py:368-474) ----------------
    sys.modules["pointops_cuda"] = types.ModuleType("pointops_cuda")
    sys.path.insert(0, os.path.join(REF, "lib/pointops/functions"))
    import pointops as ref_pointops
    xyz = clouds(g, 2, 128, 3)
    new_xyz = clouds(g, 2, 40, 3)
    idx_naive = ref_pointops.KNNQueryNaive.forward(None, 20, xyz, new_xyz)
    idx_self = ref_pointops.KNNQueryNaive.forward(None, 8, xyz, None)
    idx_excl = ref_pointops.KNNQueryExclude.forward(None, 8, xyz, None)
    np.savez_compressed(os.path.join(OUT, "knn_naive.npz"), xyz=xyz.numpy(), new_xyz=new_xyz.numpy(),
                        idx_naive=idx_naive.numpy(), idx_self=idx_self.numpy(), idx_excl=idx_excl.numpy())
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
