"""Round-2 kNN kernel (csrc/knn_gram.cu) on the B200 box: exactness against the CPU oracle over the parity-suite shapes it is
eligible for (forced with PDGN_B200_TUNE=1 PDGN_KNN_IMPL=gram), then timings against the select kernel.  tools/ only.
Usage: PDGN_B200_TUNE=1 PDGN_KNN_IMPL=gram python tools/knn_check.py [--time-only]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import clouds_sphere, clouds_ties, clouds_uniform  # noqa: E402
from oracle import cpu as ocpu  # noqa: E402
from pdgn_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
impl = os.environ.get("PDGN_KNN_IMPL", "default")


def coherent(rng, b, n, _3):
    """index-coherent cloud: points sorted along a space-filling-ish key (neighbours in space are neighbours in index)"""
    v = clouds_sphere(rng, b, n, 3)
    for i in range(b):
        key = np.floor((v[i] + 1.2) * 4).astype(np.int64)
        order = np.lexsort((v[i][:, 2], key[:, 2], key[:, 1], key[:, 0]))
        v[i] = v[i][order]
    return v


def dup_heavy(rng, b, n, _3):
    v = clouds_uniform(rng, b, n, 3)
    v[:, n // 2:] = v[:, : n - n // 2]       # every point twice
    v[:, : n // 8] = v[:, :1]                # and one point n/8 times
    return v


def far_origin(rng, b, n, _3):
    return (clouds_uniform(rng, b, n, 3) * 0.5 + np.array([40.0, -25.0, 10.0], np.float32)).astype(np.float32)


CASES = [("U-self", clouds_uniform, 2, 300, None, 20), ("S-1000x257", clouds_sphere, 3, 1000, 257, 20), ("T-ties-512", clouds_ties, 2, 512, None, 20),
         ("S-2048-self", clouds_sphere, 2, 2048, None, 20), ("T-ties-2048", clouds_ties, 2, 2048, 600, 20), ("U-k12", clouds_uniform, 2, 2048, 300, 12),
         ("U-ss4-ragged", clouds_uniform, 2, 501, 130, 20), ("U-ss8-ragged", clouds_uniform, 2, 1001, 130, 20), ("U-ss16-ragged", clouds_uniform, 2, 2039, 130, 16),
         ("S-k1", clouds_sphere, 2, 2048, 2048, 1), ("U-k3", clouds_uniform, 2, 1024, 2048, 3), ("C-coherent-2048", coherent, 2, 2048, None, 20),
         ("C-coherent-1024", coherent, 2, 1024, 700, 20), ("D-dup-heavy", dup_heavy, 2, 2048, 515, 20), ("F-far-origin", far_origin, 2, 2048, 300, 20),
         ("U-m-ragged-513", clouds_uniform, 1, 1500, 513, 20), ("U-257", clouds_uniform, 2, 257, 257, 20)]

if "--time-only" not in sys.argv:
    bad = 0
    for name, maker, b, n, m, k in CASES:
        rng = np.random.default_rng(abs(hash(name)) % 1000)
        xyz = maker(rng, b, n, 3)
        q = xyz if m is None else maker(rng, b, m, 3)
        idx, d2 = ops.knn_xyz(k, torch.from_numpy(xyz).to(dev), torch.from_numpy(q).to(dev), return_dist=True)
        oi, od = ocpu.knn_xyz(xyz, q, k)
        ok = np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(d2.cpu().numpy(), od)
        bad += not ok
        print("%-18s b%d n%d m%s k%d  %s" % (name, b, n, m, k, "exact" if ok else "MISMATCH (%d idx, %d d2 entries differ)" % (
            (idx.cpu().numpy() != oi).sum(), (d2.cpu().numpy() != od).sum())), flush=True)
    # NaN / inf handling
    rng = np.random.default_rng(5)
    xyz = clouds_uniform(rng, 1, 600, 3)
    q = clouds_uniform(rng, 1, 40, 3)
    xyz[0, 7] = np.nan; xyz[0, 100, 1] = np.inf; xyz[0, 200] = 1e30; q[0, 3, 0] = np.nan; q[0, 5] = np.inf; q[0, 9] = 3e19
    idx, d2 = ops.knn_xyz(20, torch.from_numpy(xyz).to(dev), torch.from_numpy(q).to(dev), return_dist=True)
    oi, od = ocpu.knn_xyz(xyz, q, 20)
    ok = np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(d2.cpu().numpy(), od)
    bad += not ok
    print("non-finite inputs   %s" % ("exact" if ok else "MISMATCH"))
    print("impl=%s: %d mismatching cases" % (impl, bad))

flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def t(fn, reps=20):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


rng = np.random.default_rng(0)
for name, maker, b, n, m, k in [("cfg2 uniform", clouds_uniform, 35, 2048, 2048, 20), ("cfg2 sphere", clouds_sphere, 35, 2048, 2048, 20),
                                ("cfg2 coherent", coherent, 35, 2048, 2048, 20), ("cfg2 ties", clouds_ties, 35, 2048, 2048, 20),
                                ("cfg2 dup-heavy", dup_heavy, 35, 2048, 2048, 20), ("n1024", clouds_uniform, 70, 1024, 1024, 20),
                                ("n512", clouds_uniform, 140, 512, 512, 20), ("k10", clouds_uniform, 35, 2048, 2048, 10)]:
    xyz = torch.from_numpy(maker(rng, b, n, 3)).to(dev)
    ms = t(lambda: ops.knn_xyz(k, xyz))
    print("%-14s impl=%-7s b%d n%d k%d  %.4f ms  %.1f Mq/s  issue-frac %.3f" % (name, impl, b, n, k, ms, b * n / ms / 1e3, b * n * n * 6.0 / (ms * 1e-3) / (148 * 128 * 1.965e9)), flush=True)

print("-- training shapes (B=35): queries m against n candidates")
for (b, n, m, k) in [(35, 2048, 1024, 20), (35, 2048, 512, 20), (35, 2048, 256, 20), (35, 1024, 1024, 20), (35, 1024, 512, 20), (35, 1024, 256, 20),
                     (35, 512, 512, 20), (35, 512, 256, 20), (8, 2048, 2048, 20), (16, 2048, 2048, 20)]:
    xyz = torch.from_numpy(clouds_uniform(rng, b, n, 3)).to(dev)
    q = torch.from_numpy(clouds_uniform(rng, b, m, 3)).to(dev)
    ms = t(lambda: ops.knn_xyz(k, xyz, q))
    print("b%d n%d m%d k%d impl=%-7s %.4f ms" % (b, n, m, k, impl, ms), flush=True)
