"""A/B of two builds of the EMD kernel: bitwise comparison + timing (tools/ only).  Usage: emd_ab.py <other .so>"""
import ctypes, sys, subprocess, os
import numpy as np
import torch
sys.path.insert(0, ".")
code = r'''
import sys, os, numpy as np, torch
sys.path.insert(0, ".")
from pdgn_b200 import _lib, ops
if os.environ.get("PDGN_LIB"): _lib.SO_PATH = os.environ["PDGN_LIB"]
rng = np.random.default_rng(5)
def sph(k, n):
    v = rng.standard_normal((k, n, 3)); v /= np.linalg.norm(v, axis=-1, keepdims=True)
    return torch.from_numpy(v.astype(np.float32)).cuda()
outs = []
for (na, nb, n, m) in [(6, 5, 2048, 2048), (3, 4, 1000, 1000), (2, 3, 300, 700), (2, 2, 2048, 512)]:
    outs.append(ops.emd_allpairs(sph(na, n), sph(nb, m)).cpu().numpy().ravel())
np.save(sys.argv[1], np.concatenate(outs))
'''
open("/tmp/emd_ab_child.py", "w").write(code)
subprocess.check_call([sys.executable, "/tmp/emd_ab_child.py", "/tmp/emd_new.npy"])
subprocess.check_call([sys.executable, "/tmp/emd_ab_child.py", "/tmp/emd_old.npy"], env=dict(os.environ, PDGN_LIB=sys.argv[1]))
a, b = np.load("/tmp/emd_new.npy"), np.load("/tmp/emd_old.npy")
print("values:", a.size, "bitwise equal:", bool((a.view(np.uint32) == b.view(np.uint32)).all()), "max rel diff: %.3e" % float(np.max(np.abs(a - b) / np.abs(b))))
