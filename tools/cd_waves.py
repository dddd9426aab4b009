"""Kernel-only time of the 1000x1000 all-pairs CD launch for the current PDGN_CD_WAVES (tools/ only)."""
import os, sys
import torch
sys.path.insert(0, ".")
import bench
from pdgn_b200 import ops
dev = torch.device("cuda:0")
a, b = bench.make_clouds(0).to(dev), bench.make_clouds(1).to(dev)
ts = []
for _ in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.cd_allpairs(a, b); e1.record(); torch.cuda.synchronize()
    ts.append(round(e0.elapsed_time(e1), 1))
print("waves", os.environ.get("PDGN_CD_WAVES"), ts, flush=True)
