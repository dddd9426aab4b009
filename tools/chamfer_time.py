"""chamfer_min timings on the training shapes (tools/ only)."""
import sys
import torch
sys.path.insert(0, ".")
import bench
import os
from pdgn_b200 import _lib, ops
if os.environ.get('PDGN_LIB'):
    _lib.SO_PATH = os.environ['PDGN_LIB']  # A/B against another build (tools only)
    print('library:', _lib.SO_PATH)
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (b, nx, ny, d) in [(35, 1024, 1024, 3), (35, 2048, 2048, 3), (35, 512, 512, 3), (35, 256, 256, 3), (35, 1024, 1024, 9), (35, 2048, 2048, 9), (35, 256, 256, 9), (1000, 2048, 2048, 3)]:
    x, y = torch.rand(b, nx, d, device=dev), torch.rand(b, ny, d, device=dev)
    ms = bench._time_ms(lambda: ops.chamfer_min(x, y), 10, flush)
    print("b%d %dx%d d%d: %.1f us (both directions)" % (b, nx, ny, d, ms * 1e3), flush=True)
