"""Edge-feature forward: staged vs plain kernel, C=256 N=1024 k=10 B=35 (tools/ only)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from pdgn_b200._lib import lib
L = lib(); dev = torch.device("cuda:0"); rng = np.random.default_rng(0)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (b, c, n, k) in [(35, 256, 1024, 10), (35, 128, 512, 10), (35, 64, 256, 10), (35, 32, 128, 10)]:
    x = torch.randn(b, c, n, device=dev)
    idx = torch.from_numpy(rng.integers(0, n, (b, n, k)).astype(np.int64)).to(dev)
    ee = torch.empty(b, 2 * c, n, k, device=dev)
    ms = bench._time_ms(lambda: L.pdgn_edge_feat_fwd(x.data_ptr(), idx.data_ptr(), b, c, n, k, ee.data_ptr(), st), 10, flush)
    gb = (ee.numel() + x.numel()) * 4 / 1e9 + idx.numel() * 8 / 1e9
    print("edge fwd B%d C%d N%d k%d: %.1f us (%.0f GB/s)" % (b, c, n, k, ms * 1e3, gb / ms * 1e3), flush=True)
