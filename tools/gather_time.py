"""Backward gather (pull) timings on the feature-sized shapes (tools/ only): grouping bwd and edge-feature bwd."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
import os
from pdgn_b200 import _lib, ops
if os.environ.get('PDGN_LIB'):
    _lib.SO_PATH = os.environ['PDGN_LIB']  # A/B against another build (tools only)
    print('library:', _lib.SO_PATH)
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (b, c, n, k) in [(35, 256, 1024, 10), (35, 128, 512, 10), (35, 64, 256, 10)]:
    go = torch.randn(b, c, n, k, device=dev)
    idx = torch.from_numpy(rng.integers(0, n, (b, n, k)).astype(np.int32)).to(dev)
    ms = bench._time_ms(lambda: ops.group_bwd(go, idx, n), 10, flush)
    gb = (go.numel() + idx.numel() + 2 * b * c * n) * 4 / 1e9
    gee = torch.randn(b, 2 * c, n, k, device=dev)
    idx64 = idx.long()
    ms2 = bench._time_ms(lambda: ops.edge_feat_bwd(gee, idx64, c), 10, flush)
    gb2 = (gee.numel() + 2 * b * c * n) * 4 / 1e9 + idx64.numel() * 8 / 1e9
    print("B%d C%d N%d k%d: group_bwd %.1f us (%.0f GB/s)   edge_feat_bwd %.1f us (%.0f GB/s)" % (b, c, n, k, ms * 1e3, gb / ms * 1e3, ms2 * 1e3, gb2 / ms2 * 1e3), flush=True)

from pdgn_b200._lib import lib
L = lib()
st = torch.cuda.current_stream().cuda_stream
for (b, c, m, n) in [(35, 256, 1024, 2048), (35, 128, 512, 1024), (35, 64, 256, 512)]:
    feat = torch.randn(b, c, m, device=dev)
    idx3 = torch.from_numpy(rng.integers(0, m, (b, n, 3)).astype(np.int32)).to(dev)
    w3 = torch.rand(b, n, 3, device=dev)
    out = torch.empty(b, c, n, device=dev)
    ms = bench._time_ms(lambda: L.pdgn_interp_fwd(feat.data_ptr(), idx3.data_ptr(), w3.data_ptr(), b, c, m, n, out.data_ptr(), st), 10, flush)
    gb = (feat.numel() + out.numel() + 2 * idx3.numel()) * 4 / 1e9
    print("interp fwd B%d C%d m%d n%d: %.1f us (%.0f GB/s) [PDGN_IS_ROW_KB=%s]" % (b, c, m, n, ms * 1e3, gb / ms * 1e3, os.environ.get("PDGN_IS_ROW_KB", "default")), flush=True)

for (b, c, m, n) in [(35, 256, 1024, 2048), (35, 128, 512, 1024), (35, 64, 256, 512)]:
    go = torch.randn(b, c, n, device=dev)
    idx3 = torch.from_numpy(rng.integers(0, m, (b, n, 3)).astype(np.int32)).to(dev)
    w3 = torch.rand(b, n, 3, device=dev)
    grad = torch.zeros(b, c, m, device=dev)
    ws_bytes = L.pdgn_interp_bwd_workspace(b, n, m)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    ms = bench._time_ms(lambda: L.pdgn_interp_bwd_ws(go.data_ptr(), idx3.data_ptr(), w3.data_ptr(), b, c, n, m, grad.data_ptr(), ws.data_ptr(), ws_bytes, st), 10, flush)
    gb = (go.numel() + grad.numel() + 2 * idx3.numel()) * 4 / 1e9
    print("interp bwd B%d C%d n%d m%d: %.1f us (%.0f GB/s)" % (b, c, n, m, ms * 1e3, gb / ms * 1e3), flush=True)
for (b, c, n, k) in [(35, 256, 1024, 10), (35, 32, 128, 10)]:
    go = torch.randn(b, c, n, k, device=dev)
    idx = torch.from_numpy(rng.integers(0, n, (b, n, k)).astype(np.int32)).to(dev)
    grad = torch.zeros(b, c, n, device=dev)
    ws_bytes = L.pdgn_group_bwd_workspace(b, n, n, k)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    ms = bench._time_ms(lambda: L.pdgn_group_bwd_ws(go.data_ptr(), idx.data_ptr(), b, c, n, n, k, grad.data_ptr(), ws.data_ptr(), ws_bytes, st), 10, flush)
    gb = (go.numel() + idx.numel() + b * c * n) * 4 / 1e9
    print("group_bwd_ws (C ABI) B%d C%d N%d k%d: %.1f us (%.0f GB/s algorithmic)" % (b, c, n, k, ms * 1e3, gb / ms * 1e3), flush=True)
