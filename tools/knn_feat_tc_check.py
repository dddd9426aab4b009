"""Bring-up / A-B tool for the tensor-core feature kNN (csrc/knn_feat_tc.cu): exactness against the oracle, list statistics
(read back from the workspace), timings against the FP32 SIMT kernel.  tools/ only."""
import ctypes, os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from oracle import cpu as ocpu
from pdgn_b200._lib import lib, check
L = lib(); dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
al = lambda v: (v + 255) & ~255

def run(x, k, skip, simt=False):
    b, c, n = x.shape
    idx = torch.empty((b, n, k), dtype=torch.int64, device=dev)
    d2 = torch.empty((b, n, k), dtype=torch.float32, device=dev)
    if simt:
        check(L.pdgn_knn_feat(x.data_ptr(), b, c, n, k, skip, idx.data_ptr(), d2.data_ptr(), st), "simt")
        return idx, d2, None
    wsb = L.pdgn_knn_feat_workspace(b, c, n)
    ws = torch.zeros((wsb,), dtype=torch.uint8, device=dev)
    check(L.pdgn_knn_feat_ws(x.data_ptr(), b, c, n, k, skip, idx.data_ptr(), d2.data_ptr(), ws.data_ptr(), wsb, st), "ws")
    torch.cuda.synchronize()
    base = (ws.data_ptr() + 255) & ~255
    off = base - ws.data_ptr() + al(b * c * 4) + al(b * ((c + 31) // 32) * 32 * n * 4) + al(b * c * n * 4) + al(b * n * 4 * 4) + al(b * n * 64 * 4)
    cnt = ws[off: off + b * n * 4].view(torch.int32).cpu().numpy()
    return idx, d2, cnt

rng = np.random.default_rng(0)
def relu_feat(b, c, n):      # generator-like: non-negative, correlated channels, a common offset
    z = rng.standard_normal((b, 8, n)).astype(np.float32)
    w = rng.standard_normal((c, 8)).astype(np.float32)
    return np.maximum(np.einsum("ck,bkn->bcn", w, z) + 1.0 + 0.1 * rng.standard_normal((b, c, n)).astype(np.float32), 0).astype(np.float32)
def coherent(b, c, n):       # neighbours are index neighbours
    t = np.linspace(0, 1, n, dtype=np.float32)
    base = np.stack([np.sin((i + 1) * t * 3.0) for i in range(c)]).astype(np.float32)
    return (base[None] + 1e-3 * rng.standard_normal((b, c, n))).astype(np.float32)
cases = [("randn", lambda: rng.standard_normal((2, 64, 256)).astype(np.float32), 10),
         ("randn-c256", lambda: rng.standard_normal((1, 256, 1024)).astype(np.float32), 10),
         ("randn-c32-n128", lambda: rng.standard_normal((3, 32, 128)).astype(np.float32), 10),
         ("randn-c8", lambda: rng.standard_normal((2, 8, 384)).astype(np.float32), 5),
         ("relu", lambda: relu_feat(2, 128, 512), 10),
         ("coherent", lambda: coherent(2, 64, 512), 10),
         ("dups", lambda: np.repeat(rng.standard_normal((2, 32, 128)).astype(np.float32), 2, axis=2), 10),
         ("c40-n640", lambda: rng.standard_normal((2, 40, 640)).astype(np.float32), 19)]
bad = 0
for name, mk, k in cases:
    x = mk()
    xt = torch.from_numpy(x).to(dev)
    idx, d2, cnt = run(xt, k, 1)
    ri, rd = ocpu.knn_feat(x, k, skip=1)
    ok = np.array_equal(idx.cpu().numpy(), ri) and np.array_equal(d2.cpu().numpy(), rd)
    bad += 0 if ok else 1
    mism = float((idx.cpu().numpy() != ri).mean())
    print("%-16s %s  %s  flagged %.2f%%  mean list %.1f  max %d  idx mismatch %.4f" % (
        name, x.shape, "exact" if ok else "MISMATCH", 100.0 * (cnt < 0).mean(), cnt[cnt >= 0].mean() if (cnt >= 0).any() else -1,
        cnt.max(), mism), flush=True)
print("mismatching cases:", bad, flush=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (b, c, n) in [(35, 256, 1024), (35, 128, 512), (35, 64, 256), (35, 32, 128)]:
    x = torch.randn(b, c, n, device=dev)
    idx = torch.empty((b, n, 10), dtype=torch.int64, device=dev)
    wsb = L.pdgn_knn_feat_workspace(b, c, n)
    ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
    t_tc = bench._time_ms(lambda: L.pdgn_knn_feat_ws(x.data_ptr(), b, c, n, 10, 1, idx.data_ptr(), None, ws.data_ptr(), wsb, st), 10, flush)
    t_si = bench._time_ms(lambda: L.pdgn_knn_feat(x.data_ptr(), b, c, n, 10, 1, idx.data_ptr(), None, st), 5, flush)
    print("B%d C%d N%d k10: tensor-core path %.3f ms, SIMT kernel %.3f ms" % (b, c, n, t_tc, t_si), flush=True)
