import sys
import numpy as np, torch
sys.path.insert(0, ".")
from pdgn_b200._lib import lib, check
L = lib(); dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
al = lambda v: (v + 255) & ~255
rng = np.random.default_rng(0)
b, c, n, k = 1, 16, 512, 10
x = rng.standard_normal((b, c, n)).astype(np.float32)
xt = torch.from_numpy(x).to(dev)
idx = torch.empty((b, n, k), dtype=torch.int64, device=dev)
wsb = L.pdgn_knn_feat_workspace(b, c, n)
ws = torch.zeros((wsb,), dtype=torch.uint8, device=dev)
check(L.pdgn_knn_feat_ws(xt.data_ptr(), b, c, n, k, 1, idx.data_ptr(), None, ws.data_ptr(), wsb, st), "ws")
torch.cuda.synchronize()
base = (ws.data_ptr() + 255) & ~255
o = base - ws.data_ptr()
o_mean = o; o += al(b * c * 4)
o_xc = o; o += al(b * c * n * 4)
o_xT = o; o += al(b * c * n * 4)
o_nrm = o; o += al(b * n * 4)
o_max = o; o += al(b * 4)
o_cand = o
xc = ws[o_xc: o_xc + b * c * n * 4].view(torch.float32).view(b, n, c).permute(0, 2, 1).contiguous().cpu().numpy()
nrm = ws[o_nrm: o_nrm + b * n * 4].view(torch.float32).cpu().numpy()
G = ws[o_cand: o_cand + 128 * 128 * 4].view(torch.float32).view(128, 128).cpu().numpy() - 1000.0
ref = xc[0].T.astype(np.float64) @ xc[0].astype(np.float64)      # [n, n]
print("xc mean abs", np.abs(xc).mean(), "centred?", np.abs(xc[0].mean(1)).max())
print("nrm ok", np.allclose(nrm, (xc[0] ** 2).sum(0), rtol=1e-4))
R = ref[:128, :128]
print("max |G - ref|", np.abs(G - R).max(), " max|ref|", np.abs(R).max())
print("G[0,:8]  ", G[0, :8]); print("ref[0,:8]", R[0, :8])
print("G[:8,0]  ", G[:8, 0]); print("ref[:8,0]", R[:8, 0])
# which reference entry does G[r, cidx] equal?
full = ref
for (r, cc) in [(0, 0), (0, 1), (1, 0), (0, 4), (4, 0), (5, 9), (37, 77)]:
    v = G[r, cc]
    hit = np.argwhere(np.abs(full - v) < 1e-3 * max(1.0, abs(v)))
    print("G[%d,%d]=%.4f matches ref at" % (r, cc, v), hit[:6].tolist())
