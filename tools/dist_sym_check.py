"""torchrun check of the symmetric multi-GPU tiling against the single-GPU matrix (tools/ only).
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_sym_check.py"""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=dev)
import bench
from pdgn_b200 import dist as pd, ops
ok = True
for n in (37, 256, 1000):
    A = bench.make_clouds(2, n).to(dev)
    full = pd.pairwise_cd(A, A)          # same tensor -> symmetric plan
    ref = ops.cd_allpairs(A, A)
    same = bool(torch.equal(full, ref))
    ok &= same
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pd.pairwise_cd(A, A); e1.record(); torch.cuda.synchronize()
    t_sym = e0.elapsed_time(e1)
    B = A.clone()
    e0.record(); pd.pairwise_cd(A, B); e1.record(); torch.cuda.synchronize()
    t_full = e0.elapsed_time(e1)
    if rank == 0:
        print("n=%d: symmetric tiling equals the single-GPU matrix bit for bit: %s;  %.1f ms vs %.1f ms for the plain tiling" % (n, same, t_sym, t_full), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
