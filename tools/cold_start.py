"""Per-step times of the all-pairs CD step in a fresh process on a fresh box, without and with nvidia-smi polling (tools/ only)."""
import subprocess, sys, time
import torch
sys.path.insert(0, ".")
import bench
from pdgn_b200 import ops
dev = torch.device("cuda:0")
a, b = bench.make_clouds(0).to(dev), bench.make_clouds(1).to(dev)
def steps(k, tag):
    out = []
    for _ in range(k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.cd_allpairs(a, b); e1.record(); torch.cuda.synchronize()
        out.append(round(e0.elapsed_time(e1), 1))
    print(tag, out, flush=True)
steps(6, "no polling   ")
p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=" + bench.ClockSampler.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
steps(8, "polling 100ms")
p.terminate(); p.communicate()
steps(4, "no polling   ")
p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=" + bench.ClockSampler.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
steps(4, "polling again")
p.terminate(); p.communicate()
