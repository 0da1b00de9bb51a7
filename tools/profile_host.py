"""Host-side cost of one small-circuit step (config 1 by default): torch.profiler table of CPU ops + CUDA kernels.
   python tools/profile_host.py [c1|c2] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench
import qandle_b200 as q

name = sys.argv[1] if len(sys.argv) > 1 else "c1"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
wl = bench.WORKLOADS[name]
torch.manual_seed(0)
circ = bench.build_circuit(q, wl).to("cuda")
params = list(circ.parameters())
x = torch.rand(wl["batch"], wl["n"], device="cuda", requires_grad=True)
g = torch.randn(wl["batch"], wl["n"], device="cuda")


def step():
    for p in params:
        p.grad = None
    x.grad = None
    out = circ(x=x)
    out.backward(g)


for _ in range(10):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=40, max_name_column_width=60))
