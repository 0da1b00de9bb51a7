"""One forward+backward of a bench workload, for ncu captures (B200_PROFILING.md):
   ncu --set full --clock-control none --import-source on -k regex:sweep_staged -s 24 -c 2 -o gpurun_out/prof python tools/profile_step.py c2
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import qandle_b200 as q

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
B = int(sys.argv[2]) if len(sys.argv) > 2 else wl["batch"]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
torch.manual_seed(0)
circ = bench.build_circuit(q, wl).to("cuda")
x = torch.rand(B, wl["n"], device="cuda", requires_grad=True)
g = torch.randn(B, wl["n"], device="cuda")
for _ in range(reps):
    out = circ(x=x)
    out.backward(g)
torch.cuda.synchronize()
print("done", float(out.sum()))
