#!/bin/bash
# One bench line of the persistent-CTA experiment build (qandle_b200/_variants/dyn, QB_DYN=1) on config 2, then 20 qubits.
out=gpurun_out
mkdir -p $out
V=$PWD/qandle_b200/_variants/dyn
for wl in c2 q20; do
  f=$out/bench_dyn_$wl.json
  QB_LIB_DIR=$V QB_DYN=1 timeout 30 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > $f 2> ${f%.json}.err
  python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], round(d["value"],1), "evals/s", round(d["ms_per_step"],2), "ms; adjoint", round(r["avg_launch_ms"],3), round(r["frac"],3), "fwd", round(r["forward_sweep"]["avg_launch_ms"],3), round(r["forward_sweep"]["frac"],3), d["clocks"]["sm_mhz"])
except Exception as e:
    print(sys.argv[1], "ERR", e); print(open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
done
