#!/bin/bash
# Last GPU check of round 1: the GPU suite and three bench lines on the committed default (streaming adjoint kernel on).
out=gpurun_out
mkdir -p $out
timeout 60 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3 | tee $out/pytest_last.log
for wl in c2 q20 c3; do
  f=$out/bench_last_$wl.json
  timeout 40 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > $f 2> ${f%.json}.err
  python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], round(d["value"],1), "evals/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["value"],1), "; adjoint", round(r["avg_launch_ms"],3), round(r["frac"],3), "fwd", round(r["forward_sweep"]["avg_launch_ms"],3), round(r["forward_sweep"]["frac"],3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
