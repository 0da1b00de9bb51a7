#!/bin/bash
# A/B of one engine knob on the GPU box: parity tests with the default, then bench lines per value.
#   usage: bash tools/gpu_ab.sh ENVVAR "v1 v2 ..." [workloads]      e.g.  bash tools/gpu_ab.sh QB_FULL_TILE "0 1" "c2 q20"
var=$1; vals=$2; wls=${3:-c2}
out=gpurun_out
mkdir -p $out
timeout 420 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > $out/pytest_ab_$var.log
cat $out/pytest_ab_$var.log
for wl in $wls; do
  for v in $vals; do
    f=$out/bench_${wl}_${var}_$v.json
    env $var=$v timeout 200 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > $f 2> ${f%.json}.err
    python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"],1), "evals/s", round(d["ms_per_step"],2), "ms; adjoint sweep", round(d["roofline"]["avg_launch_ms"],3), "fwd sweep", round(d["roofline"]["forward_sweep"]["avg_launch_ms"],3), d["clocks"]["sm_mhz"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
  done
done
