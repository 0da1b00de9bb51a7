#!/bin/bash
# Parity tests + the standard timing set after a kernel change.  usage: bash tools/gpu_check.sh <tag>
tag=${1:-chk}
out=gpurun_out
mkdir -p $out
timeout 420 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for wl in c2 c2 q20 c3; do
  f=$out/bench_${wl}_$tag.json
  timeout 200 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > $f 2> ${f%.json}.err
  python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"],1), "evals/s", round(d["ms_per_step"],2), "ms; adjoint sweep", round(d["roofline"]["avg_launch_ms"],3), "fwd sweep", round(d["roofline"]["forward_sweep"]["avg_launch_ms"],3), d["clocks"]["sm_mhz"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
timeout 300 python tools/run_sharded.py --qubits 27 --layers 10 --dtype c128 --backward --out $out/c128_27q_$tag.json 2>&1 | tail -1 | cut -c1-260
