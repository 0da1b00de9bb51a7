#!/bin/bash
# A/B of the narrow-barrier planner (QB_NARROW_SYNC): parity tests with the default, then bench lines with it off / on /
# "u" (timing-only upper bound: every inner barrier a __syncwarp, results wrong).   usage: bash tools/gpu_ab_narrow.sh
out=gpurun_out
mkdir -p $out
timeout 420 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > $out/pytest_narrow.log
for v in 0 1 u; do
  QB_NARROW_SYNC=$v timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/bench_c2_narrow_$v.json 2> $out/bench_c2_narrow_$v.err
done
for v in 0 1; do
  QB_NARROW_SYNC=$v timeout 200 python bench.py --workload q20 --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_q20_narrow_$v.json 2> $out/bench_q20_narrow_$v.err
done
cat $out/pytest_narrow.log
for f in $out/bench_*_narrow_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["forward_sweep"]["avg_launch_ms"], d["clocks"])
except Exception as e:
    print("ERR", e)
PY
done
