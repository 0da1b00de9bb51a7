#!/bin/bash
# A/B of the streaming adjoint kernel (QB_ADJ_STREAM=1): GPU parity subset under the knob, then bench lines off / on.
out=gpurun_out
mkdir -p $out
t0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - t0 ))s] $*"; }
summ() {
python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]
    print(sys.argv[1], round(d["value"],1), "evals/s", round(d["ms_per_step"],2), "ms; adjoint sweep", round(r["avg_launch_ms"],3), "frac", round(r["frac"],3),
          "fwd sweep", round(r["forward_sweep"]["avg_launch_ms"],3), "fp32 adj frac", round(r["fp32"]["adjoint_sweeps"]["frac"],3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
}
QB_ADJ_STREAM=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_packed_sel.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4 | tee $out/pytest_stream.log
el "pytest stream done"
for wl in c2 q20 c3; do
  for v in 0 1; do
    f=$out/bench_stream${v}_$wl.json
    QB_ADJ_STREAM=$v timeout 150 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > $f 2> ${f%.json}.err; summ $f
  done
done
el "bench done"
QB_ADJ_STREAM=1 timeout 240 ncu --set full --clock-control none --import-source on -k regex:sweep_flat -s 37 -c 3 -o $out/prof_stream_bwd -f python tools/profile_step.py c2 4096 2 > $out/ncu_stream_bwd.log 2>&1
el "ncu done"
