#!/bin/bash
# A/B of the CTAs-per-sample rule (capi.cu: choose_cps; QB_CPS_MODEL=0 = old "two waves" rule): GPU suite on the new rule,
# then bench lines old / new (and new + streaming adjoint).
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
    print(sys.argv[1], round(d["value"],1), "evals/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["value"],1), "; adjoint sweep", round(r["avg_launch_ms"],3), "frac", round(r["frac"],3),
          "fwd sweep", round(r["forward_sweep"]["avg_launch_ms"],3), "frac", round(r["forward_sweep"]["frac"],3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
}
timeout 200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3 | tee $out/pytest_cps.log
el "pytest done"
for wl in q20 c3; do
  f=$out/bench_cps_old_$wl.json
  QB_CPS_MODEL=0 timeout 100 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > $f 2> ${f%.json}.err; summ $f
  f=$out/bench_cps_new_$wl.json
  timeout 100 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > $f 2> ${f%.json}.err; summ $f
  f=$out/bench_cps_new_stream_$wl.json
  QB_ADJ_STREAM=1 timeout 100 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > $f 2> ${f%.json}.err; summ $f
done
f=$out/bench_cps_new_c2.json
timeout 100 python bench.py --steps 10 --warmup 3 > $f 2> ${f%.json}.err; summ $f
el "bench done"
