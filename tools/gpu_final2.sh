#!/bin/bash
# Final GPU pass of round 1 on the committed default (new shared-memory swizzle): whole GPU suite, bench lines of every
# workload (c2 with the CPU baseline), the streaming-adjoint knob for the record, ncu launch list + full captures.
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
          "fwd sweep", round(r["forward_sweep"]["avg_launch_ms"],3), "frac", round(r["forward_sweep"]["frac"],3), "fp32 step", round(r["fp32"]["step"]["frac"],3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
}
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4 | tee $out/pytest_final2.log
el "pytest done"
timeout 150 python bench.py --steps 10 --warmup 3 > $out/bench_final2_c2.json 2> $out/bench_final2_c2.err; summ $out/bench_final2_c2.json
for wl in q20 c3 c1; do
  f=$out/bench_final2_$wl.json
  timeout 150 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > $f 2> ${f%.json}.err; summ $f
done
for wl in c2 q20; do
  f=$out/bench_final2_stream_$wl.json
  QB_ADJ_STREAM=1 timeout 150 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > $f 2> ${f%.json}.err; summ $f
done
el "bench done"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_final2.csv python tools/profile_step.py c2 4096 1 > $out/ncu_launch_final2.log 2>&1
el "launch list done"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:sweep_flat -s 37 -c 3 -o $out/prof_final2_bwd -f python tools/profile_step.py c2 4096 2 > $out/ncu_final2_bwd.log 2>&1
el "ncu adjoint done"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:sweep_flat -s 24 -c 2 -o $out/prof_final2_fwd -f python tools/profile_step.py c2 4096 2 > $out/ncu_final2_fwd.log 2>&1
el "ncu forward done"
