#!/bin/bash
# Last GPU pass of round 1 (tight budget): parity tests of the committed state, the default bench line, the A/B of the
# interleaved adjoint reduction (QB_ADJ_INTERLEAVE=1: parity subset + bench), the issue-model micro-benchmark with the
# scalar-broadcast FFMA2 form, then (if time is left) refreshed ncu captures.  Every step has its own timeout and writes
# into gpurun_out/ as it goes, so a clamped call still leaves results.   usage: bash tools/gpu_final.sh
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
          "fwd sweep", round(r["forward_sweep"]["avg_launch_ms"],3), "fp32 step frac", round(r.get("fp32",{}).get("step",{}).get("frac",0),3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
}
el start
timeout 450 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -5 | tee $out/pytest_final.log
el "pytest default done"
timeout 150 python bench.py --steps 10 --warmup 3 > $out/bench_final_c2.json 2> $out/bench_final_c2.err; summ $out/bench_final_c2.json
el "bench c2 default done"
QB_ADJ_INTERLEAVE=1 timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "random_circuits or config2 or expectation or norm_preserved or ir_golden" 2>&1 | tail -3 | tee $out/pytest_interleave.log
el "pytest interleave done"
QB_ADJ_INTERLEAVE=1 timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/bench_final_c2_interleave.json 2> $out/bench_final_c2_interleave.err; summ $out/bench_final_c2_interleave.json
el "bench c2 interleave done"
[ -x tools/ubench/issue_model ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/issue_model tools/ubench/issue_model.cu
timeout 60 tools/ubench/issue_model > $out/ubench_issue_model.txt 2>&1; grep -A5 "scalar-broadcast" $out/ubench_issue_model.txt | head -6
el "ubench done"
for v in 0 1; do
  QB_ADJ_INTERLEAVE=$v timeout 120 python bench.py --workload q20 --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_final_q20_il$v.json 2> $out/bench_final_q20_il$v.err; summ $out/bench_final_q20_il$v.json
done
el "bench q20 done"
timeout 150 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_final_c3.json 2> $out/bench_final_c3.err; summ $out/bench_final_c3.json
el "bench c3 done"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_final.csv python tools/profile_step.py c2 4096 1 > $out/ncu_launch_final.log 2>&1
el "launch list done"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:sweep_flat -s 37 -c 3 -o $out/prof_final_bwd -f python tools/profile_step.py c2 4096 2 > $out/ncu_final_bwd.log 2>&1
el "ncu adjoint done"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:sweep_flat -s 24 -c 2 -o $out/prof_final_fwd -f python tools/profile_step.py c2 4096 2 > $out/ncu_final_fwd.log 2>&1
el "ncu forward done"
