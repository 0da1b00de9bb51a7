#!/bin/bash
# Standard GPU pass of a build (run through gpurun on one B200):
#   bash tools/gpu_pass.sh [tag] [steps...]     steps: pytest bench full launches ncu  (default: pytest bench launches ncu)
# parity suite, the bench lines of config 2 / 20 qubits / config 3, the ncu launch list of one config-2 step and ncu --set full
# captures of the default adjoint and forward sweeps.  Every step has its own timeout and writes into gpurun_out/ as it goes.
tag=${1:-pass}; shift
steps=${@:-pytest bench launches ncu}
out=gpurun_out
mkdir -p $out
t0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - t0 ))s] $*"; }
summ() {
python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], round(d["value"],1), "evals/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["value"],1), "; adjoint", round(r["avg_launch_ms"],3), round(r["frac"],3),
          "fwd", round(r["forward_sweep"]["avg_launch_ms"],3), round(r["forward_sweep"]["frac"],3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[1], "ERR", e); print(open(sys.argv[1].replace(".json",".err")).read()[-800:])
PY
}
for s in $steps; do case $s in
pytest)
  rm -f $out/parity_errors.jsonl
  timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -25 | tee $out/${tag}_pytest.log
  el "pytest done";;
bench)
  for wl in c2 q20 c3; do
    f=$out/${tag}_bench_$wl.json
    timeout 200 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > $f 2> ${f%.json}.err; summ $f
  done
  el "bench done";;
full)  # the driver's command lines: default bench (secondary shapes + CPU baseline) and the reference arm
  f=$out/${tag}_bench_full.json
  timeout 600 python bench.py > $f 2> ${f%.json}.err; summ $f; python -c "
import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print({k:(v.get('value'),v.get('ms_per_step'),(v.get('roofline') or {}).get('frac'),(v.get('cpu_baseline') or {}).get('value'),v.get('cuda_graph')) for k,v in d.get('secondary',{}).items()}); print(d.get('cpu_baseline'))"
  timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref_c2.json 2> $out/${tag}_bench_ref_c2.err; cut -c1-400 $out/${tag}_bench_ref_c2.json
  timeout 100 python bench.py --impl reference --workload c1 --steps 200 --warmup 5 > $out/${tag}_bench_ref_c1.json 2> $out/${tag}_bench_ref_c1.err; cut -c1-400 $out/${tag}_bench_ref_c1.json
  el "full bench + reference arm done";;
host)  # where the host time of a config-1 step goes
  timeout 120 python tools/profile_host.py c1 50 > $out/${tag}_host_profile_c1.txt 2>&1; head -45 $out/${tag}_host_profile_c1.txt | cut -c1-200
  el "host profile done";;
launches)
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_c2.csv python tools/profile_step.py c2 4096 1 > $out/${tag}_launches.log 2>&1
  el "launch list done";;
ncu)
  # config 2, second step (20 sweep launches per step with the default plan: 10 forward, 10 adjoint): three adjoint sweeps from the
  # middle of the backward, two forward sweeps from the middle of the forward
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_flat -s 33 -c 3 -o $out/${tag}_prof_bwd -f python tools/profile_step.py c2 4096 2 > $out/${tag}_ncu_bwd.log 2>&1
  el "ncu adjoint done"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_flat -s 24 -c 2 -o $out/${tag}_prof_fwd -f python tools/profile_step.py c2 4096 2 > $out/${tag}_ncu_fwd.log 2>&1
  el "ncu forward done";;
esac; done
