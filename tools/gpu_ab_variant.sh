#!/bin/bash
# A/B of a second build of the native libraries (qandle_b200/_variants/<name>, loaded with QB_LIB_DIR): the whole GPU suite on
# the variant, then bench lines default / variant (and variant + streaming adjoint).   usage: bash tools/gpu_ab_variant.sh swz765
name=${1:-swz765}
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
          "fwd sweep", round(r["forward_sweep"]["avg_launch_ms"],3), "frac", round(r["forward_sweep"]["frac"],3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
}
V=$PWD/qandle_b200/_variants/$name
QB_LIB_DIR=$V timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4 | tee $out/pytest_variant_$name.log
el "pytest variant done"
for wl in c2 q20 c3; do
  f=$out/bench_${name}_base_$wl.json
  timeout 150 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > $f 2> ${f%.json}.err; summ $f
  f=$out/bench_${name}_var_$wl.json
  QB_LIB_DIR=$V timeout 150 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > $f 2> ${f%.json}.err; summ $f
  f=$out/bench_${name}_var_stream_$wl.json
  QB_LIB_DIR=$V QB_ADJ_STREAM=1 timeout 150 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > $f 2> ${f%.json}.err; summ $f
done
el "bench done"
