#!/bin/bash
# First GPU call of round 2: the experiment builds that round 1 validated on the kernel emulator only.
# Build them HERE first (they travel with the snapshot):
#   python -m qandle_b200.csrc.build --variant dyn -DQB_DYN_KERNELS
#   python -m qandle_b200.csrc.build --variant stream_loop -DQB_STREAM_LOOP
#   python -m qandle_b200.csrc.build --variant fuse_init -DQB_FUSE_INIT
#   python -m qandle_b200.csrc.build --variant fuse_seed -DQB_FUSE_SEED
#   python -m qandle_b200.csrc.build --variant fuse_probs -DQB_FUSE_PROBS
#   python -m qandle_b200.csrc.build --variant all -DQB_DYN_KERNELS -DQB_STREAM_LOOP -DQB_FUSE_INIT -DQB_FUSE_SEED -DQB_FUSE_PROBS
# then: gpurun --timeout 1500 -- 'bash tools/gpu_round2_open.sh'
# Per variant: the whole GPU parity suite on that build, then bench lines default / variant on config 2, 20 qubits, config 3.
out=gpurun_out
mkdir -p $out
t0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - t0 ))s] $*"; }
summ() {
python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], round(d["value"],1), "evals/s", round(d["ms_per_step"],2), "ms; adjoint", round(r["avg_launch_ms"],3), round(r["frac"],3),
          "fwd", round(r["forward_sweep"]["avg_launch_ms"],3), round(r["forward_sweep"]["frac"],3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[1], "ERR", e); print(open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
}
bench() {  # bench <tag> <workload> [ENV=VAL ...]
  local tag=$1 wl=$2; shift 2
  local f=$out/r2_open_${tag}_$wl.json
  env "$@" timeout 150 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > $f 2> ${f%.json}.err; summ $f
}
for wl in c2 q20 c3; do bench default $wl QB_NOOP=1; done
el "default bench done"
for name in dyn stream_loop fuse_init fuse_seed fuse_probs all; do
  V=$PWD/qandle_b200/_variants/$name
  [ -d $V ] || { echo "variant $name not built"; continue; }
  extra="QB_NOOP=1"
  case $name in dyn|all) extra="QB_DYN=1";; esac
  env QB_LIB_DIR=$V $extra timeout 420 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4 | tee $out/r2_open_pytest_$name.log
  el "pytest $name done"
  for wl in c2 q20 c3; do bench $name $wl QB_LIB_DIR=$V $extra; done
  el "bench $name done"
done
