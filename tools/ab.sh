#!/bin/bash
# A/B of in-tree builds on the same GPU box: bash tools/ab.sh "<variant|default> ..." [workloads...]
# Prints forward / adjoint sweep totals (tools/sweep_times.py) of each build, round-robin, two rounds.
vs=$1; shift; wls=${@:-c2 q20 c3}
mkdir -p gpurun_out
run() {
  local v=$1 wl=$2
  if [ "$v" = default ]; then unset QB_LIB_DIR; else export QB_LIB_DIR=$PWD/qandle_b200/_variants/$v; fi
  timeout 200 python tools/sweep_times.py --workload $wl --backward 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', '$wl', 'fwd', d['fwd_ms'], 'bwd', d['bwd_ms'], 'sum', round(d['fwd_ms']+d['bwd_ms'],2))"
}
for wl in $wls; do for rep in 1 2; do for v in $vs; do run $v $wl; done; done; done
