#!/bin/bash
# sweep engine options on the bench workload: tools/tune.sh c2
wl=${1:-c2}
for tb in 10 11 12 13; do for lb in 4 5; do
  echo -n "tile_bits=$tb low_bits=$lb: "
  QB_TILE_BITS=$tb QB_LOW_BITS=$lb timeout 120 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), 'evals/s', round(d['ms_per_step'],1),'ms sweeps',d['config']['sweeps'],'fwd/bwd ms per sweep',round(d['roofline']['forward_sweep']['avg_launch_ms'],2),round(d['roofline']['avg_launch_ms'],2))"
done; done
echo -n "unstaged: "; QB_STAGED=0 timeout 120 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-120
