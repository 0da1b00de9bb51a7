#!/bin/bash
# One GPU-box pass for the round's evidence: parity tests, bench lines, ncu launch list + full captures.
# usage (on the box, via gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $out/pytest_$tag.log
timeout 300 python bench.py --steps 10 --warmup 3 > $out/bench_${tag}_c2.json 2> $out/bench_${tag}_c2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_${tag}_c2_reference.json 2>> $out/bench_${tag}_c2.err
for wl in q20 c3 c1; do
  timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_${tag}_$wl.json 2> $out/bench_${tag}_$wl.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv python tools/profile_step.py c2 4096 1 > $out/ncu_launch_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_flat -s 24 -c 2 -o $out/prof_flat_fwd -f python tools/profile_step.py c2 4096 2 > $out/ncu_flat_fwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_flat -s 37 -c 3 -o $out/prof_flat_bwd -f python tools/profile_step.py c2 4096 2 > $out/ncu_flat_bwd.log 2>&1
cat $out/pytest_$tag.log
for f in $out/bench_${tag}_*.json; do echo $f; cut -c1-260 $f; done
