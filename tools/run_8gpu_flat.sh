#!/bin/bash
# One 8-GPU box with the flat kernels: config 4 (complex128 flat) and config 5 (complex64 flat, n_local = 33) amplitude-sharded,
# checked against the committed single-GPU / interpreted-kernel results, + the batch-DP scaling point.  Outputs: gpurun_out/.
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== C4 30q c128 x50, 8 GPUs"; timeout 300 $T --nproc-per-node 8 --master-port 29611 tools/run_sharded.py --qubits 30 --layers 50 --dtype c128 --backward --pieces 2 --out gpurun_out/c4_w8_flat.json --check profiles/r1_c4_30q_c128_w1.json 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:v for k,v in d.items() if k not in ('probs','grads','grad_head')})"
echo "== C5 36q c64 x10, 8 GPUs, forward+backward"; timeout 500 $T --nproc-per-node 8 --master-port 29614 tools/run_sharded.py --qubits 36 --layers 10 --dtype c64 --backward --pieces 8 --reps 0 --out gpurun_out/c5_w8_flat.json --check profiles/r1_c5_w8_fwd.json 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:v for k,v in d.items() if k not in ('probs','grads','grad_head')})"
echo "== bench batch-DP N=8"; timeout 200 $T --nproc-per-node 8 --master-port 29613 bench.py --gpus 8 --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_r1flat_c2_n8.json; cut -c1-220 gpurun_out/bench_r1flat_c2_n8.json
