#!/bin/bash
# One 8-GPU box: configs 4 / 5 amplitude-sharded + batch-DP scaling point.  Outputs under gpurun_out/.
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L | wc -l > gpurun_out/ngpu.txt
echo "== C4 30q c128 x50, 8 GPUs"; timeout 300 $T --nproc-per-node 8 --master-port 29611 tools/run_sharded.py --qubits 30 --layers 50 --dtype c128 --backward --pieces 2 --out gpurun_out/c4_w8.json --check profiles/r1_c4_30q_c128_w1.json 2>&1 | tail -1 | cut -c1-900
echo "== C5 36q c64 x10, 8 GPUs, forward"; timeout 400 $T --nproc-per-node 8 --master-port 29612 tools/run_sharded.py --qubits 36 --layers 10 --dtype c64 --pieces 8 --out gpurun_out/c5_w8_fwd.json 2>&1 | tail -2 | cut -c1-900
echo "== bench batch-DP N=8"; timeout 200 $T --nproc-per-node 8 --master-port 29613 bench.py --gpus 8 --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-400
echo "== C5 36q c64 x10, 8 GPUs, forward+backward"; timeout 500 $T --nproc-per-node 8 --master-port 29614 tools/run_sharded.py --qubits 36 --layers 10 --dtype c64 --backward --pieces 8 --reps 0 --out gpurun_out/c5_w8_bwd.json --check gpurun_out/c5_w8_fwd.json 2>&1 | tail -2 | cut -c1-900
echo "== C4 4 GPUs"; timeout 200 $T --nproc-per-node 4 --master-port 29615 tools/run_sharded.py --qubits 30 --layers 50 --dtype c128 --backward --pieces 2 --out gpurun_out/c4_w4.json --check profiles/r1_c4_30q_c128_w1.json 2>&1 | tail -1 | cut -c1-900
echo "== C4 2 GPUs"; timeout 200 $T --nproc-per-node 2 --master-port 29616 tools/run_sharded.py --qubits 30 --layers 50 --dtype c128 --backward --pieces 2 --out gpurun_out/c4_w2.json --check profiles/r1_c4_30q_c128_w1.json 2>&1 | tail -1 | cut -c1-900
echo "== sharded pytest"; timeout 200 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout=150 -p no:cacheprovider 2>&1 | tail -2
