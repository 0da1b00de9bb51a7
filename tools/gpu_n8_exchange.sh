#!/bin/bash
# 8 GPUs: the push exchange (TMA bulk copies into peer staging) on config 4 and config 5; compare with the bench's p2p numbers.
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
short() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k:v for k,v in d.items() if k not in ('probs','grads','grad_head','probs_head')})"; }
echo "== C4 push"; timeout 200 $T --master-port 29631 tools/run_sharded.py --qubits 30 --layers 50 --dtype c128 --backward --exchange push --out gpurun_out/r2_c4_w8_push.json 2>&1 | short
echo "== C4 p2p"; timeout 200 $T --master-port 29632 tools/run_sharded.py --qubits 30 --layers 50 --dtype c128 --backward --exchange p2p --out gpurun_out/r2_c4_w8_p2p.json --check gpurun_out/r2_c4_w8_push.json 2>&1 | short
echo "== C5 push pieces 8"; timeout 300 $T --master-port 29633 tools/run_sharded.py --qubits 36 --layers 10 --dtype c64 --backward --exchange push --pieces 8 --reps 1 --out gpurun_out/r2_c5_w8_push.json 2>&1 | short
