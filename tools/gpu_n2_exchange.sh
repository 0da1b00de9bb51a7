mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_sharded.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -6
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
short() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k:v for k,v in d.items() if k not in ('probs','grads','grad_head','probs_head')})"; }
echo "== push TMA"; timeout 200 $T --master-port 29621 tools/run_sharded.py --qubits 30 --layers 50 --dtype c128 --backward --exchange push --out gpurun_out/r2_c4_w2_push_tma.json 2>&1 | short
echo "== push ldst"; QB_EXCHANGE_TMA=0 timeout 200 $T --master-port 29622 tools/run_sharded.py --qubits 30 --layers 50 --dtype c128 --backward --exchange push --out gpurun_out/r2_c4_w2_push_ldst.json --check gpurun_out/r2_c4_w2_push_tma.json 2>&1 | short
echo "== p2p"; timeout 200 $T --master-port 29623 tools/run_sharded.py --qubits 30 --layers 50 --dtype c128 --backward --exchange p2p --out gpurun_out/r2_c4_w2_p2p.json --check gpurun_out/r2_c4_w2_push_tma.json 2>&1 | short
echo "== push TMA pieces 4, 32q c64 (8 GiB shards)"; timeout 200 $T --master-port 29624 tools/run_sharded.py --qubits 31 --layers 10 --dtype c64 --backward --exchange push --pieces 4 --out gpurun_out/r2_q31_w2_push.json 2>&1 | short
