timeout 420 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for v in 0 1; do
  QB_FULL_TILE=$v timeout 300 python tools/run_sharded.py --qubits 27 --layers 10 --dtype c128 --backward --out gpurun_out/c128_27q_full_$v.json > gpurun_out/c128_27q_full_$v.log 2>&1
done
QB_FULL_TILE=1 timeout 300 python tools/run_sharded.py --qubits 27 --layers 10 --dtype c128 --backward --check gpurun_out/c128_27q_full_0.json --out gpurun_out/c128_27q_full_1b.json > gpurun_out/c128_27q_full_1b.log 2>&1
python - <<'PY'
import json
for v in ("0","1","1b"):
    try:
        d=json.load(open(f"gpurun_out/c128_27q_full_{v}.json"))
        print(v, {k:d[k] for k in d if k in ("forward_ms","backward_ms","sweeps","max_abs_diff_probs","max_abs_diff_grads","check_probs","check_grad")}, [k for k in d if "check" in k or "diff" in k])
    except Exception as e: print(v,"ERR",e)
PY
tail -3 gpurun_out/c128_27q_full_1b.log
