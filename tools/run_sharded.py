"""BASELINE configs 4 / 5: one large state, amplitude-sharded over the ranks of a torchrun launch.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/run_sharded.py \
        --qubits 30 --layers 50 --dtype c128 [--backward] [--pieces 4] [--out gpurun_out/c4_w8.json] [--check other.json]

Circuit: L x [one of RX/RY/RZ per qubit (random.choice, seed 0); CZ brickwork even then odd] +
MeasureProbability, weights U[0, 2pi) (torch.manual_seed(0)), start |0...0>.  N = 1 runs the same code unsharded.
Prints one JSON line on rank 0: timings (CUDA events, max over ranks), probabilities, gradient checksum, and -- with
--check -- the deviation from another run's probabilities (single GPU vs sharded: oracle tier T3).
"""
import argparse
import json
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def build_layers(q, n, L, seed=0):
    rng = random.Random(seed)
    layers = []
    for _ in range(L):
        for k in range(n):
            layers.append(getattr(q, rng.choice(["RX", "RY", "RZ"]))(k, remapping=None))
        if os.environ.get("QB_EXPERIMENT_NO_CZ") != "1":  # experiment knob: cost of the CZ brickwork
            for k in list(range(0, n - 1, 2)) + list(range(1, n - 1, 2)):
                layers.append(q.CZ(k, k + 1))
    layers.append(q.MeasureProbability())
    return layers


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", dest="n", type=int, default=30)
    ap.add_argument("--layers", type=int, default=50)
    ap.add_argument("--dtype", default="c128")
    ap.add_argument("--backward", action="store_true")
    ap.add_argument("--pieces", type=int, default=1)
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--out", default="")
    ap.add_argument("--check", default="")
    ap.add_argument("--exchange", default="auto", choices=["auto", "nccl", "p2p", "push"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    import qandle_b200 as q
    from qandle_b200.distributed import ShardedCircuit

    real = torch.float64 if args.dtype == "c128" else torch.float32
    torch.manual_seed(0)
    sc = ShardedCircuit(build_layers(q, args.n, args.layers), num_qubits=args.n, pieces=args.pieces, exchange=args.exchange)
    with torch.no_grad():
        for p in sc.parameters():
            p.mul_(2 * 3.141592653589793)
    sc = sc.to(dev)
    torch.manual_seed(2)
    g = torch.randn(args.n, device=dev, dtype=real)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    times_f, times_b = [], []
    out = None
    for rep in range(args.reps + 1):  # first iteration = warm-up (plan build, NCCL setup)
        for p in sc.parameters():
            p.grad = None
        barrier()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        out = sc(dtype=real)
        e1.record()
        if args.backward:
            out.backward(g)
        e2.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1), e1.elapsed_time(e2)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rep > 0 or args.reps == 0:
            times_f.append(float(t[0]))
            times_b.append(float(t[1]))
    probs = out.detach().double().cpu()
    n_ex = sum(1 for s in sc.step_types if s == 1)
    ex_ms = sc.exchange_time_ms(real) if world > 1 else None
    res = {
        "exchange_ms": ex_ms,
        "exchange_GBps": ((2**args.n) * (16 if args.dtype == "c128" else 8) / world * (world - 1) / world / (ex_ms / 1e3) / 1e9) if ex_ms else None,
        "n": args.n, "layers": args.layers, "dtype": args.dtype, "world": world, "pieces": args.pieces, "exchange": sc.exchange,
        "sweeps": sc.plan.num_sweeps, "exchanges": n_ex, "gates": len(sc.seg.rows),
        "forward_ms": min(times_f) if times_f else None, "backward_ms": min(times_b) if (times_b and args.backward) else None,
        "probs": probs.tolist(), "probs_in_unit_interval": bool((probs > -1e-5).all() and (probs < 1 + 1e-5).all()),
        "max_mem_GiB": torch.cuda.max_memory_allocated() / 2**30,
    }
    state_bytes = (2**args.n) * (16 if args.dtype == "c128" else 8)
    if res["forward_ms"]:
        res["forward_algorithmic_GBps_per_gpu"] = sc.plan.num_sweeps * 2 * state_bytes / world / (res["forward_ms"] / 1e3) / 1e9
    if args.backward:
        gr = torch.stack([p.grad.detach().double().reshape(()) for p in sc.parameters()]).cpu()
        res["grad_l2"] = float(gr.norm())
        res["grad_head"] = gr[:8].tolist()
        res["grads"] = gr.tolist()
    if args.check and os.path.exists(args.check) and rank == 0:
        ref = json.load(open(args.check))
        res["max_abs_prob_diff_vs_" + os.path.basename(args.check)] = float((probs - torch.tensor(ref["probs"], dtype=torch.float64)).abs().max())
        if args.backward and "grads" in ref:
            res["max_abs_grad_diff_vs_" + os.path.basename(args.check)] = float((gr - torch.tensor(ref["grads"], dtype=torch.float64)).abs().max())
    if rank == 0:
        if args.out:
            os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
            json.dump(res, open(args.out, "w"))
        short = {k: v for k, v in res.items() if k not in ("probs", "grads")}
        short["probs_head"] = res["probs"][:4]
        print(json.dumps(short))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
