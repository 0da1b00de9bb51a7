"""Per-sweep CUDA-event times of a plan on ONE GPU, with each sweep's tile bits and op counts.

   python tools/sweep_times.py --qubits 33 --layers 10 --dtype c64 [--rank-bits 3] [--backward] [--workload c2 --batch 4096]

With --rank-bits g the plan is the one rank 0 of a 2^g-GPU amplitude-sharded run executes on a shard of 2^(qubits) amplitudes (the circuit
has qubits + g qubits; exchange steps are skipped, so the numbers in the state are meaningless: this is a TIMING tool for the local sweeps of
configs 4 / 5 without the 8-GPU box).  Prints one line per sweep: tile bits, 2x2s / diagonal / controlled ops, ms, algorithmic GB/s."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import qandle_b200 as q
from qandle_b200 import config, engine, qcircuit


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=30, help="LOCAL qubits (the shard)")
    ap.add_argument("--layers", type=int, default=10)
    ap.add_argument("--dtype", default="c64")
    ap.add_argument("--rank-bits", type=int, default=0)
    ap.add_argument("--backward", action="store_true")
    ap.add_argument("--workload", default="")
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--out", default="")
    ap.add_argument("--low-bits", type=int, default=0, help="lowest tile bits that are always staged (0: the default, 5 complex64 / 4 complex128)")
    ap.add_argument("--phases", action="store_true", help="QB_LIB_DIR = a -DQB_PHASE_TIMES variant build: per-phase cycles of thread 0, per tile")
    args = ap.parse_args()
    dev = engine.require_cuda()
    ops = engine.load_ops()
    g = args.rank_bits
    if args.workload:
        wl = bench.WORKLOADS[args.workload]
        n, depth = wl["n"], wl["depth"]
        if args.workload == "c3":
            layers = [q.AngleEmbedding(name="x", qubits=list(range(n)))]
            for _ in range(depth):
                layers += [q.RY(k, remapping=None) for k in range(n)] + [q.RZ(k, remapping=None) for k in range(n)]
                layers += [q.CNOT(k, k + 1) for k in range(n - 1)]
            layers.append(q.MeasureProbability())
        else:
            layers = [q.AngleEmbedding(name="x", qubits=list(range(n))),
                      q.StronglyEntanglingLayer(qubits=list(range(n)), depth=depth, remapping=None), q.MeasureProbability()]
        c128 = False
        if args.batch == 1:
            args.batch = wl["batch"]
    else:
        n = args.qubits + g
        layers, _ = bench.sharded_layers(q, n, args.layers)
        c128 = args.dtype == "c128"
    real = torch.float64 if c128 else torch.float32
    circ = qcircuit.UnsplittedCircuit(n, layers)
    seg = qcircuit.lower_modules(circ.layers, n)[0]
    prog = torch.tensor(seg.rows, dtype=torch.int32).reshape(-1, 4)
    opts = (0, args.low_bits, 0, n - g, 0, 0, 1, config.ENGINE_MAX_OPS_PER_SWEEP, 0, 0, 0, 0, 1 if g else 0)
    plan = engine.Plan(prog, n, engine.C128 if c128 else engine.C64, opts)
    d = engine.parse_plan_dump(plan.dump().tolist())
    st = plan.step_types()
    B = args.batch
    torch.manual_seed(0)
    with torch.no_grad():
        for p in circ.parameters():
            p.mul_(bench.PI2)
    shared = qcircuit._gather_weights(seg, dev, real).detach().contiguous()
    nb = len(seg.batch_cols)
    batch_angles = torch.rand(B, nb, device=dev, dtype=real) if nb else torch.zeros(0, device=dev, dtype=real)
    mats = (torch.view_as_real(torch.stack(seg.mats).to(device=dev, dtype=torch.complex128 if c128 else torch.complex64)).contiguous()
            if seg.mats else torch.zeros(0, device=dev, dtype=real))
    ws = torch.empty(ops.workspace_bytes(plan.handle, B) + 256, dtype=torch.uint8, device=dev)
    ops.prepare(plan.handle, B, shared, batch_angles, mats, ws)
    cd = torch.complex128 if c128 else torch.complex64
    state = torch.empty(B, 2 ** (n - g), dtype=cd, device=dev)
    ops.init_zero(plan.handle, B, state, 0)
    lam = None
    if args.backward:
        lam = torch.empty_like(state)
        gr = torch.randn(B, n, device=dev, dtype=real)
    S = 2 ** (n - g) * (16 if c128 else 8) * B
    sweeps = d["sweeps"]
    rows = []
    sw_idx = [i for i, t in enumerate(st) if t == 0]

    lib = engine.core()

    def phases():
        import ctypes

        buf = (ctypes.c_ulonglong * 12)()
        assert lib.qb_debug_phase_times(buf, 1) == 0
        v = list(buf)
        tiles = max(v[6], 1)
        names = ["item_setup", "tile_tables", "tile_wait", "stages", "epilogue+store", "end_barrier"]
        d = {nm: round(v[i] / tiles) for i, nm in enumerate(names)}
        d["tile_offsets+cp.async_issue"] = round(v[8] / tiles)
        d["cta_cycles_per_tile"] = round(v[7] / tiles)
        return d

    def timed(fn):
        best = 1e30
        for _ in range(args.reps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    for k, s in enumerate(sw_idx):
        # repeated forward application of one sweep: every gate is unitary, so the state stays a state
        if args.phases:
            phases()
        ms = timed(lambda: ops.apply_forward(plan.handle, s, s + 1, B, state, ws, 0))
        sw = sweeps[k]
        kinds = [o["kind"] for o in sw["ops"]]
        rows.append(dict(sweep=k, step=s, tile_bits=sw["tile_bits"], u1=kinds.count(1), diag=kinds.count(2) + kinds.count(3),
                         ctl=len(kinds) - kinds.count(1) - kinds.count(2) - kinds.count(3), stages=len(sw["stages"]),
                         fwd_ms=round(ms, 3), fwd_GBps=round(2 * S / ms / 1e6, 1)))
        if args.phases:
            rows[-1]["fwd_phase_cycles_per_tile"] = phases()
    if args.backward:
        ops.seed_probs(plan.handle, B, state, gr.contiguous(), lam, 0)
        ops.backward_begin(plan.handle, B, ws)
        for k, s in reversed(list(enumerate(sw_idx))):
            if args.phases:
                phases()
            ms = timed(lambda: ops.apply_backward(plan.handle, s, s + 1, B, state, lam, ws, 0))
            if args.phases:
                rows[k]["bwd_phase_cycles_per_tile"] = phases()
            rows[k]["bwd_ms"] = round(ms, 3)
            rows[k]["bwd_GBps"] = round(4 * S / ms / 1e6, 1)
    for r in rows:
        print(json.dumps(r))
    tot_f = sum(r["fwd_ms"] for r in rows)
    summ = dict(n_local=n - g, rank_bits=g, dtype="c128" if c128 else "c64", batch=B, sweeps=len(rows), exchanges=sum(1 for t in st if t == 1),
                fwd_ms=round(tot_f, 2), fwd_GBps=round(len(rows) * 2 * S / tot_f / 1e6, 1))
    if args.backward:
        tot_b = sum(r["bwd_ms"] for r in rows)
        summ.update(bwd_ms=round(tot_b, 2), bwd_GBps=round(len(rows) * 4 * S / tot_b / 1e6, 1))
    print(json.dumps(summ))
    if args.out:
        json.dump(dict(summary=summ, sweeps=rows), open(args.out, "w"))


if __name__ == "__main__":
    main()
