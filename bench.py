#!/usr/bin/env python
"""Benchmark of the state-vector hot path (BASELINE.json metric: circuit evals/s, fwd+bwd, and HBM GB/s as a
fraction of the roofline).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c2|c3|q20]

Default workload = BASELINE.json configs[1] ("c2"): 16-qubit strongly-entangling ansatz, 10 layers, batch 4096
per GPU, complex64, forward+backward, batch data-parallel (weak scaling; the only collective is the all-reduce
of the 480 weight gradients).  One "step" = zero_grad -> forward -> backward of the whole batch; one "eval" =
one batch element.  Prints ONE JSON line (rank 0).  Under torchrun (N > 1) every rank runs its own batch shard.

--impl reference times the CPU implementation of the same path on the host cores: the reference itself is pure
Python and cannot travel to the GPU box nor instantiate a 16-qubit gate (32 GiB per matrix), so
the arm runs the oracle port (oracle/statevec.py: same math, O(2^n) per gate, torch CPU autograd) on a bounded
sample of the same workload, with all host threads.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (n_qubits, depth, per-GPU batch, description)
    "c1": dict(n=4, depth=1, batch=64, desc="4q README-style hybrid layer (AngleEmbedding + RX/RY + CNOT ring + MeasureProbability), batch 64, c64, fwd+bwd"),
    "c2": dict(n=16, depth=10, batch=4096, desc="16q strongly-entangling ansatz x10, AngleEmbedding, MeasureProbability, batch 4096/GPU, c64, fwd+bwd"),
    "q20": dict(n=20, depth=10, batch=256, desc="20q strongly-entangling ansatz x10, AngleEmbedding, MeasureProbability, batch 256/GPU, c64, fwd+bwd"),
    "c3": dict(n=24, depth=20, batch=16, desc="24q hardware-efficient ansatz x20 (RY,RZ + CNOT chain), batch 16/GPU, c64, fwd+bwd"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                self.samples.append(float(parts[0]))
                self.max_mhz = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(nm)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        s = self.samples
        return {"sm_mhz": statistics.median(s) if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def build_circuit(q, wl):
    n, depth = wl["n"], wl["depth"]
    if wl is WORKLOADS["c1"]:
        layers = [q.AngleEmbedding(name="x", qubits=list(range(n)))] + [q.RX(k) for k in range(n)] + [q.RY(k) for k in range(n)]
        layers += [q.CNOT(k, (k + 1) % n) for k in range(n)] + [q.MeasureProbability()]
    elif wl is WORKLOADS["c3"]:
        layers = [q.AngleEmbedding(name="x", qubits=list(range(n)))]
        for _ in range(depth):
            layers += [q.RY(k, remapping=None) for k in range(n)] + [q.RZ(k, remapping=None) for k in range(n)]
            layers += [q.CNOT(k, k + 1) for k in range(n - 1)]
        layers.append(q.MeasureProbability())
    else:
        layers = [q.AngleEmbedding(name="x", qubits=list(range(n))),
                  q.StronglyEntanglingLayer(qubits=list(range(n)), depth=depth, remapping=None),
                  q.MeasureProbability()]
    return q.Circuit(layers=layers, num_qubits=n)


def oracle_rows(wl):
    from oracle import statevec as O

    n, depth = wl["n"], wl["depth"]
    rows = [(O.OP_RX | O.FLAG_BATCH, k, -1, k) for k in range(n)]
    if wl is WORKLOADS["c1"]:
        rows += [(O.OP_RX, k, -1, k) for k in range(n)] + [(O.OP_RY, k, -1, n + k) for k in range(n)]
        rows += [(O.OP_CNOT, k, (k + 1) % n, 0) for k in range(n)]
        return rows, 2 * n
    if wl is WORKLOADS["c3"]:
        s = 0
        for _ in range(depth):
            rows += [(O.OP_RY, k, -1, s + k) for k in range(n)] + [(O.OP_RZ, k, -1, s + n + k) for k in range(n)]
            s += 2 * n
            rows += [(O.OP_CNOT, k, k + 1, 0) for k in range(n - 1)]
        n_w = s
    else:
        rows += O.sel_program(list(range(n)), depth)
        n_w = depth * n * 3
    return rows, n_w


def cpu_port_time(wl, sample_batch, steps, warmup, budget_s=25.0):
    """Time the oracle port (CPU torch, all host threads) on a bounded sample of the workload: fwd+bwd."""
    from oracle import statevec as O

    n = wl["n"]
    rows, n_w = oracle_rows(wl)
    torch.manual_seed(0)
    w = (torch.rand(n_w) * 2 * 3.141592653589793).requires_grad_(True)
    torch.manual_seed(1)
    x = torch.rand(sample_batch, n).requires_grad_(True)
    torch.manual_seed(2)
    g = torch.randn(sample_batch, n)
    times = []
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        w.grad = None
        x.grad = None
        t0 = time.perf_counter()
        out = O.run_program(rows, n, w, x, None, None, sample_batch, O.MEASURE_PROBS)
        out.backward(g)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_start > budget_s and len(times) >= 1:
            break
    return sample_batch / statistics.median(times), len(times)


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    sample = 8 if wl["n"] <= 16 else (2 if wl["n"] <= 20 else 1)
    evals, nsteps = cpu_port_time(wl, sample, args.steps, args.warmup, budget_s=150.0)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "circuit evals/sec (fwd+bwd)", "value": evals, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": nsteps, "warmup": args.warmup, "ms_per_step": 1000.0 * sample / evals, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "complex64 (f32 arithmetic)", "data": "synthetic",
        "config": {"workload": wl["desc"], "sample": f"batch {sample} of the per-GPU batch {wl['batch']} per step"},
        "cpu_baseline": {"value": evals, "unit": "evals/s", "cores": cores, "kind": "port",
                         "sample": f"oracle/statevec.py (O(2^n)/gate restatement of the reference, torch CPU autograd), batch {sample}, {nsteps} timed steps"},
        "e2e": {"value": evals, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default=os.environ.get("QB_WORKLOAD", "c2"))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    import qandle_b200 as q
    from qandle_b200 import distributed as qdist
    from qandle_b200 import config, engine

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    n, B = wl["n"], (args.batch or wl["batch"])

    torch.manual_seed(0)
    circ = build_circuit(q, wl)
    with torch.no_grad():
        for p in circ.parameters():
            p.mul_(2 * 3.141592653589793)  # random-angle circuits
    circ = circ.to(dev)
    params = list(circ.parameters())
    torch.manual_seed(1 + rank)
    x_host = torch.rand(B, n).pin_memory()
    torch.manual_seed(2 + rank)
    g_host = torch.randn(B, n).pin_memory()
    x = x_host.to(dev).requires_grad_(True)
    g = g_host.to(dev)

    def step_resident():
        for p in params:
            p.grad = None
        x.grad = None
        out = circ(x=x)
        out.backward(g)
        if world > 1:
            qdist.allreduce_gradients(params)
        return out

    def step_e2e():
        # public API with HOST buffers: H2D of the inputs, D2H of the result and of the input gradient, every step
        for p in params:
            p.grad = None
        xh = x_host.detach().requires_grad_(True)
        out = circ(x=xh)  # CPU tensor in -> CPU tensor out
        out.backward(g_host)
        if world > 1:
            qdist.allreduce_gradients(params)
        return out, xh.grad

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed(step_resident, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms / 1000.0)

    for _ in range(2):
        step_e2e()
    t0 = time.perf_counter()
    barrier()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * B * args.steps / e2e_s

    # ---- roofline of the dominant kernel (backward sweep), timed alone with CUDA events on the launch stream ------
    seg = circ.circuit._qb_segments[0]
    plan = next(iter(seg.plans.values()))
    ops = torch.ops.qandle_b200
    ws = torch.empty(ops.workspace_bytes(plan.handle, B) + 256, dtype=torch.uint8, device=dev)
    from qandle_b200 import qcircuit

    shared = qcircuit._gather_weights(seg, dev, torch.float32).detach().contiguous()
    batch_angles = x.detach().contiguous()
    mats = torch.zeros(0, device=dev)
    state = torch.empty(B, 2**n, dtype=torch.complex64, device=dev)
    lam = torch.empty_like(state)
    ops.prepare(plan.handle, B, shared, batch_angles, mats, ws)
    ops.init_zero(plan.handle, B, state, 0)
    ops.apply_forward(plan.handle, 0, plan.num_steps, B, state, ws, 0)
    gprob = g.contiguous()
    ops.seed_probs(plan.handle, B, state, gprob, lam, 0)
    ops.backward_begin(plan.handle, B, ws)
    torch.cuda.synchronize()
    reps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # (unitary sweeps keep |psi| and |lambda| bounded, so the kernels can be re-run back to back)
    e0.record()
    for _ in range(reps):
        ops.apply_backward(plan.handle, 0, plan.num_steps, B, state, lam, ws, 0)
    e1.record()
    torch.cuda.synchronize()
    bwd_ms = e0.elapsed_time(e1) / reps
    e0.record()
    for _ in range(reps):
        ops.apply_forward(plan.handle, 0, plan.num_steps, B, state, ws, 0)
    e1.record()
    torch.cuda.synchronize()
    fwd_ms = e0.elapsed_time(e1) / reps
    bytes_bwd = plan.algorithmic_bytes(B, True)
    bytes_fwd = plan.algorithmic_bytes(B, False)
    peak, peak_src = peaks()
    traffic = None  # measured only under ncu (profiles/): never a constant carried in the bench
    achieved = bytes_bwd / (bwd_ms / 1000.0) / 1e9
    S = (2**n) * 8
    n_gates = len(seg.rows)
    unfused_bytes_per_eval = (6 * n_gates + 2) * S
    # FP32 companion of the HBM roofline: the gate arithmetic of the fused plan (FMA per amplitude: a 2x2 costs 8 forward
    # and 8 + 8 + 6 in the adjoint sweep (psi, lambda, Pauli sums); a diagonal phase 4 / 4 + 4 + 2) against the CUDA-core
    # FP32 peak = SMs x 128 lanes x 2 flop x the SM clock sampled during the timed region.
    dump = engine.parse_plan_dump(plan.dump().tolist())
    n_u1 = sum(1 for sw in dump["sweeps"] for o in sw["ops"] if o["kind"] == 1)
    n_d1 = sum(1 for sw in dump["sweeps"] for o in sw["ops"] if o["kind"] in (2, 3))
    amps = float(B) * 2**n
    flop_fwd = 2.0 * amps * (8 * n_u1 + 4 * n_d1)
    flop_bwd = 2.0 * amps * (22 * n_u1 + 10 * n_d1)
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    sm_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
    fp32_peak = sm_count * 128 * 2 * sm_mhz * 1e6 / 1e12
    fp32 = {"unit": "TFLOP/s", "peak": fp32_peak, "peak_source": f"{sm_count} SMs x 128 FP32 lanes x 2 x {sm_mhz:.0f} MHz (CUDA cores; not a tensor-core figure)",
            "fused_2x2_ops": n_u1, "fused_diag_ops": n_d1,
            "adjoint_sweeps": {"achieved": flop_bwd / (bwd_ms / 1000.0) / 1e12, "frac": flop_bwd / (bwd_ms / 1000.0) / 1e12 / fp32_peak},
            "forward_sweeps": {"achieved": flop_fwd / (fwd_ms / 1000.0) / 1e12, "frac": flop_fwd / (fwd_ms / 1000.0) / 1e12 / fp32_peak},
            "step": {"achieved": (flop_fwd + flop_bwd) / (ms / args.steps / 1000.0) / 1e12,
                     "frac": (flop_fwd + flop_bwd) / (ms / args.steps / 1000.0) / 1e12 / fp32_peak},
            "hbm_time_over_fp32_time": ((bytes_fwd + bytes_bwd) / (peak * 1e9)) / ((flop_fwd + flop_bwd) / (fp32_peak * 1e12))}

    line = {
        "metric": "circuit evals/sec (fwd+bwd)", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "complex64 (f32 arithmetic)", "data": "synthetic",
        "config": {"workload": wl["desc"], "n_qubits": n, "per_gpu_batch": B, "gates": n_gates, "weights": len(params),
                   "parallelism": f"batch-dp{world}", "l2": f"state+adjoint working set {2 * B * S / 2**30:.1f} GiB per GPU >> 126 MB L2 (no flush needed)",
                   "sweeps": plan.num_sweeps, "tile_bits": min(config.ENGINE_TILE_BITS or 12, n)},
        "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": int(2 * B * n * 4), "d2h_bytes_per_step": int(2 * B * n * 4)},
        "gpu_launches": int((plan.launches_fwd + plan.launches_bwd) * args.steps),
        "roofline": {"bound": "hbm", "kernel": "fl::sweep_flat_kernel<true> (adjoint sweep)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "note": "with maximal fusion a sweep applies ~27 fused gates: the FP32 floor of that work (FFMA2 at 2 issue cycles) is above the HBM time, so the sweeps are latency / FP bound, not HBM bound (ncu: profiles/r1_final_summary.md; DESIGN.md 4; the fp32 object quantifies it); unfused_equivalent_GBps is the gate-per-pass equivalent",
                     "algorithmic_bytes_per_launch": bytes_bwd / max(plan.num_sweeps, 1), "launches_per_step": plan.num_sweeps,
                     "avg_launch_ms": bwd_ms / max(plan.num_sweeps, 1),
                     "forward_sweep": {"achieved": bytes_fwd / (fwd_ms / 1000.0) / 1e9, "frac": bytes_fwd / (fwd_ms / 1000.0) / 1e9 / peak,
                                       "avg_launch_ms": fwd_ms / max(plan.num_sweeps, 1)},
                     "unfused_equivalent_GBps": value / world * unfused_bytes_per_eval / 1e9,
                     "fp32": fp32},
        "clocks": clocks,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = 8 if n <= 16 else (2 if n <= 20 else 1)
        cpu_evals, nst = cpu_port_time(wl, sample, 5, 1, budget_s=25.0)
        line["cpu_baseline"] = {"value": cpu_evals, "unit": "evals/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"oracle/statevec.py on batch {sample} (same circuit, fwd+bwd), {nst} timed steps"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
