#!/usr/bin/env python
"""Benchmark of the state-vector hot path (BASELINE.json metric: circuit evals/s, fwd+bwd, and HBM GB/s as a
fraction of the roofline).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c2|c3|q20|c1]

Default workload = BASELINE.json configs[1] ("c2"): 16-qubit strongly-entangling ansatz, 10 layers, batch 4096
per GPU, complex64, forward+backward, batch data-parallel (weak scaling; the only collective is the all-reduce
of the 480 weight gradients).  One "step" = zero_grad -> forward -> backward of the whole batch; one "eval" =
one batch element.  Prints ONE JSON line (rank 0).  Under torchrun (N > 1) every rank runs its own batch shard.

Besides the contract's keys the line carries
  secondary  (N = 1)  the other single-GPU shapes, measured in the same run: "q20" (the north_star's roofline shape: 20 qubits,
             batch 256), "c3" (BASELINE configs[2], batch 16) with value + roofline, and "c1" (configs[0], 4 qubits, batch 64:
             launch-latency bound) next to the UNMODIFIED reference timed on the host cores (oracle/_ref, kind "reference")
  sharded    (N > 1)  the amplitude-sharded path on the same N GPUs (strong scaling): BASELINE configs[3] (30 qubits, 50 layers,
             complex128) against the single-GPU run of the same circuit (speed-up, max deviation of probabilities and gradients),
             exchange bandwidth against NVLink, and -- at N = 8 -- configs[4] (36 qubits, complex64) with a light-cone oracle check.

--impl reference times the CPU implementation of the same path on the host cores, with all host threads: config 1 runs the
unmodified reference (oracle/_ref, made by oracle/make_ref.py in the build container); the other configs cannot be instantiated
by the reference (a 16-qubit gate is a 32 GiB matrix), so its port (oracle/statevec.py: same math, O(2^n) per
gate, torch CPU autograd) runs on a bounded sample of the workload.
"""
import argparse
import hashlib
import json
import os
import random
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
# amplitude-sharded single-state circuits: L x [random RX/RY/RZ per qubit; CZ brickwork even then odd] + MeasureProbability
SHARDED = {
    "c4": dict(n=30, layers=50, dtype="c128", desc="30q single state, 50 layers, complex128 (BASELINE configs[3])"),
    "c5": dict(n=36, layers=10, dtype="c64", desc="36q single state, 10 layers, complex64 (BASELINE configs[4])"),
}
NVLINK_GBPS = 770.0  # measured peer copy per direction (B200_PROFILING.md); 900 nominal
PI2 = 2 * 3.141592653589793


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


def ncu_traffic(workload, batch):
    """roofline.traffic: dram__bytes_read + dram__bytes_write per launch of the dominant kernel.  DRAM counters exist only under ncu, so
    the number comes from the committed capture profiles/ncu_traffic.json (tools/ncu_traffic.py) -- and only while the kernel sources
    hash to what that capture ran; otherwise null."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")
    try:
        t = json.load(open(path))
        h = hashlib.sha256()
        for f in t["kernel_sources"]:
            h.update(open(os.path.join(os.path.dirname(path), "..", "qandle_b200", "csrc", f), "rb").read())
        if t["workload"] != workload or t["batch"] != batch:
            return {"traffic": None, "traffic_note": f"profiles/ncu_traffic.json holds the capture of {t['workload']} batch {t['batch']}"}
        if h.hexdigest()[:16] != t["kernel_sources_sha16"]:
            return {"traffic": None, "traffic_note": "profiles/ncu_traffic.json is from an older kernel build (source hash differs): re-capture"}
        return {"traffic": t["dram_bytes_per_launch"],
                "traffic_note": f"{t['capture']}: mean dram__bytes_read.sum + dram__bytes_write.sum of {len(t['launches'])} launches of "
                                f"{t['kernel']}, kernel sources sha {t['kernel_sources_sha16']} = this build (profiles/ncu_traffic.json)"}
    except Exception as e:  # noqa: BLE001
        return {"traffic": None, "traffic_note": f"no usable profiles/ncu_traffic.json ({type(e).__name__})"}


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


# ---------------------------------------------------------------------------------------------------------------------
# CPU arms (the only places that execute oracle/)
def host_threads():
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1: undo that for the CPU arm)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def cpu_port_time(wl, sample_batch, steps, warmup, budget_s=25.0):
    """Time the oracle port (CPU torch, all host threads) on a bounded sample of the workload: fwd+bwd."""
    from oracle import statevec as O

    n = wl["n"]
    rows, n_w = oracle_rows(wl)
    torch.manual_seed(0)
    w = (torch.rand(n_w) * PI2).requires_grad_(True)
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


def reference_c1_time(steps=200, warmup=5, budget_s=20.0):
    """The UNMODIFIED reference (oracle/_ref: /root/reference/src/qandle + the qw_map stand-in) on config 1, fwd+bwd on the
    host cores (reference qcircuit.py:163-174).  Returns (evals/s, timed steps) or None when oracle/_ref does not exist."""
    from oracle import make_ref

    ref = make_ref.import_reference()
    if ref is None:
        return None
    wl = WORKLOADS["c1"]
    torch.manual_seed(0)
    circ = build_circuit(ref, wl)
    torch.manual_seed(1)
    x = torch.rand(wl["batch"], wl["n"], requires_grad=True)
    torch.manual_seed(2)
    g = torch.randn(wl["batch"], wl["n"])
    times = []
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        circ.zero_grad()
        x.grad = None
        t0 = time.perf_counter()
        out = circ(x=x)
        out.backward(g)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_start > budget_s and len(times) >= 10:
            break
    return wl["batch"] / statistics.median(times), len(times)


def cpu_baseline_for(name, steps=5, warmup=1, budget_s=25.0):
    """cpu_baseline object of a workload: the reference itself where it can run (config 1), its port on a bounded sample otherwise."""
    wl = WORKLOADS[name]
    cores = host_threads()
    if name == "c1":
        r = reference_c1_time(budget_s=budget_s)
        if r is not None:
            return {"value": r[0], "unit": "evals/s", "cores": cores, "kind": "reference",
                    "sample": f"unmodified reference (oracle/_ref) on config 1 as specified: batch 64, fwd+bwd, {r[1]} timed steps (median)"}
    sample = 64 if wl["n"] <= 4 else (8 if wl["n"] <= 16 else (2 if wl["n"] <= 20 else 1))
    evals, nst = cpu_port_time(wl, sample, steps, warmup, budget_s=budget_s)
    return {"value": evals, "unit": "evals/s", "cores": cores, "kind": "port",
            "sample": f"oracle/statevec.py (O(2^n)/gate restatement of the reference, torch CPU autograd) on batch {sample} of the same circuit, fwd+bwd, {nst} timed steps"}


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    cb = cpu_baseline_for(args.workload, steps=args.steps, warmup=args.warmup, budget_s=150.0)
    evals = cb["value"]
    line = {
        "impl": "reference", "metric": "circuit evals/sec (fwd+bwd)", "value": evals, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * wl["batch"] / evals, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "complex64 (f32 arithmetic)", "data": "synthetic",
        "config": {"workload": wl["desc"], "sample": cb["sample"],
                   "note": "ms_per_step = per-GPU batch / measured evals/s (the CPU arm runs on rank 0 only, with all host threads)"},
        "cpu_baseline": cb,
        "e2e": {"value": evals, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self, rank, world, dev):
        self.rank, self.world, self.dev = rank, world, dev

    def barrier(self):
        if self.world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world > 1:
            t = torch.tensor([v], device=self.dev, dtype=torch.float64)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            return float(t.item())
        return v


def measure_workload(ctx, name, B, steps, warmup, with_e2e=True, with_clocks=True):
    """One batch-data-parallel workload on this rank's GPU: device-resident throughput, e2e through the public API with
    host buffers, and the roofline of the dominant kernel (adjoint sweep) timed alone with CUDA events."""
    import qandle_b200 as q
    from qandle_b200 import distributed as qdist
    from qandle_b200 import config, engine, qcircuit

    wl = WORKLOADS[name]
    n = wl["n"]
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    torch.manual_seed(0)
    circ = build_circuit(q, wl)
    with torch.no_grad():
        for p in circ.parameters():
            p.mul_(PI2)  # random-angle circuits
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

    def timed(fn, k):
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        ctx.barrier()
        return ctx.max_over_ranks(e0.elapsed_time(e1))

    for _ in range(warmup):
        step_resident()
    sampler = ClockSampler(dev.index or 0)
    if rank == 0 and with_clocks:
        sampler.start()
    ms = timed(step_resident, steps)
    clocks = sampler.stop() if (rank == 0 and with_clocks) else None
    res = {"value": world * B * steps / (ms / 1000.0), "ms_per_step": ms / steps, "clocks": clocks}

    if with_e2e:
        for _ in range(2):
            step_e2e()
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step_e2e()
        ctx.barrier()
        e2e_s = ctx.max_over_ranks(time.perf_counter() - t0)
        res["e2e"] = {"value": world * B * steps / e2e_s, "unit": "evals/s", "h2d_bytes_per_step": int(2 * B * n * 4),
                      "d2h_bytes_per_step": int(2 * B * n * 4)}

    seg = qcircuit.segments_of(circ.circuit)[0]
    plan = next(iter(seg.plans.values()))
    res["gpu_launches"] = int((plan.launches_fwd + plan.launches_bwd) * steps)
    res["config"] = {"workload": wl["desc"], "n_qubits": n, "per_gpu_batch": B, "gates": len(seg.rows), "weights": len(params),
                     "parallelism": f"batch-dp{world}", "sweeps": plan.num_sweeps, "tile_bits": min(config.ENGINE_TILE_BITS or 12, n),
                     "l2": f"state+adjoint working set {2 * B * (2**n) * 8 / 2**30:.3f} GiB per GPU"
                           + (" >> 126 MB L2 (no flush needed)" if 2 * B * (2**n) * 8 > 4 * 126e6 else " (fits L2: launch-latency bound, no HBM roofline)")}
    if n < 12:
        return res  # config 1: 128-byte states, launch-latency bound -- no sweep roofline to report

    # ---- roofline of the dominant kernel (adjoint sweep), timed alone with CUDA events on the launch stream ------------
    ops = torch.ops.qandle_b200
    ws = torch.empty(ops.workspace_bytes(plan.handle, B) + 256, dtype=torch.uint8, device=dev)
    shared = qcircuit._gather_weights(seg, dev, torch.float32).detach().contiguous()
    batch_angles = x.detach().contiguous()
    mats = torch.zeros(0, device=dev)
    state = torch.empty(B, 2**n, dtype=torch.complex64, device=dev)
    lam = torch.empty_like(state)
    ops.prepare(plan.handle, B, shared, batch_angles, mats, ws)
    ops.init_zero(plan.handle, B, state, 0)
    ops.apply_forward(plan.handle, 0, plan.num_steps, B, state, ws, 0)
    ops.seed_probs(plan.handle, B, state, g.contiguous(), lam, 0)
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
    step_s = ms / steps / 1000.0
    fp32 = {"unit": "TFLOP/s", "peak": fp32_peak, "peak_source": f"{sm_count} SMs x 128 FP32 lanes x 2 x {sm_mhz:.0f} MHz (CUDA cores; not a tensor-core figure)",
            "fused_2x2_ops": n_u1, "fused_diag_ops": n_d1,
            "adjoint_sweeps": {"achieved": flop_bwd / (bwd_ms / 1000.0) / 1e12, "frac": flop_bwd / (bwd_ms / 1000.0) / 1e12 / fp32_peak},
            "forward_sweeps": {"achieved": flop_fwd / (fwd_ms / 1000.0) / 1e12, "frac": flop_fwd / (fwd_ms / 1000.0) / 1e12 / fp32_peak},
            "step": {"achieved": (flop_fwd + flop_bwd) / step_s / 1e12, "frac": (flop_fwd + flop_bwd) / step_s / 1e12 / fp32_peak},
            "hbm_time_over_fp32_time": ((bytes_fwd + bytes_bwd) / (peak * 1e9)) / ((flop_fwd + flop_bwd) / (fp32_peak * 1e12))}
    res["roofline"] = {
        "bound": "hbm", "limiter": "fp32-issue" if fp32["hbm_time_over_fp32_time"] < 1 else "hbm",
        "kernel": "fl::sweep_flat_kernel<BWD, FULL, STREAM, DYN> (streaming adjoint sweep, persistent CTAs)", "achieved": achieved, "peak": peak,
        "unit": "GB/s", "frac": achieved / peak,
        **ncu_traffic(name, B),
        "peak_source": peak_src,
        "note": "algorithmic bytes = 4 x state bytes x batch per adjoint-sweep launch (read + write of psi and lambda), launch time = CUDA events over "
                "all adjoint sweeps of the plan re-run back to back / launches; with maximal fusion a sweep applies many fused gates, so the "
                "FP32 floor of the work is above the HBM time whenever hbm_time_over_fp32_time < 1 (fp32 object; DESIGN.md 4)",
        "algorithmic_bytes_per_launch": bytes_bwd / max(plan.num_sweeps, 1), "launches_per_step": plan.num_sweeps,
        "avg_launch_ms": bwd_ms / max(plan.num_sweeps, 1),
        "forward_sweep": {"achieved": bytes_fwd / (fwd_ms / 1000.0) / 1e9, "frac": bytes_fwd / (fwd_ms / 1000.0) / 1e9 / peak,
                          "avg_launch_ms": fwd_ms / max(plan.num_sweeps, 1)},
        "step_algorithmic_GBps": (bytes_fwd + bytes_bwd) / step_s / 1e9, "step_frac": (bytes_fwd + bytes_bwd) / step_s / 1e9 / peak,
        "unfused_equivalent_GBps": res["value"] / world * unfused_bytes_per_eval / 1e9,
        "fp32": fp32}
    return res


def measure_graphed(ctx, name, steps, warmup=5):
    """The same step through qandle_b200.cuda_graph (forward and adjoint backward replayed from CUDA graphs): device-resident and
    e2e with host buffers (H2D of the inputs, D2H of the result and of the input gradient every step)."""
    import qandle_b200 as q

    wl = WORKLOADS[name]
    n, B, dev = wl["n"], wl["batch"], ctx.dev
    torch.manual_seed(0)
    circ = build_circuit(q, wl)
    with torch.no_grad():
        for p in circ.parameters():
            p.mul_(PI2)
    circ = circ.to(dev)
    params = list(circ.parameters())
    torch.manual_seed(1)
    x_host = torch.rand(B, n).pin_memory()
    torch.manual_seed(2)
    g_host = torch.randn(B, n).pin_memory()
    x = x_host.to(dev).requires_grad_(True)
    g = g_host.to(dev)
    fast = q.cuda_graph(circ, x=x)

    def step():
        for p in params:
            p.grad = None
        x.grad = None
        out = fast(x=x)
        out.backward(g)
        return out

    def step_e2e():
        for p in params:
            p.grad = None
        x.grad = None
        with torch.no_grad():
            x.copy_(x_host, non_blocking=True)
            g.copy_(g_host, non_blocking=True)
        out = fast(x=x)
        out.backward(g)
        return out.detach().cpu(), x.grad.cpu()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    for _ in range(3):
        step_e2e()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) / steps * 1e3
    return {"value": B / (ms / 1e3), "unit": "evals/s", "ms_per_step": ms,
            "e2e": {"value": B / (e2e_ms / 1e3), "unit": "evals/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(2 * B * n * 4),
                    "d2h_bytes_per_step": int(2 * B * n * 4)},
            "note": "qandle_b200.cuda_graph(circuit, x=example): one CUDA-graph launch forward, one backward"}


# ---------------------------------------------------------------------------------------------------------------------
# amplitude-sharded leg (N > 1)
def sharded_layers(q, n, L, seed=0):
    rng = random.Random(seed)
    layers, rows = [], []
    for _ in range(L):
        for k in range(n):
            kind = rng.choice(["RX", "RY", "RZ"])
            layers.append(getattr(q, kind)(k, remapping=None))
            rows.append((kind, k, -1))
        for k in list(range(0, n - 1, 2)) + list(range(1, n - 1, 2)):
            layers.append(q.CZ(k, k + 1))
            rows.append(("CZ", k, k + 1))
    layers.append(q.MeasureProbability())
    return layers, rows


def light_cone_oracle(rows, n, thetas, g):
    """Exact P(q = 0) of EVERY qubit, and the gradient of sum_q g_q P_q w.r.t. every angle, of a shallow random-rotation + CZ
    circuit on n qubits (n = 36 here) from the CPU oracle: P_q only depends on the gates in qubit q's backward light cone (the
    rest cancels as U^+ U), which for a few brickwork layers spans <= ~15 qubits.  Checker only (oracle/statevec.py)."""
    from oracle import statevec as O

    th = thetas.detach().double().clone().requires_grad_(True)
    slot_of, s = {}, 0
    for i, r in enumerate(rows):
        if r[0] != "CZ":
            slot_of[i] = s
            s += 1
    probs = []
    for qt in range(n):
        cone, keep = {qt}, []
        for i in range(len(rows) - 1, -1, -1):
            kind, a, b = rows[i]
            if kind == "CZ":
                if a in cone or b in cone:
                    cone.update((a, b))
                    keep.append(i)
            elif a in cone:
                keep.append(i)
        keep.reverse()
        qs = sorted(cone)
        pos = {qq: j for j, qq in enumerate(qs)}
        prog = []
        for i in keep:
            kind, a, b = rows[i]
            if kind == "CZ":
                prog.append((O.OP_CZ, pos[a], pos[b], 0))
            else:
                prog.append(({"RX": O.OP_RX, "RY": O.OP_RY, "RZ": O.OP_RZ}[kind], pos[a], -1, slot_of[i]))
        p = O.run_program(prog, len(qs), th, None, None, None, 1, O.MEASURE_PROBS)
        probs.append(p[0, pos[qt]])
    probs = torch.stack(probs)
    (probs * g.double().cpu()).sum().backward()
    return probs.detach(), th.grad


def run_sharded_case(ctx, name, layers_override=None, verify_single=False, verify_light_cone=False, reps=1):
    import qandle_b200 as q
    from qandle_b200.distributed import ShardedCircuit

    cfg = SHARDED[name]
    n, L = cfg["n"], layers_override or cfg["layers"]
    real = torch.float64 if cfg["dtype"] == "c128" else torch.float32
    sz = 16 if cfg["dtype"] == "c128" else 8
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    torch.manual_seed(0)
    layers, rows = sharded_layers(q, n, L)
    shard_gib = (2**n) * sz / world / 2**30
    pieces = 1 if shard_gib <= 16 else 8  # 36 qubits: psi + lambda leave no room for full-size exchange staging
    sc = ShardedCircuit(layers, num_qubits=n, pieces=pieces, exchange=os.environ.get("QB_BENCH_EXCHANGE", "auto"))
    with torch.no_grad():
        for p in sc.parameters():
            p.mul_(PI2)
    sc = sc.to(dev)
    torch.manual_seed(2)
    g = torch.randn(n, device=dev, dtype=real)
    tf, tb = [], []
    out = None
    for rep in range(reps + 1):  # first iteration = warm-up (plan build, NCCL / symmetric-memory setup)
        for p in sc.parameters():
            p.grad = None
        ctx.barrier()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        out = sc(dtype=real)
        e1.record()
        out.backward(g)
        e2.record()
        ctx.barrier()
        f, b = ctx.max_over_ranks(e0.elapsed_time(e1)), ctx.max_over_ranks(e1.elapsed_time(e2))
        if rep > 0:
            tf.append(f)
            tb.append(b)
    probs = out.detach().double().cpu()
    grads = torch.stack([p.grad.detach().double().reshape(()) for p in sc.parameters()]).cpu()
    thetas = torch.stack([p.detach().reshape(()) for p in sc.parameters()]).cpu()
    n_ex = sum(1 for s in sc.step_types if s == 1)
    peak, _ = peaks()
    shard_bytes = (2**n) * sz / world
    fwd_ms, bwd_ms = min(tf), min(tb)
    ex_ms = sc.exchange_time_ms(real)
    ex_bytes = shard_bytes * (world - 1) / world  # sent (= received) per GPU and exchange
    res = {"workload": cfg["desc"] if not layers_override else cfg["desc"].replace(f"{cfg['layers']} layers", f"{L} layers"),
           "n_qubits": n, "layers": L, "gates": len(rows), "dtype": cfg["dtype"], "world": world, "exchange": sc.exchange, "pieces": pieces,
           "shard_GiB": shard_gib, "sweeps": sc.plan.num_sweeps, "exchanges": n_ex,
           "fwd_ms": fwd_ms, "bwd_ms": bwd_ms, "evals_per_s": 1000.0 / (fwd_ms + bwd_ms),
           "hbm_frac_fwd": sc.plan.num_sweeps * 2 * shard_bytes / (fwd_ms / 1e3) / 1e9 / peak,
           "hbm_frac_bwd": sc.plan.num_sweeps * 4 * shard_bytes / (bwd_ms / 1e3) / 1e9 / peak,
           "exchange_ms": ex_ms, "exchange_GBps": ex_bytes / (ex_ms / 1e3) / 1e9, "nvlink_frac": ex_bytes / (ex_ms / 1e3) / 1e9 / NVLINK_GBPS,
           "exchange_share_of_fwd": n_ex * ex_ms / fwd_ms,
           "probs_in_unit_interval": bool((probs > -1e-5).all() and (probs < 1 + 1e-5).all()),
           "max_mem_GiB": torch.cuda.max_memory_allocated(dev) / 2**30}
    del sc, out
    torch.cuda.empty_cache()
    if verify_single:
        # the same circuit unsharded on rank 0 (the single-GPU engine is itself checked against the oracle at <= 24 qubits by the
        # GPU parity suite: oracle tier T3)
        if rank == 0:
            torch.manual_seed(0)
            layers1, _ = sharded_layers(q, n, L)
            c1 = q.Circuit(layers=layers1, num_qubits=n)
            with torch.no_grad():
                for p in c1.parameters():
                    p.mul_(PI2)
            c1 = c1.to(dev)
            st0 = torch.zeros(2**n, dtype=torch.complex128 if real == torch.float64 else torch.complex64, device=dev)
            st0[0] = 1
            single_ms = None
            for rep in range(2):  # first iteration = warm-up (plan build: ~0.6 s of host time with the sweep-size search), like the sharded run
                for p in c1.parameters():
                    p.grad = None
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                o1 = c1(st0)
                o1.backward(g.to(o1.dtype))
                e1.record()
                torch.cuda.synchronize()
                single_ms = e0.elapsed_time(e1)
            p1 = o1.detach().double().cpu()
            g1 = torch.stack([p.grad.detach().double().reshape(()) for p in c1.parameters()]).cpu()
            res.update({"single_gpu_fwd_bwd_ms": single_ms, "speedup_vs_1": single_ms / (fwd_ms + bwd_ms),
                        "max_abs_diff_vs_single_gpu": {"probs": float((probs - p1).abs().max()), "grads": float((grads - g1).abs().max())}})
            del c1, o1, st0
            torch.cuda.empty_cache()
        ctx.barrier()
    if verify_light_cone and rank == 0:
        p_ref, g_ref = light_cone_oracle(rows, n, thetas, g.cpu())
        res["light_cone_oracle"] = {"max_abs_prob_diff": float((probs - p_ref).abs().max()),
                                    "max_abs_grad_diff": float((grads - g_ref).abs().max()),
                                    "grad_scale": float(g_ref.abs().max()),
                                    "note": "exact marginals and gradients of all qubits from the CPU oracle on each qubit's backward light cone"}
    return res


def sharded_leg(ctx):
    out = {"scaling": "strong", "nvlink_peak_GBps": NVLINK_GBPS}
    try:
        out["c4"] = run_sharded_case(ctx, "c4", verify_single=True)
    except Exception as e:  # noqa: BLE001
        out["c4"] = {"error": repr(e)[:300]}
    if ctx.world >= 8:
        try:
            out["c5_check"] = run_sharded_case(ctx, "c5", layers_override=3, verify_light_cone=True, reps=1)
            out["c5"] = run_sharded_case(ctx, "c5", reps=1)
        except Exception as e:  # noqa: BLE001
            out["c5"] = {"error": repr(e)[:300]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default=os.environ.get("QB_WORKLOAD", "c2"))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary single-GPU shapes / the sharded leg")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    ctx = Ctx(rank, world, dev)
    B = args.batch or wl["batch"]
    m = measure_workload(ctx, args.workload, B, args.steps, args.warmup)
    line = {
        "metric": "circuit evals/sec (fwd+bwd)", "value": m["value"], "unit": "evals/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "complex64 (f32 arithmetic)", "data": "synthetic", "config": m["config"], "e2e": m["e2e"],
        "gpu_launches": m["gpu_launches"], "roofline": m.get("roofline"), "clocks": m["clocks"],
    }
    if not args.no_secondary:
        if world == 1:
            sec = {}
            for name in ("q20", "c3", "c1"):
                if name == args.workload:
                    continue
                try:
                    s = measure_workload(ctx, name, WORKLOADS[name]["batch"], max(3, min(args.steps, 5)) if name != "c1" else 200, 3, with_e2e=(name == "c1"))
                    entry = {"value": s["value"], "unit": "evals/s", "ms_per_step": s["ms_per_step"], "config": s["config"], "clocks": s["clocks"]}
                    if "roofline" in s:
                        entry["roofline"] = s["roofline"]
                    if "e2e" in s:
                        entry["e2e"] = s["e2e"]
                    if name == "c1":
                        try:
                            entry["cuda_graph"] = measure_graphed(ctx, "c1", 200)
                        except Exception as e:  # noqa: BLE001
                            entry["cuda_graph"] = {"error": repr(e)[:300]}
                    if name == "c1" and not args.no_cpu_baseline:
                        entry["cpu_baseline"] = cpu_baseline_for("c1", budget_s=10.0)
                    sec[name] = entry
                except Exception as e:  # noqa: BLE001
                    sec[name] = {"error": repr(e)[:300]}
                torch.cuda.empty_cache()
            # the same workload on the planner's plain greedy plan (no sweep-size search, 256-byte chunks): more, lighter sweeps -- a HIGHER
            # per-launch HBM fraction at a LOWER throughput; shows which way the two move
            from qandle_b200 import config as _cfg

            try:
                if not isinstance(line.get("roofline"), dict):
                    raise RuntimeError("no sweep roofline for this workload")
                _cfg.ENGINE_SWEEP_SEARCH = False
                s = measure_workload(ctx, args.workload, B, max(3, min(args.steps, 5)), 3, with_e2e=False, with_clocks=False)
                line["roofline"]["greedy_plan"] = {"value": s["value"], "unit": "evals/s", "ms_per_step": s["ms_per_step"], "sweeps": s["config"]["sweeps"],
                                                   "frac": s["roofline"]["frac"], "avg_launch_ms": s["roofline"]["avg_launch_ms"],
                                                   "forward_sweep_frac": s["roofline"]["forward_sweep"]["frac"],
                                                   "note": "QB_SWEEP_SEARCH=0: greedy sweep fill; the default plan does the same gates in fewer, heavier sweeps"}
            except Exception as e:  # noqa: BLE001
                if isinstance(line.get("roofline"), dict):
                    line["roofline"]["greedy_plan"] = {"error": repr(e)[:300]}
            finally:
                _cfg.ENGINE_SWEEP_SEARCH = True
            torch.cuda.empty_cache()
            line["secondary"] = sec
        else:
            line["sharded"] = sharded_leg(ctx)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_for(args.workload)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
