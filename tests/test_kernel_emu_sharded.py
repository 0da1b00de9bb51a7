"""Amplitude-sharded plans on the kernel emulator (tests/kernel_emu): the REAL sweep / measurement / seed kernels run with a
non-zero rank (rank bits = out-of-shard index bits), every rank of a 2- or 4-way sharding simulated one after the other in this
process through the step-level C ABI (qb_prepare_dev, qb_init_zero_dev, qb_apply_forward_dev, qb_measure_probs_dev, qb_seed_probs_dev,
qb_apply_backward_dev, qb_finalize_grads_dev), the exchange steps done in numpy (chunk c of rank r <-> chunk r of rank c: what
distributed.exchange_inplace / qb_exchange_p2p_dev do over NCCL / NVLink).  Checked against the float64 oracle.

TEST INFRASTRUCTURE (see test_kernel_emu.py).  Complements tests/test_gpu_sharded.py, which needs >= 2 GPUs."""
import ctypes
import random

import numpy as np
import pytest
import torch

from oracle import statevec as O
from test_kernel_emu import _PlanOpts, _load, _rel


@pytest.fixture(scope="module")
def emu():
    return _load()


def _layers(n, depth, seed):
    """BASELINE configs 4 / 5 structure: random RX/RY/RZ per qubit, CZ brickwork; plus a CNOT onto qubit 0 (a rank bit) every second layer."""
    rng = random.Random(seed)
    rows, slot = [], 0
    for d in range(depth):
        for k in range(n):
            rows.append((rng.choice([O.OP_RX, O.OP_RY, O.OP_RZ]), k, -1, slot))
            slot += 1
        for k in list(range(0, n - 1, 2)) + list(range(1, n - 1, 2)):
            rows.append((O.OP_CZ, k, k + 1, 0))
        if d % 2 == 1:
            rows.append((O.OP_CNOT, n - 1, 0, 0))
    return rows, slot


def _exchange(shards, world, bits, f64):
    """in place on the per-rank raw arrays [B][2^n_local * 2 reals]: rank bit j <-> local index bit bits[j].  The data is in the
    engine's internal layout: complex128 interleaved (one amplitude per 16-byte vector), complex64 pack-planar (two amplitudes --
    index bit 0 -- per 16-byte vector), so the exchange moves whole vectors and never looks inside them."""
    B = shards[0].shape[0]
    shift = 0 if f64 else 1
    per = 2 if f64 else 4  # reals per vector
    v = [s.reshape(B, -1, per) for s in shards]
    nv = v[0].shape[1]
    x = np.arange(nv)
    vb = [b - shift for b in bits]
    assert min(vb) >= 0
    dest = np.zeros_like(x)
    for j, b in enumerate(vb):
        dest |= ((x >> b) & 1) << j
    old = [a.copy() for a in v]
    for r in range(world):
        xr = x.copy()
        for j, b in enumerate(vb):
            xr = (xr & ~(1 << b)) | (((r >> j) & 1) << b)
        for c in range(world):
            sel = dest == c
            v[r][:, x[sel], :] = old[c][:, xr[sel], :]


@pytest.mark.parametrize("any_bit", [0, 1])
@pytest.mark.parametrize("n,world,depth,real", [(14, 2, 3, torch.float32), (15, 4, 3, torch.float32), (13, 2, 3, torch.float64)])
def test_sharded_plan_step_by_step_matches_oracle(emu, n, world, depth, real, any_bit):
    lib = emu
    f64 = real == torch.float64
    rt = np.float64 if f64 else np.float32
    g_bits = world.bit_length() - 1
    n_local = n - g_bits
    rows, n_slots = _layers(n, depth, seed=n + world)
    gen = torch.Generator().manual_seed(n)
    thetas = (torch.rand(n_slots, generator=gen, dtype=torch.float64) * 6.283).to(real)
    prog = np.ascontiguousarray(np.asarray(rows, dtype=np.int32).reshape(-1, 4))
    po = _PlanOpts(n_local=n_local, final_layout=1, exchange_any_bit=any_bit)
    plan = ctypes.c_void_p()
    assert lib.qb_plan_create(prog.ctypes.data_as(ctypes.c_void_p), len(prog), n, 1 if f64 else 0, ctypes.byref(po), ctypes.byref(plan)) == 0, lib.qb_last_error()
    lib.qb_workspace_bytes.restype = ctypes.c_int64
    steps = [lib.qb_plan_step_type(plan, i) for i in range(lib.qb_plan_num_steps(plan))]
    assert 1 in steps, "the plan has no exchange step"
    B = 1
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    sh = np.ascontiguousarray(thetas.numpy())
    ws_bytes = lib.qb_workspace_bytes(plan, ctypes.c_int64(B))
    ws = [np.zeros(ws_bytes + 256, np.uint8) for _ in range(world)]
    wsp = [ctypes.c_void_p((w.ctypes.data + 255) & ~255) for w in ws]
    psi = [np.zeros((B, 2**n_local * 2), rt) for _ in range(world)]
    lam = [np.zeros((B, 2**n_local * 2), rt) for _ in range(world)]
    B64 = ctypes.c_int64(B)
    ok = lambda rc: (rc == 0) or pytest.fail(lib.qb_last_error().decode())
    stream_before = lib.qb_emu_stream_launches()
    for r in range(world):
        ok(lib.qb_prepare_dev(plan, B64, P(sh), None, 0, None, wsp[r], None))
        ok(lib.qb_init_zero_dev(plan, B64, P(psi[r]), r, None))
    # maximal runs of sweep steps, exchanges in between (distributed.ShardedDriver)
    runs, i = [], 0
    while i < len(steps):
        j = i
        while j < len(steps) and steps[j] == steps[i] == 0:
            j += 1
        runs.append(("sweeps", i, j) if steps[i] == 0 else ("exchange", i, i + 1))
        i = max(j, i + 1)

    def xbits(step):
        buf = (ctypes.c_int32 * 4)()
        g = lib.qb_plan_exchange_bits(plan, step, buf)
        assert g == g_bits
        return list(buf)[:g]

    if any_bit:
        assert any(xbits(s0) != list(range(n_local - g_bits, n_local)) for kind, s0, _ in runs if kind == "exchange"), "planner kept the top bits"
    for kind, s0, s1 in runs:
        if kind == "sweeps":
            for r in range(world):
                ok(lib.qb_apply_forward_dev(plan, s0, s1, B64, P(psi[r]), wsp[r], r, None))
        else:
            _exchange(psi, world, xbits(s0), f64)
    probs = np.zeros((B, n), rt)
    for r in range(world):
        part = np.zeros((B, n), rt)
        ok(lib.qb_measure_probs_dev(plan, B64, P(psi[r]), P(part), wsp[r], r, None))
        probs += part
    th64 = thetas.double().requires_grad_(True)
    ref = O.run_program(rows, n, th64, None, None, None, B, O.MEASURE_PROBS)
    tol = 1e-12 if f64 else 1e-5
    assert _rel(torch.from_numpy(probs).double(), ref.detach()) < tol
    # adjoint backward: the same runs in reverse, psi and lambda both exchanged
    g = torch.linspace(-1, 1, n, dtype=torch.float64).reshape(1, n)
    ref.backward(g)
    gn = np.ascontiguousarray(g.numpy().astype(rt))
    for r in range(world):
        ok(lib.qb_seed_probs_dev(plan, B64, P(psi[r]), P(gn), P(lam[r]), r, None))
        ok(lib.qb_backward_begin_dev(plan, B64, wsp[r], None))
    for kind, s0, s1 in reversed(runs):
        if kind == "sweeps":
            for r in range(world):
                ok(lib.qb_apply_backward_dev(plan, s0, s1, B64, P(psi[r]), P(lam[r]), wsp[r], r, None))
        else:
            _exchange(psi, world, xbits(s0), f64)
            _exchange(lam, world, xbits(s0), f64)
    grads = np.zeros(n_slots, rt)
    for r in range(world):
        gs = np.zeros(n_slots, rt)
        ok(lib.qb_finalize_grads_dev(plan, B64, P(sh), None, 0, None, wsp[r], P(gs), n_slots, None, None))
        grads += gs
    assert float(np.abs(grads - th64.grad.numpy()).max()) < tol * max(float(th64.grad.abs().max()), 1e-30) * (1 if f64 else 1)
    # un-computed state: |0...0> on rank 0, nothing elsewhere
    assert abs(psi[0][0, 0] - 1) < (1e-12 if f64 else 1e-5) and all(float(np.abs(psi[r]).max()) < 1e-5 for r in range(1, world))
    if not f64 and n_local >= 12 and not any_bit:
        assert lib.qb_emu_stream_launches() > stream_before, "sharded complex64 adjoint sweeps did not take the streaming kernel"
    lib.qb_plan_destroy(plan)
