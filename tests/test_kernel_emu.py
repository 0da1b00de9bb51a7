"""The REAL CUDA kernels, run on the CPU by the kernel emulator (tests/kernel_emu), against the oracle.

TEST INFRASTRUCTURE.  tests/kernel_emu compiles the unmodified sources of qandle_b200/csrc with g++ (a stand-in
<cuda_runtime.h>, kernel launches / dynamic shared memory / the eight inline-PTX sites rewritten textually) and runs
every CUDA thread as a cooperative fiber with real barrier, named-barrier, warp-barrier and shuffle semantics.  The same
C ABI (include/qandle_b200.h: qb_plan_create, qb_run_host, ...) is called through ctypes, so what is checked here is the
kernels' and launchers' LOGIC -- addressing tables, absorbed CNOT maps, the adjoint linearisation, the warp reductions,
the deterministic gradient reduction -- on a machine without a GPU.  It is not a product path: qandle_b200 never loads
this library (engine.py / `test_no_cpu_fallback`), and the GPU parity tests (`-m gpu`) remain the parity gate.
Tolerances as on the GPU (north_star): 1e-5 relative complex64, 1e-12 complex128.
"""
import ctypes
import os
import random
import shutil

import numpy as np
import pytest
import torch

from kernel_emu import build_emu
from oracle import statevec as O
from test_planner_emulation import rand_state, rand_unitaries, random_program


def _load(env=None, tag="default", defines=()):
    """Each knob setting gets its own copy of the library (the knobs are read once per loaded image)."""
    lib_path = build_emu.build(defines=defines, tag="_".join(d.lower() for d in defines))
    if env:
        alt = os.path.join(os.path.dirname(lib_path), f"libqandle_b200_emu_{tag}.so")
        if not os.path.exists(alt) or os.path.getmtime(alt) < os.path.getmtime(lib_path):
            shutil.copy(lib_path, alt)
        lib_path = alt
    old = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        lib = ctypes.CDLL(lib_path)
        lib.qb_last_error.restype = ctypes.c_char_p
        # the knobs are function-local statics evaluated on first use: run a tiny plan now, while the environment is set
        _run(lib, 3, 1, [(O.OP_RX, 0, -1, 0), (O.OP_CNOT, 0, 1, 0)], torch.tensor([0.3]), None, None, None, O.MEASURE_PROBS, torch.float32,
             torch.ones(1, 3))
        _run(lib, 12, 1, [(O.OP_RX, 0, -1, 0), (O.OP_CNOT, 0, 11, 0)], torch.tensor([0.3]), None, None, None, O.MEASURE_PROBS, torch.float32,
             torch.ones(1, 12))
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return lib


@pytest.fixture(scope="module")
def emu():
    return _load()


@pytest.fixture(scope="module")
def emu_dyn():
    """The persistent full-tile kernels with the grid capped at 2 CTAs (test hook QB_DYN_GRID): (sample, tile-subset) work items are
    claimed from the atomic queue inside one CTA."""
    return _load({"QB_DYN_GRID": "2"}, "dyn_grid2")


@pytest.fixture(scope="module")
def emu_unfused():
    """Test hook QB_FUSE=0: separate |0...0>, MeasureProbability and adjoint-seed passes."""
    return _load({"QB_FUSE": "0"}, "unfused")


@pytest.fixture(scope="module")
def emu_reverse_order():
    """The default build with the fibers of a CTA resumed in reverse thread order (KEMU_ORDER=reverse)."""
    return _load({"KEMU_ORDER": "reverse"}, "reverse")


@pytest.fixture(scope="module")
def emu_two_cta_adjoint():
    return _load({"QB_ADJ_STREAM": "0"}, "nostream")


@pytest.fixture(scope="module")
def emu_swizzle_identity():
    return _load(defines=("QB_SWIZZLE_IDENTITY",))


class _PlanOpts(ctypes.Structure):
    """qb_plan_opts (include/qandle_b200.h)"""
    _fields_ = [(k, ctypes.c_int32) for k in ("tile_bits", "low_bits", "fuse", "n_local", "host_only", "swap_relabel", "final_layout",
                                              "max_ops_per_sweep", "staged", "packed", "flat", "narrow_sync", "exchange_any_bit")] + [("reserved", ctypes.c_int32 * 3)]


def _run(lib, n, B, prog, shared, batch, mats_engine, init, measure, real, g, opts=None):
    """qb_run_host on host buffers.  Returns (out, grad_shared, grad_batch, grad_init) as torch tensors."""
    f64 = real == torch.float64
    rt, ct = (np.float64, np.complex128) if f64 else (np.float32, np.complex64)
    prog = np.ascontiguousarray(np.asarray(prog, dtype=np.int32).reshape(-1, 4))
    plan = ctypes.c_void_p()
    po = _PlanOpts(**opts) if opts else None
    rc = lib.qb_plan_create(prog.ctypes.data_as(ctypes.c_void_p), len(prog), n, 1 if f64 else 0, ctypes.byref(po) if po else None, ctypes.byref(plan))
    assert rc == 0, lib.qb_last_error()
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else None
    sh = np.ascontiguousarray(shared.detach().numpy().astype(rt)) if shared is not None and shared.numel() else None
    ba = np.ascontiguousarray(batch.detach().numpy().astype(rt)) if batch is not None else None
    fm = np.ascontiguousarray(mats_engine.numpy().astype(ct)) if mats_engine is not None else None
    ini = np.ascontiguousarray(init.detach().numpy().astype(ct)) if init is not None else None
    N = 2**n
    if measure == O.MEASURE_PROBS:
        out = np.zeros((B, n), rt)
    elif measure == O.MEASURE_JOINT:
        out = np.zeros((B, N), rt)
    else:
        out = None
    final = np.zeros((B, N), ct) if measure == O.MEASURE_STATE else None
    gn = None
    if g is not None:
        gn = np.ascontiguousarray(g.numpy().astype(ct if measure == O.MEASURE_STATE else rt))
    gs = np.zeros(sh.shape, rt) if sh is not None and g is not None else None
    gb = np.zeros(ba.shape, rt) if ba is not None and g is not None else None
    gi = np.zeros((B, N), ct) if ini is not None and g is not None else None
    rc = lib.qb_run_host(plan, ctypes.c_int64(B), P(sh), 0 if sh is None else sh.size, P(ba), 0 if ba is None else ba.shape[1], P(fm),
                         0 if fm is None else fm.shape[0], P(ini), measure, P(out), P(final), P(gn), P(gs), P(gb), P(gi))
    assert rc == 0, lib.qb_last_error()
    lib.qb_plan_destroy(plan)
    res = final if measure == O.MEASURE_STATE else out
    T = lambda a: torch.from_numpy(a) if a is not None else None
    return T(res), T(gs), T(gb), T(gi)


def _rel(a, b):
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)


def _case(lib, n, B, G, seed, measure, real, with_init=False, n_mats=2, opts=None):
    """Same construction as test_gpu_parity.test_random_circuits_vs_oracle: the float64 oracle is the truth."""
    rng = random.Random(seed)
    gen = torch.Generator().manual_seed(seed)
    prog = random_program(rng, n, G, 6, 2, n_mats)
    shared = ((torch.rand(6, generator=gen, dtype=torch.float64) - 0.5) * 6).requires_grad_(True)
    batch = ((torch.rand(B, 2, generator=gen, dtype=torch.float64) - 0.5) * 6).requires_grad_(True)
    mats = rand_unitaries(gen, n_mats) if n_mats else None
    init = rand_state(gen, B, n).requires_grad_(True) if with_init else None
    ref = O.run_program(prog, n, shared, batch, mats, init, B, measure)
    g = torch.randn(ref.shape, generator=gen, dtype=torch.float64)
    if ref.is_complex():
        g = torch.complex(g, torch.randn(ref.shape, generator=gen, dtype=torch.float64))
    ref.backward(g)
    out, gs, gb, gi = _run(lib, n, B, prog, shared, batch, mats, init, measure, real, g, opts)
    # north_star tolerances: 1e-5 relative (complex64), 1e-12 (complex128); the gradient tensors of one backward share one scale
    tol = 1e-12 if real == torch.float64 else 1e-5
    assert _rel(out.to(ref.dtype), ref.detach()) < tol
    gscale = max(float(shared.grad.abs().max()), float(batch.grad.abs().max()), float(init.grad.abs().max()) if with_init else 0.0, 1e-30)
    assert float((gs.double() - shared.grad).abs().max()) < tol * gscale
    assert float((gb.double() - batch.grad).abs().max()) < tol * gscale
    if with_init:  # qb_run_host returns torch's convention (2 dL/dpsi0*)
        assert float((gi.to(torch.complex128) - init.grad).abs().max()) < tol * gscale


CASES = [
    # n, B, gates, measure, dtype, init   -- 12 = one full flat tile; 13 / 14 = several tiles per state, out-of-tile controls
    (5, 3, 60, O.MEASURE_PROBS, torch.float32, False),
    (9, 2, 120, O.MEASURE_JOINT, torch.float32, True),
    (12, 2, 140, O.MEASURE_PROBS, torch.float32, False),
    (13, 1, 160, O.MEASURE_STATE, torch.float32, True),
    (14, 1, 160, O.MEASURE_PROBS, torch.float32, False),
    (6, 2, 60, O.MEASURE_PROBS, torch.float64, True),
    (11, 1, 120, O.MEASURE_JOINT, torch.float64, False),
    (12, 1, 140, O.MEASURE_PROBS, torch.float64, False),
]


@pytest.mark.parametrize("n,B,G,measure,real,with_init", CASES)
def test_real_kernels_on_cpu_emulator_match_oracle(emu, n, B, G, measure, real, with_init):
    _case(emu, n, B, G, 100 + n, measure, real, with_init)


def test_strongly_entangling_ansatz_on_emulator(emu):
    """The bench circuit's shape (AngleEmbedding + SEL + MeasureProbability) at 13 qubits: flat full-tile kernels, fused
    RZ RY RZ groups, the CNOT ring absorbed into the stage addressing, forward + adjoint."""
    n, B, depth = 13, 2, 2
    gen = torch.Generator().manual_seed(3)
    prog = [(O.OP_RX | O.FLAG_BATCH, k, -1, k) for k in range(n)] + O.sel_program(list(range(n)), depth)
    w = (torch.rand(depth * n * 3, generator=gen) * 6.283).requires_grad_(True)
    x = torch.rand(B, n, generator=gen).requires_grad_(True)
    ref = O.run_program(prog, n, w, x, None, None, B, O.MEASURE_PROBS)
    g = torch.randn(B, n, generator=gen)
    ref.backward(g)
    out, gs, gb, _ = _run(emu, n, B, prog, w, x, None, None, O.MEASURE_PROBS, torch.float32, g)
    assert _rel(out, ref.detach()) < 1e-5
    assert float((gs - w.grad).abs().max()) < 5e-5 * max(1.0, float(w.grad.abs().max()))
    assert float((gb - x.grad).abs().max()) < 5e-5 * max(1.0, float(x.grad.abs().max()))


def _sel_case(lib, n, B, depth, seed, extra=(), opts=None):
    gen = torch.Generator().manual_seed(seed)
    prog = [(O.OP_RX | O.FLAG_BATCH, k, -1, k) for k in range(n)] + O.sel_program(list(range(n)), depth) + list(extra)
    w = (torch.rand(depth * n * 3, generator=gen, dtype=torch.float64) * 6.283).requires_grad_(True)
    x = torch.rand(B, n, generator=gen, dtype=torch.float64).requires_grad_(True)
    ref = O.run_program(prog, n, w, x, None, None, B, O.MEASURE_PROBS)
    g = torch.randn(B, n, generator=gen, dtype=torch.float64)
    ref.backward(g)
    out, gs, gb, _ = _run(lib, n, B, prog, w, x, None, None, O.MEASURE_PROBS, torch.float32, g, opts)
    scale = max(1.0, float(w.grad.abs().max()), float(x.grad.abs().max()))
    assert _rel(out.double(), ref.detach()) < 1e-5
    assert float((gs.double() - w.grad).abs().max()) < 2e-5 * scale
    assert float((gb.double() - x.grad).abs().max()) < 2e-5 * scale


@pytest.mark.parametrize("n,B,depth", [(12, 2, 1), (12, 1, 2), (13, 2, 2), (14, 1, 3)])
def test_streaming_adjoint_kernel_matches_oracle(emu, n, B, depth):
    """The streaming adjoint kernel (flat64.cuh: run_stages_stream): lambda streamed from shared memory for the Pauli sums, re-loaded
    for its own 2x2s, second barrier in stages that move amplitudes between threads, in-place fix-up pre-pass."""
    before = emu.qb_emu_stream_launches()
    _sel_case(emu, n, B, depth, 40 + n, extra=[(O.OP_CZ, 0, n - 1, 0), (O.OP_CNOT, n - 1, 1, 0), (O.OP_SWAP, 2, n - 2, 0)])
    assert emu.qb_emu_stream_launches() > before, "the streaming kernel was not selected"


def test_streaming_adjoint_falls_back_for_parametrised_diagonals(emu):
    """Sweeps with a gradient-carrying diagonal (bare RZ) run on the default adjoint kernel; mixed programs stay correct."""
    _case(emu, 12, 2, 140, 31, O.MEASURE_PROBS, torch.float32, with_init=True)
    _case(emu, 13, 1, 160, 32, O.MEASURE_STATE, torch.float32)


@pytest.mark.parametrize("n,B,G,measure,real,with_init", [c for c in CASES if c[4] == torch.float32])
def test_alternative_swizzle_build_matches_oracle(emu_swizzle_identity, n, B, G, measure, real, with_init):
    """-DQB_SWIZZLE_IDENTITY (packed64.cuh: slot_off): the old GF(2)-linear fold of unit bits 3-5 into the bank-group bits, kept
    for A/B builds; every table, tile fill and drain goes through the same function, so results must not change."""
    _case(emu_swizzle_identity, n, B, G, 100 + n, measure, real, with_init)


def test_alternative_swizzle_sel_circuit(emu_swizzle_identity):
    _sel_case(emu_swizzle_identity, 13, 2, 2, 53, extra=[(O.OP_CZ, 0, 12, 0), (O.OP_CNOT, 12, 1, 0)])


@pytest.mark.parametrize("n,B,depth", [(12, 2, 1), (13, 2, 2)])
def test_two_cta_adjoint_kernel_matches_oracle(emu_two_cta_adjoint, n, B, depth):
    """Test hook QB_ADJ_STREAM=0: the generic adjoint sweep with psi and lambda both in registers (128 registers, 2 CTAs / SM), which also
    serves every sweep the streaming kernel does not take."""
    before = emu_two_cta_adjoint.qb_emu_stream_launches()
    _sel_case(emu_two_cta_adjoint, n, B, depth, 60 + n, extra=[(O.OP_CZ, 0, n - 1, 0), (O.OP_CNOT, n - 1, 1, 0)])
    assert emu_two_cta_adjoint.qb_emu_stream_launches() == before


def test_results_do_not_depend_on_thread_execution_order(emu, emu_reverse_order):
    """Between two barriers the emulator runs a CTA's threads one after the other; a store / load pair that needs a barrier it
    does not have shows up as a result that changes when that order is reversed.  Streaming adjoint kernel, several tiles,
    stages that move amplitudes between threads: bit-identical in both orders."""
    n, B, depth = 13, 1, 3
    gen = torch.Generator().manual_seed(91)
    prog = [(O.OP_RY | O.FLAG_BATCH, k, -1, k) for k in range(n)] + O.sel_program(list(range(n)), depth) + [(O.OP_CZ, 0, 12, 0), (O.OP_SWAP, 3, 9, 0)]
    w = torch.rand(depth * n * 3, generator=gen) * 6.283
    x = torch.rand(B, n, generator=gen)
    g = torch.randn(B, n, generator=gen)
    a = _run(emu, n, B, prog, w, x, None, None, O.MEASURE_PROBS, torch.float32, g)
    b = _run(emu_reverse_order, n, B, prog, w, x, None, None, O.MEASURE_PROBS, torch.float32, g)
    for u, v in zip(a[:3], b[:3]):
        assert torch.equal(u, v)
    _case(emu_reverse_order, 12, 2, 140, 112, O.MEASURE_PROBS, torch.float32)
    _case(emu_reverse_order, 13, 1, 160, 113, O.MEASURE_STATE, torch.float32, True)
    _case(emu_reverse_order, 12, 1, 140, 112, O.MEASURE_PROBS, torch.float64)


@pytest.mark.parametrize("n,B,depth", [(12, 5, 1), (13, 3, 2), (14, 2, 2)])
def test_persistent_cta_work_queue_matches_oracle(emu_dyn, n, B, depth):
    """Persistent CTAs with the grid capped at 2: the emulator runs CTAs one after the other, so the
    first CTA claims every work item beyond the second -- all sample switches (per-sample matrices reloaded, accumulators
    restarted, per-item partial sums) happen inside one CTA.  Several samples, several tiles per sample, per-sample embedding angles and their gradients."""
    _sel_case(emu_dyn, n, B, depth, 70 + n, extra=[(O.OP_CZ, 0, n - 1, 0), (O.OP_CNOT, n - 1, 1, 0)])


def test_persistent_cta_build_mixed_programs(emu_dyn):
    _case(emu_dyn, 13, 2, 160, 33, O.MEASURE_PROBS, torch.float32, with_init=True)
    _case(emu_dyn, 12, 3, 140, 34, O.MEASURE_JOINT, torch.float32)


KERNEL_FAMILIES = [
    # every sweep-kernel family behind the plan options of include/qandle_b200.h (the GPU suite's BIG list, smaller)
    ("interpreted packed complex64 stage bodies", 12, 2, 120, torch.float32, dict(flat=-1)),
    ("interpreted packed, partial tiles", 9, 2, 100, torch.float32, dict(tile_bits=5, low_bits=2, flat=-1)),
    ("scalar staged complex64", 11, 2, 120, torch.float32, dict(packed=-1)),
    ("generic one-pass-per-op kernels complex64", 10, 2, 100, torch.float32, dict(tile_bits=6, low_bits=2, staged=-1)),
    ("flat complex128, several tiles", 13, 1, 140, torch.float64, dict(tile_bits=7, low_bits=2)),
    ("interpreted staged complex128", 11, 2, 120, torch.float64, dict(flat=-1)),
    ("generic kernels complex128", 9, 2, 100, torch.float64, dict(tile_bits=5, low_bits=2, staged=-1)),
    ("flat complex64, CTA barriers only, no fusion, SWAPs move data", 13, 1, 140, torch.float32, dict(narrow_sync=-1, fuse=-1, swap_relabel=-1)),
    ("flat complex64, capped sweeps", 12, 2, 140, torch.float32, dict(max_ops_per_sweep=6)),
]


@pytest.mark.parametrize("name,n,B,G,real,opts", KERNEL_FAMILIES, ids=[k[0] for k in KERNEL_FAMILIES])
@pytest.mark.parametrize("measure", [O.MEASURE_PROBS, O.MEASURE_STATE])
def test_every_kernel_family_on_emulator(emu, name, n, B, G, real, opts, measure):
    _case(emu, n, B, G, 500 + n + measure, measure, real, with_init=True, opts=opts)


@pytest.mark.parametrize("n,B,depth", [(12, 2, 1), (13, 2, 2), (14, 1, 3)])
def test_streaming_adjoint_more_shapes_match_oracle(emu, n, B, depth):
    before = emu.qb_emu_stream_launches()
    _sel_case(emu, n, B, depth, 80 + n, extra=[(O.OP_CZ, 0, n - 1, 0), (O.OP_CNOT, n - 1, 1, 0), (O.OP_SWAP, 2, n - 2, 0)])
    assert emu.qb_emu_stream_launches() > before


@pytest.mark.parametrize("n,B,depth", [(12, 2, 1), (13, 2, 2), (14, 1, 3)])
def test_fused_zero_init_matches_oracle(emu, n, B, depth):
    """qb_forward_dev skips init_zero_kernel and the first flat sweep starts from tiles it
    builds in shared memory; 13 / 14 qubits = several tiles per state, only tile 0 holds the 1."""
    before = emu.qb_emu_fused_inits()
    _sel_case(emu, n, B, depth, 90 + n, extra=[(O.OP_CZ, 0, n - 1, 0), (O.OP_CNOT, n - 1, 1, 0), (O.OP_SWAP, 2, n - 2, 0)])
    assert emu.qb_emu_fused_inits() > before, "the init pass was not skipped"


def test_fused_zero_init_mixed_programs_and_fallbacks(emu):
    """Random programs from |0...0> (fused) for all three measurements; partial tiles; a caller-supplied state, complex128 and the
    non-flat kernel families keep the separate init pass."""
    before = emu.qb_emu_fused_inits()
    _case(emu, 12, 2, 140, 41, O.MEASURE_PROBS, torch.float32)
    _case(emu, 13, 1, 160, 42, O.MEASURE_STATE, torch.float32)
    _case(emu, 14, 1, 120, 43, O.MEASURE_JOINT, torch.float32)
    _case(emu, 9, 2, 100, 44, O.MEASURE_PROBS, torch.float32, opts=dict(tile_bits=5, low_bits=2))
    assert emu.qb_emu_fused_inits() >= before + 3
    mid = emu.qb_emu_fused_inits()
    _case(emu, 12, 2, 140, 45, O.MEASURE_PROBS, torch.float32, with_init=True)
    _case(emu, 12, 1, 120, 46, O.MEASURE_PROBS, torch.float64)
    _case(emu, 11, 2, 120, 47, O.MEASURE_PROBS, torch.float32, opts=dict(flat=-1))
    _case(emu, 10, 2, 100, 48, O.MEASURE_PROBS, torch.float32, opts=dict(tile_bits=6, low_bits=2, staged=-1))
    assert emu.qb_emu_fused_inits() == mid



@pytest.mark.parametrize("n,B,depth", [(12, 2, 1), (13, 2, 2), (14, 1, 3)])
def test_fused_adjoint_seed_matches_oracle(emu, emu_two_cta_adjoint, n, B, depth):
    """qb_backward_dev skips seed_probs_kernel; the first adjoint sweep (streaming kernel, and
    the two-CTA kernel with QB_ADJ_STREAM=0) builds lambda = w (.) psi in shared memory.  Several tiles per state at 13 / 14
    qubits (out-of-tile weights), permuted final layout."""
    # final_layout=1 as the engine plans MeasureProbability segments (qcircuit.py): the SWAP stays a relabelling and the final layout a
    # permutation; with the identity layout restored, the restoring tail sweep runs on the generic kernel and keeps the seed pass
    extra = [(O.OP_CZ, 0, n - 1, 0), (O.OP_SWAP, 2, n - 2, 0), (O.OP_CNOT, n - 1, 1, 0)]
    for lib in (emu, emu_two_cta_adjoint):
        before = lib.qb_emu_fused_seeds()
        _sel_case(lib, n, B, depth, 60 + n, extra=extra, opts=dict(final_layout=1))
        assert lib.qb_emu_fused_seeds() > before, "the seed pass was not skipped"


def test_fused_adjoint_seed_mixed_programs_and_fallbacks(emu):
    before = emu.qb_emu_fused_seeds()
    perm = dict(final_layout=1)
    _case(emu, 12, 2, 140, 51, O.MEASURE_PROBS, torch.float32, with_init=True, opts=perm)  # bare RZ: tile dots, two-CTA kernel
    _case(emu, 13, 2, 160, 52, O.MEASURE_PROBS, torch.float32, opts=perm)
    _case(emu, 14, 1, 120, 53, O.MEASURE_PROBS, torch.float32, with_init=True, opts=perm)
    _case(emu, 9, 2, 100, 54, O.MEASURE_PROBS, torch.float32, opts=dict(tile_bits=5, low_bits=2, final_layout=1))  # partial tiles
    _case(emu, 13, 1, 140, 55, O.MEASURE_PROBS, torch.float32, opts=dict(max_ops_per_sweep=6, final_layout=1))
    assert emu.qb_emu_fused_seeds() >= before + 4
    mid = emu.qb_emu_fused_seeds()
    _case(emu, 12, 2, 140, 56, O.MEASURE_STATE, torch.float32)
    _case(emu, 12, 2, 140, 57, O.MEASURE_JOINT, torch.float32)
    _case(emu, 12, 1, 120, 58, O.MEASURE_PROBS, torch.float64)
    _case(emu, 11, 2, 120, 59, O.MEASURE_PROBS, torch.float32, opts=dict(flat=-1))
    assert emu.qb_emu_fused_seeds() == mid


@pytest.mark.parametrize("n,B,depth", [(12, 3, 1), (13, 3, 2), (14, 2, 2)])
def test_persistent_ctas_with_all_fusions_match_oracle(emu_dyn, n, B, depth):
    lib = emu_dyn
    b0, b1, b2, b3 = lib.qb_emu_fused_inits(), lib.qb_emu_fused_seeds(), lib.qb_emu_stream_launches(), lib.qb_emu_fused_probs()
    _sel_case(lib, n, B, depth, 20 + n, extra=[(O.OP_CZ, 0, n - 1, 0), (O.OP_CNOT, n - 1, 1, 0), (O.OP_SWAP, 2, n - 2, 0)], opts=dict(final_layout=1))
    assert lib.qb_emu_fused_inits() > b0 and lib.qb_emu_fused_seeds() > b1 and lib.qb_emu_stream_launches() > b2
    assert lib.qb_emu_fused_probs() > b3
    if n == 13:
        _case(lib, 13, 2, 160, 61, O.MEASURE_PROBS, torch.float32, opts=dict(final_layout=1))
        _case(lib, 12, 2, 140, 62, O.MEASURE_STATE, torch.float32, with_init=True)


@pytest.mark.parametrize("n,B,depth", [(12, 2, 1), (13, 2, 2), (14, 1, 3)])
def test_fused_probability_reduction_matches_oracle(emu, n, B, depth):
    """qb_forward_dev skips probs_partial_kernel; the last forward sweep squares each finished tile
    and writes probs_partial_kernel's row per CTA (S1 of the 12 tile-index bits by layout bit, the tile total for the out-of-tile bits of
    the tile's base), probs_finalize_kernel unchanged.  Permuted final layout as the engine plans it."""
    before = emu.qb_emu_fused_probs()
    _sel_case(emu, n, B, depth, 30 + n, extra=[(O.OP_CZ, 0, n - 1, 0), (O.OP_SWAP, 2, n - 2, 0), (O.OP_CNOT, n - 1, 1, 0)], opts=dict(final_layout=1))
    assert emu.qb_emu_fused_probs() > before, "the partial-sum pass was not skipped"


def test_fused_probability_reduction_mixed_programs_and_fallbacks(emu):
    lib = emu
    before = lib.qb_emu_fused_probs()
    _case(lib, 12, 2, 140, 71, O.MEASURE_PROBS, torch.float32, with_init=True, opts=dict(final_layout=1))
    _case(lib, 14, 2, 160, 72, O.MEASURE_PROBS, torch.float32, opts=dict(final_layout=1))
    _case(lib, 9, 2, 100, 73, O.MEASURE_PROBS, torch.float32, opts=dict(tile_bits=5, low_bits=2, final_layout=1))  # partial tiles, 64-thread CTAs
    _case(lib, 13, 3, 140, 74, O.MEASURE_PROBS, torch.float32, opts=dict(max_ops_per_sweep=6, final_layout=1))
    assert lib.qb_emu_fused_probs() >= before + 3
    mid = lib.qb_emu_fused_probs()
    _case(lib, 12, 2, 140, 76, O.MEASURE_STATE, torch.float32)
    _case(lib, 12, 1, 120, 77, O.MEASURE_PROBS, torch.float64)
    _case(lib, 11, 2, 120, 78, O.MEASURE_PROBS, torch.float32, opts=dict(flat=-1))
    assert lib.qb_emu_fused_probs() == mid


def test_fused_init_and_probabilities_in_a_one_sweep_plan(emu_dyn):
    """One sweep that builds |0...0>, applies the gates and reduces the probabilities (both flags on the same launch)."""
    lib = emu_dyn
    b0, b3 = lib.qb_emu_fused_inits(), lib.qb_emu_fused_probs()
    _sel_case(lib, 12, 2, 1, 97, opts=dict(final_layout=1))
    assert lib.qb_emu_fused_inits() > b0 and lib.qb_emu_fused_probs() > b3


def test_separate_measurement_passes_still_match_oracle(emu_unfused):
    """Test hook QB_FUSE=0: init_zero_kernel, probs_partial_kernel and seed_probs_kernel as separate passes (what sharded plans, complex128 and
    the non-flat kernel families always use)."""
    lib = emu_unfused
    b0, b1, b3 = lib.qb_emu_fused_inits(), lib.qb_emu_fused_seeds(), lib.qb_emu_fused_probs()
    _sel_case(lib, 13, 2, 2, 131, extra=[(O.OP_CZ, 0, 12, 0), (O.OP_SWAP, 2, 11, 0)], opts=dict(final_layout=1))
    _case(lib, 12, 2, 140, 132, O.MEASURE_PROBS, torch.float32)
    assert (lib.qb_emu_fused_inits(), lib.qb_emu_fused_seeds(), lib.qb_emu_fused_probs()) == (b0, b1, b3)


def test_persistent_launch_is_the_default_for_full_tiles(emu):
    before = emu.qb_emu_dyn_launches()
    _sel_case(emu, 13, 2, 1, 133)
    assert emu.qb_emu_dyn_launches() > before
