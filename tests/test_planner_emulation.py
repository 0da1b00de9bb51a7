"""CPU tests of the host-side planner (qandle_b200/csrc/plan.cpp) through the C ABI.

The plan is dumped (qb_plan_dump) and interpreted by tests/plan_emulator.py, which mirrors the CUDA sweep
kernels op for op; the result must equal the oracle.  This validates tiling, ext-controlled ops, 1-qubit
fusion, SWAP relabelling, exchange steps and the fused-group adjoint gradient math without a GPU.
No compute entry point of the library is called here.
"""
import random

import numpy as np
import pytest
import torch

import plan_emulator as E
from oracle import statevec as O
from qandle_b200 import engine


def make_plan(prog, n, dtype=engine.C128, tile_bits=0, low_bits=0, fuse=0, n_local=0, swap_relabel=0, final_layout=0,
              max_ops=0, flat=0, narrow_sync=0, exchange_any_bit=0, sweep_search=0):
    program = torch.tensor(prog, dtype=torch.int32).reshape(-1, 4)
    plan = engine.Plan(program, n, dtype, (tile_bits, low_bits, fuse, n_local, 1, swap_relabel, final_layout, max_ops, 0, 0, flat, narrow_sync,
                                           exchange_any_bit, sweep_search))
    return plan, engine.parse_plan_dump(plan.dump().tolist())


def random_program(rng, n, G, n_shared, n_batch, n_mats, p2=0.4):
    rows = []
    for _ in range(G):
        r = rng.random()
        if n >= 2 and r < p2:
            a, b = rng.sample(range(n), 2)
            rows.append((rng.choice([O.OP_CNOT, O.OP_CZ, O.OP_SWAP]), a, b, 0))
        elif n_mats and r < p2 + 0.08:
            rows.append((O.OP_U, rng.randrange(n), -1, rng.randrange(n_mats)))
        elif n_batch and r < p2 + 0.25:
            rows.append((rng.choice([O.OP_RX, O.OP_RY, O.OP_RZ]) | O.FLAG_BATCH, rng.randrange(n), -1, rng.randrange(n_batch)))
        else:
            rows.append((rng.choice([O.OP_RX, O.OP_RY, O.OP_RZ, O.OP_RZ]), rng.randrange(n), -1, rng.randrange(n_shared)))
    return rows


def rand_unitaries(gen, k):
    out = []
    for _ in range(k):
        a = torch.complex(torch.randn(2, 2, generator=gen, dtype=torch.float64), torch.randn(2, 2, generator=gen, dtype=torch.float64))
        q, r = torch.linalg.qr(a)
        out.append(q * (torch.diagonal(r) / torch.diagonal(r).abs()))
    return torch.stack(out)


def rand_state(gen, B, n):
    st = torch.complex(torch.randn(B, 2**n, generator=gen, dtype=torch.float64), torch.randn(B, 2**n, generator=gen, dtype=torch.float64))
    return st / torch.linalg.norm(st, dim=-1, keepdim=True)


CONFIGS = [
    # n, G, B, tile_bits, low_bits, fuse(0 on / -1 off), swap_relabel(0 on / -1 off)
    (1, 10, 2, 0, 0, 0, 0),
    (2, 20, 2, 0, 0, 0, 0),
    (3, 30, 3, 2, 1, 0, 0),
    (5, 60, 2, 3, 1, 0, 0),
    (5, 60, 2, 3, 1, -1, -1),
    (6, 80, 2, 4, 2, 0, -1),
    (7, 120, 1, 3, 1, 0, 0),
    (8, 150, 2, 5, 2, 0, 0),
    (9, 100, 1, 4, 1, -1, 0),
    (10, 200, 1, 8, 2, 0, 0),
    (11, 250, 1, 7, 3, 0, -1),
]


@pytest.mark.parametrize("cfg", CONFIGS)
@pytest.mark.parametrize("seed", [0, 1])
@pytest.mark.parametrize("dtype", [engine.C128, engine.C64], ids=["c128", "c64flat"])
def test_plan_forward_and_adjoint_match_oracle(cfg, seed, dtype):
    """dtype only selects the planner path here (complex64 plans use flat stages with a separate backward
    linearisation); the emulator always computes in float64."""
    n, G, B, tb, lb, fuse, relabel = cfg
    rng = random.Random(100 * seed + n)
    gen = torch.Generator().manual_seed(7 + seed)
    n_shared, n_batch, n_mats = 6, 2, 2
    prog = random_program(rng, n, G, n_shared, n_batch, n_mats)
    shared = ((torch.rand(n_shared, generator=gen, dtype=torch.float64) - 0.5) * 6).requires_grad_(True)
    batch = ((torch.rand(B, n_batch, generator=gen, dtype=torch.float64) - 0.5) * 6).requires_grad_(True)
    mats = rand_unitaries(gen, n_mats)
    init = rand_state(gen, B, n).requires_grad_(True)

    _plan, pd = make_plan(prog, n, dtype=dtype, tile_bits=tb, low_bits=lb, fuse=fuse, swap_relabel=relabel)
    assert pd["final_pos"] == [n - 1 - q for q in range(n)]  # final_layout=0 restores the identity layout

    ref = O.run_program(prog, n, shared, batch, mats, init, B, O.MEASURE_STATE)
    got = E.emulate_forward(pd, init.detach().numpy(), shared.detach().numpy(), batch.detach().numpy(), mats.numpy())
    assert np.allclose(got, ref.detach().numpy(), atol=1e-12)

    # adjoint backward with a random complex cotangent on the state
    g = torch.complex(torch.randn(B, 2**n, generator=gen, dtype=torch.float64), torch.randn(B, 2**n, generator=gen, dtype=torch.float64))
    ref.backward(g)
    lam = 0.5 * g.numpy()  # seed_state_kernel
    gs, gb, lam0, psi0 = E.emulate_backward(pd, got, lam, shared.detach().numpy(), batch.detach().numpy(), mats.numpy(),
                                            n_shared, n_batch)
    assert np.allclose(psi0, init.detach().numpy(), atol=1e-12)  # un-computation returns the initial state
    assert np.allclose(gs, shared.grad.numpy(), atol=1e-10)
    assert np.allclose(gb, batch.grad.numpy(), atol=1e-10)
    assert np.allclose(2 * lam0, init.grad.numpy(), atol=1e-10)


@pytest.mark.parametrize("n,tb", [(4, 2), (6, 3), (8, 4)])
def test_plan_probs_with_permuted_final_layout(n, tb):
    rng = random.Random(n)
    gen = torch.Generator().manual_seed(n)
    prog = random_program(rng, n, 60, 5, 0, 0, p2=0.5)
    shared = ((torch.rand(5, generator=gen, dtype=torch.float64) - 0.5) * 6).requires_grad_(True)
    B = 2
    _plan, pd = make_plan(prog, n, tile_bits=tb, low_bits=1, final_layout=1)
    ref = O.run_program(prog, n, shared, None, None, None, B, O.MEASURE_PROBS)
    st0 = O.zero_state(n, B, torch.float64).numpy()
    full = E.emulate_forward(pd, st0, shared.detach().numpy(), None, None)
    assert np.allclose(E.probs_from_physical(pd, full), ref.detach().numpy(), atol=1e-12)
    g = torch.randn(B, n, generator=gen, dtype=torch.float64)
    ref.backward(g)
    lam = E.seed_probs(pd, full, g.numpy())
    gs, _gb, _l0, psi0 = E.emulate_backward(pd, full, lam, shared.detach().numpy(), None, None, 5, 0)
    assert np.allclose(gs, shared.grad.numpy(), atol=1e-10)
    assert np.allclose(psi0, st0, atol=1e-12)


@pytest.mark.parametrize("any_bit", [0, 1])
@pytest.mark.parametrize("n,world,tb", [(6, 2, 3), (8, 4, 3), (9, 8, 4), (13, 2, 5), (14, 4, 6)])
def test_plan_amplitude_sharded_matches_oracle(n, world, tb, any_bit):
    """north_star (d): top log2(world) qubits are rank bits; exchange steps swap them with local bits -- the top ones, or (plan option
    exchange_any_bit) the ones the planner picks per exchange."""
    g_bits = world.bit_length() - 1
    n_local = n - g_bits
    rng = random.Random(n + world)
    gen = torch.Generator().manual_seed(n * world)
    prog = random_program(rng, n, 90, 6, 0, 0, p2=0.4)
    shared = ((torch.rand(6, generator=gen, dtype=torch.float64) - 0.5) * 6).requires_grad_(True)
    B = 1
    _plan, pd = make_plan(prog, n, tile_bits=tb, low_bits=1, n_local=n_local, final_layout=1, exchange_any_bit=any_bit)
    assert pd["n_local"] == n_local
    n_ex = sum(1 for s in pd["steps"] if s["type"] == E.STEP_EXCHANGE)
    assert n_ex >= 1 and n_ex == len(pd["exchanges"])
    top = list(range(n_local - g_bits, n_local))
    assert all(sorted(set(ex)) == ex and len(ex) == g_bits and all(0 <= b < n_local for b in ex) for ex in pd["exchanges"])
    if not any_bit:
        assert all(ex == top for ex in pd["exchanges"])
    assert [_plan.exchange_bits(i) for i, s in enumerate(pd["steps"]) if s["type"] == E.STEP_EXCHANGE] == pd["exchanges"]
    ref = O.run_program(prog, n, shared, None, None, None, B, O.MEASURE_PROBS)
    st0 = O.zero_state(n, B, torch.float64).numpy()
    full = E.emulate_forward(pd, st0, shared.detach().numpy(), None, None, world=world)
    assert np.allclose(E.probs_from_physical(pd, full), ref.detach().numpy(), atol=1e-12)
    g = torch.randn(B, n, generator=gen, dtype=torch.float64)
    ref.backward(g)
    lam = E.seed_probs(pd, full, g.numpy())
    gs, _gb, _l0, psi0 = E.emulate_backward(pd, full, lam, shared.detach().numpy(), None, None, 6, 0, world=world)
    assert np.allclose(gs, shared.grad.numpy(), atol=1e-10)


def _overhead_model(pd):
    """plan.cpp: sweep_overhead_ms (the measured per-sweep cost model the sweep-size search minimises)"""
    fix = lambda sts: sum(1 for st in sts if st["d_end"] > st["pre_end"])
    return sum(1.22 + 0.146 * len(sw["stages"]) + 0.087 * fix(sw["stages"]) + 0.283 * len(sw["stages_bwd"]) + 0.245 * fix(sw["stages_bwd"])
               for sw in pd["sweeps"])


@pytest.mark.parametrize("n,depth", [(13, 5), (15, 4)])
def test_sweep_size_search_keeps_the_circuit_and_never_costs_more(n, depth):
    """plan option sweep_search (default on): where each sweep ends is searched a few sweeps ahead, and 128-byte chunks (low_bits 4) are a
    candidate; the kept plan has the smallest modelled overhead of all candidates -- never more than the greedy plan's -- and is the same
    operator (forward and adjoint against the oracle)."""
    rows = [(O.OP_RX | O.FLAG_BATCH, k, -1, k) for k in range(n)] + O.sel_program(list(range(n)), depth)
    _pg, greedy = make_plan(rows, n, dtype=engine.C64, final_layout=1, sweep_search=-1)
    _ps, searched = make_plan(rows, n, dtype=engine.C64, final_layout=1)
    assert _overhead_model(searched) <= _overhead_model(greedy) + 1e-9
    assert len(searched["sweeps"]) <= len(greedy["sweeps"])
    n_ops = lambda d: sorted((o["kind"], o["mat"]) for sw in d["sweeps"] for o in sw["ops"] if o["kind"] in (1, 2, 3))
    assert n_ops(searched) == n_ops(greedy)  # the same fused 2x2 groups, scheduled differently
    gen = torch.Generator().manual_seed(n)
    B = 2
    n_shared = 3 * n * depth
    shared = ((torch.rand(n_shared, generator=gen, dtype=torch.float64) - 0.5) * 6).requires_grad_(True)
    batch = ((torch.rand(B, n, generator=gen, dtype=torch.float64) - 0.5) * 6).requires_grad_(True)
    ref = O.run_program(rows, n, shared, batch, None, None, B, O.MEASURE_PROBS)
    st0 = O.zero_state(n, B, torch.float64).numpy()
    g = torch.randn(B, n, generator=gen, dtype=torch.float64)
    ref.backward(g)
    for pd in (greedy, searched):
        full = E.emulate_forward(pd, st0, shared.detach().numpy(), batch.detach().numpy(), None)
        assert np.allclose(E.probs_from_physical(pd, full), ref.detach().numpy(), atol=1e-12)
        lam = E.seed_probs(pd, full, g.numpy())
        gs, gb, _l0, psi0 = E.emulate_backward(pd, full, lam, shared.detach().numpy(), batch.detach().numpy(), None, n_shared, n)
        assert np.allclose(gs, shared.grad.numpy(), atol=1e-10) and np.allclose(gb, batch.grad.numpy(), atol=1e-10)
        assert np.allclose(psi0, st0, atol=1e-12)


def test_drained_exchange_plans_need_few_exchanges():
    """With planner-chosen exchange bits the plan drains the runnable gates before it exchanges: a brickwork circuit of
    16 qubits on 4 ranks needs fewer exchanges than with top-bit exchanges, and no more sweeps."""
    import random as _r

    n, world = 16, 4
    rng = _r.Random(0)
    rows, slot = [], 0
    for _ in range(12):
        for k in range(n):
            rows.append((rng.choice([O.OP_RX, O.OP_RY, O.OP_RZ]), k, -1, slot))
            slot += 1
        for k in list(range(0, n - 1, 2)) + list(range(1, n - 1, 2)):
            rows.append((O.OP_CZ, k, k + 1, 0))
    count = lambda pd: (len(pd["sweeps"]), len(pd["exchanges"]))
    _p0, top = make_plan(rows, n, tile_bits=6, low_bits=2, n_local=n - 2, final_layout=1, exchange_any_bit=0)
    _p1, anyb = make_plan(rows, n, tile_bits=6, low_bits=2, n_local=n - 2, final_layout=1, exchange_any_bit=1)
    assert count(anyb)[1] < count(top)[1] and count(anyb)[0] <= count(top)[0], (count(anyb), count(top))


def test_fusion_reduces_groups_and_sel_sweeps():
    rows = O.sel_program(list(range(16)), depth=10)
    plan, pd = make_plan(rows, 16, dtype=engine.C64)
    # RZ RY RZ on each qubit fuse into one 2x2: 16 groups per layer
    assert len(pd["groups"]) == 160
    assert all(g["member_count"] == 3 for g in pd["groups"])
    assert plan.num_sweeps <= 40
    plan_nf, pd_nf = make_plan(rows, 16, dtype=engine.C64, fuse=-1)
    assert len(pd_nf["groups"]) == 480


def test_bad_programs_fail_loudly():
    with pytest.raises(RuntimeError):
        make_plan([(O.OP_RX, 5, -1, 0)], 3)
    with pytest.raises(RuntimeError):
        make_plan([(O.OP_CNOT, 1, 1, 0)], 3)
    with pytest.raises(RuntimeError):
        make_plan([(99, 0, -1, 0)], 3)


@pytest.mark.parametrize("dtype,rb", [(engine.C64, 4), (engine.C128, 3)])
def test_stage_metadata_is_consistent(dtype, rb):
    """Register-blocked stages (plan.h: Stage): every non-diagonal target is a register bit of its stage and the
    r / rc fields index the stage's register bits."""
    rng = random.Random(3)
    n = 10
    prog = random_program(rng, n, 300, 6, 2, 2, p2=0.45)
    _plan, pd = make_plan(prog, n, dtype=dtype, tile_bits=8, low_bits=2, swap_relabel=0, flat=-1)
    n_staged = 0
    for sw in pd["sweeps"]:
        if not sw["stages"]:
            assert len(sw["tile_bits"]) < rb or any(op["kind"] == E.K_SWAP for op in sw["ops"])
            continue
        n_staged += 1
        covered = 0
        for st in sw["stages"]:
            assert len(st["regbits"]) == rb and sorted(st["regbits"]) == st["regbits"]
            if st["low"]:
                assert st["regbits"] == list(range(rb))
            assert st["op_begin"] == covered
            covered = st["op_end"]
            for op in sw["ops"][st["op_begin"]:st["op_end"]]:
                if op["kind"] in (E.K_U1, E.K_CX, E.K_CX_EXT):
                    assert op["r"] >= 0 and st["regbits"][op["r"]] == op["a"]
                if op["a"] >= 0:
                    assert (op["r"] >= 0) == (op["a"] in st["regbits"])
                    if op["r"] >= 0:
                        assert st["regbits"][op["r"]] == op["a"]
                if op["c"] >= 0:
                    assert (op["rc"] >= 0) == (op["c"] in st["regbits"])
                    if op["rc"] >= 0:
                        assert st["regbits"][op["rc"]] == op["c"]
        assert covered == len(sw["ops"])
    assert n_staged > 0


KIND_SIGN = (E.K_CZ, E.K_CZ_EXT1, E.K_CZ_EXT2)


def _check_flat_stages(ops, stages, packed=True):
    covered = 0
    for st in stages:
        assert st["flat"] == 1
        if packed:
            assert len(st["regbits"]) == 4 and st["regbits"][0] == 0
        else:
            assert len(st["regbits"]) == 3
        assert st["op_begin"] == covered
        assert st["op_begin"] <= st["pre_end"] <= st["la_end"] <= st["d_end"] <= st["suf_begin"] <= st["op_end"]
        covered = st["op_end"]
        for i in range(st["op_begin"], st["op_end"]):
            op = ops[i]
            for fld, rf in (("a", "r"), ("c", "rc")):
                if op[fld] >= 0:
                    assert (op[rf] >= 0) == (op[fld] in st["regbits"])
                    if op[rf] >= 0:
                        assert st["regbits"][op[rf]] == op[fld]
            lane_cx = packed and op["kind"] in (E.K_CX, E.K_CX_EXT) and (op["a"] == 0 or (op["kind"] == E.K_CX and op["c"] == 0))
            if i < st["pre_end"] or i >= st["suf_begin"]:
                assert op["kind"] in (E.K_CX, E.K_CX_EXT) and not lane_cx  # absorbed into the addressing
            elif i < st["la_end"]:
                assert lane_cx
                if op["kind"] == E.K_CX and op["c"] == 0:
                    assert op["r"] >= 1  # lanes are exchanged between two of the thread's packs
            elif i < st["d_end"]:
                assert op["kind"] in KIND_SIGN or op["kind"] == E.K_D1_EXT or (op["kind"] == E.K_D1 and op["r"] < 0)
            else:
                assert op["kind"] in (E.K_U1, E.K_D1) and op["r"] >= 0 and st["u_op"][op["r"]] == i
        n_u = sum(1 for u in st["u_op"] if u >= 0)
        assert n_u == st["suf_begin"] - st["d_end"] and st["shape"] == sum(1 << r for r in range(4) if st["u_op"][r] >= 0)
        dops = ops[st["la_end"]:st["d_end"]]
        assert st["n_sign"] == sum(1 for o in dops if o["kind"] in KIND_SIGN) and st["n_sign"] + st["n_phase"] == len(dops)
    assert covered == len(ops)


@pytest.mark.parametrize("dtype", [engine.C64, engine.C128], ids=["c64", "c128"])
def test_flat_stage_metadata_is_consistent(dtype):
    """flat stages (plan.h: Stage::flat) for the forward order and for the backward linearisation: complex64 (pack lane +
    3 register bits) and complex128 (3 register bits, every CNOT absorbed)."""
    rng = random.Random(5)
    n = 10
    for trial in range(4):
        prog = random_program(rng, n, 300, 6, 2, 2, p2=0.45)
        _plan, pd = make_plan(prog, n, dtype=dtype, tile_bits=8 - trial, low_bits=2, swap_relabel=0)
        n_flat = 0
        for sw in pd["sweeps"]:
            if not sw["stages"]:
                assert not sw["ops_bwd"]
                continue
            n_flat += 1
            assert len(sw["ops_bwd"]) == len(sw["ops"])
            key = lambda o: (o["kind"], o["a"], o["c"], o["mat"], o["ext_mask"], o["kslot"])
            assert sorted(map(key, sw["ops"])) == sorted(map(key, sw["ops_bwd"]))
            _check_flat_stages(sw["ops"], sw["stages"], packed=dtype == engine.C64)
            _check_flat_stages(sw["ops_bwd"], sw["stages_bwd"], packed=dtype == engine.C64)
        assert n_flat > 0


# ---- narrow barriers (plan.cpp: sync_cost / schedule_flat_stages; flat64.cuh: group_barrier) ------------------------
def _flat_sweeps(pd):
    return [sw for sw in pd["sweeps"] if sw["stages"] and sw["stages"][0]["flat"]]


@pytest.mark.parametrize("n,dtype", [(13, engine.C64), (14, engine.C64), (12, engine.C128), (13, engine.C128)],
                         ids=["13q-c64", "14q-c64", "12q-c128", "13q-c128"])
@pytest.mark.parametrize("seed", [0, 1])
def test_narrow_barrier_plans_match_oracle_and_cover_every_handover(n, dtype, seed):
    """Full 256-thread tiles: the planner relabels the staged bits and replaces CTA barriers between flat stages by
    warp / sub-CTA barriers.  (i) the relabelled, re-scheduled plan still equals the oracle (forward and adjoint);
    (ii) a thread-level model of the kernels' shared-memory addressing shows that every hand-over of amplitudes between
    threads stays inside the thread group of the barrier the planner declared."""
    rng = random.Random(1000 * seed + n)
    gen = torch.Generator().manual_seed(11 + seed)
    n_shared, n_batch, n_mats, B = 8, 2, 2, 1
    prog = random_program(rng, n, 260, n_shared, n_batch, n_mats, p2=0.35)
    prog = [r for r in prog if (r[0] & 0xFF) != O.OP_SWAP]  # SWAPs relabel; keep the sweeps on the flat kernels
    shared = ((torch.rand(n_shared, generator=gen, dtype=torch.float64) - 0.5) * 6).requires_grad_(True)
    batch = ((torch.rand(B, n_batch, generator=gen, dtype=torch.float64) - 0.5) * 6).requires_grad_(True)
    mats = rand_unitaries(gen, n_mats)
    init = rand_state(gen, B, n).requires_grad_(True)
    _plan, pd = make_plan(prog, n, dtype=dtype)
    packed = dtype == engine.C64
    counts = [0, 0, 0, 0]
    flat = _flat_sweeps(pd)
    assert flat, "expected flat sweeps"
    for sw in flat:
        L = 4  # lowest bits that are always staged: 5 (256-byte HBM chunks) or 4 (the planner's 128-byte-chunk candidate) for complex64
        assert sw["tile_bits"][:L] == list(range(L)) and sorted(sw["tile_bits"]) == sorted(set(sw["tile_bits"]))
        for bwd in (False, True):
            counts = [a + b for a, b in zip(counts, E.check_flat_sync(sw, bwd, packed))]
    assert counts[1] + counts[2] + counts[3] > 0, "no barrier was narrowed"
    ref = O.run_program(prog, n, shared, batch, mats, init, B, O.MEASURE_STATE)
    got = E.emulate_forward(pd, init.detach().numpy(), shared.detach().numpy(), batch.detach().numpy(), mats.numpy())
    assert np.allclose(got, ref.detach().numpy(), atol=1e-12)
    g = torch.complex(torch.randn(B, 2**n, generator=gen, dtype=torch.float64), torch.randn(B, 2**n, generator=gen, dtype=torch.float64))
    ref.backward(g)
    gs, gb, lam0, psi0 = E.emulate_backward(pd, got, 0.5 * g.numpy(), shared.detach().numpy(), batch.detach().numpy(),
                                            mats.numpy(), n_shared, n_batch)
    assert np.allclose(psi0, init.detach().numpy(), atol=1e-12)
    assert np.allclose(gs, shared.grad.numpy(), atol=1e-10)
    assert np.allclose(gb, batch.grad.numpy(), atol=1e-10)
    assert np.allclose(2 * lam0, init.grad.numpy(), atol=1e-10)


def test_narrow_barriers_on_the_sel_workload_shape():
    """BASELINE config 2's circuit shape (16-qubit strongly-entangling ansatz): the narrowed plan keeps the stage count
    of the CTA-barrier plan and narrows a good part of the inner barriers."""
    n, depth = 16, 4
    rows = [(O.OP_RX | O.FLAG_BATCH, k, -1, k) for k in range(n)] + O.sel_program(list(range(n)), depth)
    _p, pd = make_plan(rows, n, dtype=engine.C64, final_layout=1)
    _p0, pd0 = make_plan(rows, n, dtype=engine.C64, final_layout=1, narrow_sync=-1)  # CTA barriers only (qb_plan_opts.narrow_sync)
    n_st = lambda d, key: sum(len(sw[key]) for sw in d["sweeps"])
    assert n_st(pd, "stages") <= n_st(pd0, "stages") and n_st(pd, "stages_bwd") <= n_st(pd0, "stages_bwd")
    assert all(st["narrow_end"] == 0 and st["narrow_x"] == 0 for sw in pd0["sweeps"] for st in sw["stages"] + sw["stages_bwd"])
    counts = [0, 0, 0, 0]
    for sw in _flat_sweeps(pd):
        for bwd in (False, True):
            counts = [a + b for a, b in zip(counts, E.check_flat_sync(sw, bwd, True))]
    assert counts[1] + counts[2] + counts[3] >= counts[0] // 2, counts
