"""CPU tests of the host side: operator surface (constructors, Parameter names / shapes vs the reference's golden
state_dicts), lowering to the gate-program IR, C-ABI symbol export, and loud failure without a CUDA device."""
import os
import re
import subprocess

import numpy as np
import pytest
import torch

import specs
import qandle_b200 as q
from qandle_b200 import engine, qcircuit

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "api_cases.npz")


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "qandle_b200.h")).read()
    declared = set(re.findall(r"\b(qb_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 25
    out = subprocess.run(["nm", "-D", "--defined-only", engine.LIB_CORE], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (qb_[a-z_0-9]+)", out))
    assert declared <= exported, sorted(declared - exported)
    lib = engine.core()
    assert b"sm_100a" in lib.qb_version()


def test_torch_library_ops_are_registered():
    ops = engine.load_ops()
    for name in ("plan_create", "circuit_forward", "circuit_backward", "apply_forward", "apply_backward", "measure_probs",
                 "seed_probs", "finalize_grads", "exchange_p2p"):
        assert hasattr(ops, name)


@pytest.mark.parametrize("name", sorted(specs.api_specs().keys()))
def test_state_dict_matches_reference(name):
    """Same Parameter names and shapes as the reference: reference checkpoints load unchanged."""
    z = np.load(GOLD)
    case = specs.api_specs()[name]
    circ = specs.build_circuit(q, case["spec"], case["num_qubits"])
    assert circ.num_qubits == int(z[f"{name}/n"])
    ref = {k[len(name) + 3:]: z[k] for k in z.files if k.startswith(f"{name}/p.")}
    sd = circ.state_dict()
    assert set(sd.keys()) == set(ref.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(ref[k].shape) and v.dtype == torch.float32
    circ.load_state_dict({k: torch.tensor(v) for k, v in ref.items()})


def test_lowering_of_hybrid_circuit():
    c = q.Circuit(layers=[q.AngleEmbedding(name="x", qubits=[0, 1, 2]), q.RX(0), q.RY(1, name="phi"), q.CNOT(0, 2), q.CZ(1, 2),
                          q.SWAP(0, 1), q.U(2, torch.eye(2)), q.StronglyEntanglingLayer(qubits=[0, 1, 2], depth=2),
                          q.MeasureProbability()], num_qubits=3)
    segs = qcircuit.lower_modules(c.circuit.layers, 3)
    assert len(segs) == 1
    s = segs[0]
    assert s.init == "zero" and s.measure == engine.MEASURE_PROBS
    kinds = [r[0] & 0xFF for r in s.rows]
    assert kinds[:3] == [engine.OP_RX] * 3 and all(r[0] & engine.FLAG_BATCH for r in s.rows[:3])
    assert s.batch_cols == [("x", 0), ("x", 1), ("x", 2), ("phi", -1)]
    assert kinds[3:9] == [engine.OP_RX, engine.OP_RY, engine.OP_CNOT, engine.OP_CZ, engine.OP_SWAP, engine.OP_U]
    assert s.rows[4][0] & engine.FLAG_BATCH and s.rows[4][3] == 3  # named RY -> batch column 3
    assert len(s.weight_mods) == 1 + 2 * 3 * 3  # RX + SEL(2 x 3 qubits x 3 rotations); the named RY owns an unused theta
    assert len(s.mats) == 1
    sel_rows = s.rows[9:]
    assert [r[0] for r in sel_rows[:9]] == [engine.OP_RZ, engine.OP_RY, engine.OP_RZ] * 3
    assert [tuple(r[1:3]) for r in sel_rows[9:12]] == [(0, 1), (1, 2), (2, 0)]  # CNOT ring, range 1
    assert [tuple(r[1:3]) for r in sel_rows[21:24]] == [(0, 2), (1, 0), (2, 1)]  # second layer: range 2


def test_embedding_resets_state_and_foreign_modules_split_segments():
    class Scale(torch.nn.Module):
        named = False

        def forward(self, state):
            return state * 1.0

    c = q.Circuit(layers=[q.RX(0), q.AmplitudeEmbedding(name="a", qubits=[0, 1], normalize=True), q.RY(1), Scale(), q.RZ(0),
                          q.MeasureState()], num_qubits=2)
    segs = qcircuit.lower_modules(c.circuit.layers, 2)
    assert len(segs) == 2
    assert segs[0].init[0] == "amp" and len(segs[0].rows) == 1 and isinstance(segs[0].foreign, Scale)  # RX(0) before the embedding is dead
    assert segs[1].init == "inherit" and len(segs[1].rows) == 1


def test_constructor_errors_match_reference():
    with pytest.raises(AssertionError):
        q.RX("a")
    with pytest.raises(AssertionError):
        q.RX(-1)
    with pytest.raises(AssertionError):
        q.CNOT(1, 1)
    with pytest.raises(AssertionError):
        q.CZ(0, 0)
    with pytest.raises(ValueError):
        q.Circuit(layers=[q.MeasureProbability()])
    with pytest.raises(ValueError):
        q.parse_rot("rw")
    with pytest.raises(AssertionError):
        q.AmplitudeEmbedding(name="a", qubits=[0]).build(num_qubits=2)
    with pytest.raises(q.UnbuiltGateError):
        q.RX(0).to_qasm()


def test_parameter_shapes_follow_reference_quirks():
    g = q.RX(0).build(num_qubits=2)
    assert tuple(g.theta.shape) == (1,)  # random init -> shape (1,)   (reference operators.py:226-227, quirk Q13)
    g = q.RX(0, theta=0.5).build(num_qubits=2)
    assert tuple(g.theta.shape) == () and g.theta.dtype == torch.float32
    g = q.RX(0, name="x").build(num_qubits=2)
    assert g.named and isinstance(g.theta, torch.nn.Parameter)  # named gates keep an unused theta (Q6)
    sel = q.StronglyEntanglingLayer(qubits=[0, 1, 2], depth=2).build(num_qubits=3)
    assert len(list(sel.parameters())) == 18 and all(tuple(p.shape) == () for p in sel.parameters())


def test_to_matrix_is_row_vector_convention():
    """forward(state) == state @ to_matrix() (reference operators.py:67-69); dense helper, small n only."""
    from oracle import statevec as O

    g = q.RY(1, theta=0.7, remapping=None).build(num_qubits=3)
    st = torch.randn(2, 8, dtype=torch.complex64)
    ref = O.apply_1q(st, O.rot_matrix(O.OP_RY, torch.tensor(0.7)), 1, 3)
    assert torch.allclose(st @ g.to_matrix(), ref, atol=1e-6)
    cn = q.CNOT(0, 2).build(num_qubits=3)
    assert torch.allclose(st @ cn.to_matrix(), O.apply_cnot(st, 0, 2, 3), atol=1e-6)
    u = torch.tensor([[0.6, -0.8], [0.8, 0.6]], dtype=torch.complex64)
    bu = q.U(1, u).build(num_qubits=3)
    assert torch.allclose(st @ bu.to_matrix(), O.apply_1q(st, u.T, 1, 3), atol=1e-6)  # quirk Q2: acts as U^T


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the CPU-only failure mode")
def test_no_cpu_fallback():
    c = q.Circuit(layers=[q.RX(0), q.MeasureProbability()], num_qubits=2)
    with pytest.raises(engine.EngineUnavailable):
        c()
    with pytest.raises(RuntimeError):
        engine.Plan(torch.tensor([[1, 0, -1, 0]], dtype=torch.int32), 2, engine.C64)  # needs a device unless host_only
