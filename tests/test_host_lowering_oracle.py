"""CPU test of the whole HOST path of the drop-in API against the reference's golden vectors.

The product has no CPU backend, so these tests inject one as TEST INFRASTRUCTURE: `engine.run_circuit` is replaced by
the oracle's interpreter of the gate-program IR (oracle/statevec.py) and plans by the lowered segment itself.  Everything
above the custom-op boundary is the product's own code: lowering of the layer list (incl. Invert / Controlled
decompositions), weight gathering through the remapping, named / embedding input columns, state handling, the
reference's output-shape rules, torch modules between segments (Reset).  Outputs, shapes, dtypes and every gradient must
equal what the real reference produced for the same constructor calls (tests/golden/api_cases.npz).
"""
import os

import numpy as np
import pytest
import torch

import specs
import qandle_b200 as q
from oracle import statevec as O
from qandle_b200 import engine, qcircuit

GOLD = os.path.join(os.path.dirname(__file__), "golden", "api_cases.npz")


@pytest.fixture()
def oracle_backend(monkeypatch):
    def run_circuit(plan, shared, batch, mats, init, B, measure):
        seg, n = plan
        fm = torch.view_as_complex(mats.reshape(-1, 2, 2, 2)) if mats.numel() else None
        out = O.run_program(seg.rows, n, shared, batch if batch.numel() else None, fm, init, B, measure)
        return out

    monkeypatch.setattr(engine, "require_cuda", lambda: torch.device("cpu"))
    monkeypatch.setattr(qcircuit, "_plan_for", lambda seg, n, real_dtype, dev=None: (seg, n))
    monkeypatch.setattr(engine, "run_circuit", run_circuit)


def rel_err(a, b):
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)


@pytest.mark.parametrize("name", sorted(specs.api_specs().keys()))
def test_host_path_on_oracle_backend_matches_reference(oracle_backend, name):
    z = np.load(GOLD)
    g = lambda k: z[f"{name}/{k}"]
    case = specs.api_specs()[name]
    circ = specs.build_circuit(q, case["spec"], case["num_qubits"])
    sd = {k[len(name) + 3:]: torch.tensor(z[k]) for k in z.files if k.startswith(f"{name}/p.")}
    circ.load_state_dict(sd)
    inputs = {k: torch.tensor(g(f"in.{k}")).requires_grad_(True) for k in case["inputs"]}
    state = torch.tensor(g("state")).requires_grad_(True) if bool(g("has_state")) else None
    out = circ(state, **inputs)
    ref = torch.tensor(g("out"))
    assert tuple(out.shape) == tuple(ref.shape) and out.dtype == ref.dtype
    assert rel_err(out.detach(), ref) < 1e-5
    out.backward(torch.tensor(g("g")))
    for k in case["inputs"]:
        refg = torch.tensor(g(f"gin.{k}"))
        assert float((inputs[k].grad - refg).abs().max()) < 2e-5 * max(1.0, float(refg.abs().max())), k
    if state is not None:
        refg = torch.tensor(g("grad_state"))
        assert float((state.grad - refg).abs().max()) < 2e-5 * max(1.0, float(refg.abs().max()))
    for k, p in circ.named_parameters():
        refg = torch.tensor(g(f"gp.{k}"))
        if torch.isnan(refg).any():
            assert p.grad is None or float(p.grad.abs().max()) == 0.0
        else:
            assert p.grad is not None, k
            assert float((p.grad - refg).abs().max()) < 2e-5 * max(1.0, float(refg.abs().max())), k


def test_controlled_rejects_what_the_engine_cannot_express():
    with pytest.raises(NotImplementedError):
        q.Controlled(1, q.RX(1)).build(num_qubits=2)  # target on the control qubit
    with pytest.raises(NotImplementedError):
        q.Controlled(0, q.U(1, torch.tensor([[1.0, 1.0], [0.0, 1.0]]))).build(num_qubits=2)  # not unitary
    with pytest.raises(NotImplementedError):
        q.Controlled(0, q.Controlled(1, q.RX(2))).build(num_qubits=3)


def test_controlled_dense_matrix_matches_lowering(oracle_backend):
    """to_matrix (row-vector convention) of a Controlled gate equals its engine lowering on random states."""
    torch.manual_seed(3)
    u = torch.linalg.qr(torch.complex(torch.randn(2, 2), torch.randn(2, 2)))[0]
    for tgt in (q.RY(0, theta=0.9, remapping=None), q.RX(2, theta=-0.4), q.U(0, u), q.CNOT(2, 0), q.CZ(0, 2), q.SWAP(0, 2)):
        gate = q.Controlled(1, tgt).build(num_qubits=3)
        st = torch.complex(torch.randn(4, 8), torch.randn(4, 8))
        got = qcircuit.run_modules(gate, [gate], 3, st, {})
        assert torch.allclose(got, st @ gate.to_matrix(), atol=1e-5), type(tgt).__name__


@pytest.mark.parametrize("pauli", ["z", "x", "y"])
def test_expectation_values_match_direct_evaluation(oracle_backend, pauli):
    """<P_q> from the probability reduction after a basis change == <psi| P_q |psi> evaluated densely."""
    torch.manual_seed(5)
    n, B = 4, 3
    layers = [q.AngleEmbedding(name="x", qubits=list(range(n)), rotation="ry"),
              q.StronglyEntanglingLayer(qubits=list(range(n)), depth=2, remapping=None)]
    circ = q.Circuit(layers=layers + [q.MeasureExpectation(pauli)], num_qubits=n)
    ref_circ = q.Circuit(layers=circ.circuit.layers[:-1] + [q.MeasureState()], num_qubits=n)
    x = torch.rand(B, n, requires_grad=True)
    out = circ(x=x)
    assert tuple(out.shape) == (B, n)
    psi = ref_circ(x=x.detach())
    P = {"x": torch.tensor([[0, 1], [1, 0]]), "y": torch.tensor([[0, -1j], [1j, 0]]), "z": torch.tensor([[1, 0], [0, -1]])}[pauli].to(torch.complex64)
    for k in range(n):
        full = O.apply_1q(psi, P, k, n)
        ev = (psi.conj() * full).sum(dim=1).real
        assert torch.allclose(out[:, k].detach(), ev, atol=1e-5)
    out.sum().backward()
    assert x.grad is not None and torch.isfinite(x.grad).all()
    with pytest.raises(ValueError):
        q.MeasureExpectation("w")
