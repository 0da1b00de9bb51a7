"""Reference-generated golden vectors at 12 and 13 qubits (tests/golden/full_tile_cases.npz, made by
tests/golden/generate_full_tile.py from the REAL reference): one sample = one / two full 2^12-amplitude tiles, so the
FULL-tile forward and streaming adjoint sweep kernels -- the ones bench.py times -- are compared with the reference's own
numbers (outputs, input / initial-state gradients, every parameter gradient; 1e-5 relative, complex64).

Backends as in test_reference_suite.py: `oracle` (CPU: pins the oracle + host path at these sizes) and `engine` (`-m gpu`).
"""
import os

import numpy as np
import pytest
import torch

import specs
import qandle_b200 as q
from oracle import statevec as O
from qandle_b200 import engine, qcircuit

GOLD = os.path.join(os.path.dirname(__file__), "golden", "full_tile_cases.npz")
TOL = 1e-5


@pytest.fixture(params=["oracle", pytest.param("engine", marks=pytest.mark.gpu)])
def dev(request, monkeypatch):
    if request.param == "engine":
        assert torch.cuda.is_available(), "the engine backend needs cuda:0"
        return torch.device("cuda:0")

    def run_circuit(plan, shared, batch, mats, init, B, measure):
        seg, n = plan
        fm = torch.view_as_complex(mats.reshape(-1, 2, 2, 2)) if mats.numel() else None
        return O.run_program(seg.rows, n, shared, batch if batch.numel() else None, fm, init, B, measure)

    monkeypatch.setattr(engine, "require_cuda", lambda: torch.device("cpu"))
    monkeypatch.setattr(qcircuit, "_plan_for", lambda seg, n, real_dtype, dev=None: (seg, n))
    monkeypatch.setattr(engine, "run_circuit", run_circuit)
    return torch.device("cpu")


@pytest.mark.parametrize("name", sorted(specs.full_tile_specs().keys()))
def test_full_tile_case_matches_reference(dev, name):
    z = np.load(GOLD)
    g = lambda k: torch.tensor(z[f"{name}/{k}"])  # noqa: E731
    case = specs.full_tile_specs()[name]
    circ = specs.build_circuit(q, case["spec"], case["num_qubits"])
    circ.load_state_dict({k[len(name) + 3:]: torch.tensor(z[k]) for k in z.files if k.startswith(f"{name}/p.")})
    circ = circ.to(dev)
    inputs = {k: g(f"in.{k}").to(dev).requires_grad_(True) for k in case["inputs"]}
    state = g("state").to(dev).requires_grad_(True)
    out = circ(state, **inputs)
    ref = g("out")
    assert tuple(out.shape) == tuple(ref.shape) and out.dtype == ref.dtype
    assert float((out.detach().cpu() - ref).abs().max()) < TOL * float(ref.abs().max())
    out.backward(g("g").to(dev))

    def same(own, refg, what):
        assert own is not None, what
        assert float((own.detach().cpu() - refg).abs().max()) < TOL * max(1.0, float(refg.abs().max())), what

    for k in case["inputs"]:
        same(inputs[k].grad, g(f"gin.{k}"), f"d/d{k}")
    same(state.grad, g("grad_state"), "d/dstate")
    for k, p in circ.named_parameters():
        refg = g(f"gp.{k}")
        if torch.isnan(refg).any():  # the reference produced no gradient for it
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
        else:
            same(p.grad, refg, k)
