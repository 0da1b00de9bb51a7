"""Packed-weight StronglyEntanglingLayer: same gate program as the per-gate-Parameter module."""
import pytest
import torch

import qandle_b200 as q
from qandle_b200 import qcircuit


def test_packed_sel_lowers_to_the_same_program():
    torch.manual_seed(0)
    w = torch.rand(3, 4, 3)
    a = q.Circuit(layers=[q.AngleEmbedding(name="x", qubits=[0, 1, 2, 3]), q.StronglyEntanglingLayer(qubits=[0, 1, 2, 3], depth=3, q_params=w),
                          q.RX(2), q.MeasureProbability()], num_qubits=4)
    b = q.Circuit(layers=[q.AngleEmbedding(name="x", qubits=[0, 1, 2, 3]), q.StronglyEntanglingLayerPacked(qubits=[0, 1, 2, 3], depth=3, q_params=w),
                          q.RX(2), q.MeasureProbability()], num_qubits=4)
    sa, sb = qcircuit.lower_modules(a.circuit.layers, 4)[0], qcircuit.lower_modules(b.circuit.layers, 4)[0]
    assert sa.rows == sb.rows
    assert sa.n_slots == sb.n_slots == 37
    assert len(list(b.parameters())) == 2 and len(list(a.parameters())) == 37
    with torch.no_grad():
        list(b.parameters())[1].fill_(0.25)
        list(a.parameters())[-1].fill_(0.25)
    ga = qcircuit._gather_weights(sa, torch.device("cpu"), torch.float32)
    gb = qcircuit._gather_weights(sb, torch.device("cpu"), torch.float32)
    assert torch.allclose(ga, gb)


@pytest.mark.gpu
def test_packed_sel_matches_per_gate_sel_on_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    torch.manual_seed(1)
    n, depth, B = 7, 4, 5
    w = torch.rand(depth, n, 3)
    mk = lambda cls: q.Circuit(layers=[q.AngleEmbedding(name="x", qubits=list(range(n)), rotation="ry"),
                                       cls(qubits=list(range(n)), depth=depth, q_params=w.clone()), q.MeasureProbability()], num_qubits=n).to("cuda")
    a, b = mk(q.StronglyEntanglingLayer), mk(q.StronglyEntanglingLayerPacked)
    x = torch.rand(B, n, device="cuda")
    g = torch.randn(B, n, device="cuda")
    oa, ob = a(x=x), b(x=x)
    assert torch.allclose(oa, ob, atol=1e-6)
    oa.backward(g)
    ob.backward(g)
    ga = torch.stack([p.grad.reshape(()) for p in a.parameters()])
    gb = list(b.parameters())[0].grad.reshape(-1)
    assert torch.allclose(ga, gb, atol=1e-5)
