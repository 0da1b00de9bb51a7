"""CPU tests of the SURVEY 8f rank-4 host features (engine-evaluated circuit matrix, sampling, OpenQASM export / import).

As in test_host_lowering_oracle.py the oracle interpreter is injected as the backend (TEST INFRASTRUCTURE): everything
above the custom-op boundary is the product's own code.  The GPU versions of these tests are in test_gpu_parity.py.
"""
import math

import pytest
import torch

import qandle_b200 as q
from oracle import statevec as O
from qandle_b200 import engine, qasm, qcircuit


@pytest.fixture()
def oracle_backend(monkeypatch):
    def run_circuit(plan, shared, batch, mats, init, B, measure):
        seg, n = plan
        fm = torch.view_as_complex(mats.reshape(-1, 2, 2, 2)) if mats.numel() else None
        return O.run_program(seg.rows, n, shared, batch if batch.numel() else None, fm, init, B, measure)

    monkeypatch.setattr(engine, "require_cuda", lambda: torch.device("cpu"))
    monkeypatch.setattr(qcircuit, "_plan_for", lambda seg, n, real_dtype, dev=None: (seg, n))
    monkeypatch.setattr(engine, "run_circuit", run_circuit)


def _circuit(n=3):
    torch.manual_seed(11)
    u = torch.linalg.qr(torch.complex(torch.randn(2, 2), torch.randn(2, 2)))[0]
    layers = [q.RX(0), q.RY(1, theta=0.3, remapping=None), q.RZ(2), q.CNOT(0, 2), q.U(1, u), q.CZ(1, 2), q.SWAP(0, 1),
              q.StronglyEntanglingLayer(qubits=list(range(n)), depth=2), q.RX(1, name="a")]
    return q.Circuit(layers=layers + [q.MeasureProbability()], num_qubits=n)


def test_matrix_on_engine_equals_dense_helper(oracle_backend):
    circ = _circuit()
    a = torch.tensor(0.7)
    dense = q.Circuit(layers=list(circ.circuit.layers[:-1]), num_qubits=3).to_matrix(a=a)
    got = circ.to_matrix_engine(a=a)
    assert got.shape == (8, 8) and got.dtype == torch.complex64
    assert torch.allclose(got, dense, atol=1e-5)
    # row-vector convention: forward(s) == s @ M
    st = torch.complex(torch.randn(8), torch.randn(8))
    fwd = q.Circuit(layers=list(circ.circuit.layers[:-1]) + [q.MeasureState()], num_qubits=3)(st, a=a)
    assert torch.allclose(fwd, st @ got, atol=1e-5)
    assert torch.allclose(got @ got.conj().T, torch.eye(8, dtype=torch.complex64), atol=1e-5)
    with pytest.raises(ValueError):
        q.Circuit(layers=[q.AngleEmbedding(name="x", qubits=[0, 1])], num_qubits=2).to_matrix_engine(x=torch.rand(2))


def test_sampling_follows_the_joint_distribution(oracle_backend):
    bell = qasm.circuit_from_qasm('OPENQASM 2.0;\ninclude "qelib1.inc";\nqreg q[2];\nh q[0];\ncx q[0],q[1];\n', q.MeasureProbability())
    g = torch.Generator().manual_seed(0)
    s = bell.sample(4000, generator=g)
    assert s.shape == (4000,) and s.dtype == torch.int64
    counts = torch.bincount(s, minlength=4)
    assert counts[1] == 0 and counts[2] == 0
    assert abs(int(counts[0]) - 2000) < 200
    # batched: per-sample angles give per-sample distributions; qubit 0 is the most significant bit
    circ = q.Circuit(layers=[q.AngleEmbedding(name="x", qubits=[0, 1]), q.MeasureProbability()], num_qubits=2)
    x = torch.tensor([[0.0, 0.0], [math.pi, 0.0], [0.0, math.pi]])
    sb = circ.sample(16, x=x, generator=g)
    assert sb.shape == (3, 16)
    assert (sb[0] == 0).all() and (sb[1] == 2).all() and (sb[2] == 1).all()


def test_qasm_export_import_round_trip(oracle_backend):
    circ = _circuit()
    text2 = circ.to_openqasm2()
    assert text2.startswith("OPENQASM 2.0;") and "qreg q[3];" in text2 and "input float" not in text2
    text3 = circ.to_openqasm3()
    assert text3.startswith("OPENQASM 3.0;") and "input float a;" in text3 and "rx(a) q[1];" in text3
    # importable subset: rotations with literal angles + cx / cz / swap
    layers = [q.RX(0, theta=0.4, remapping=None), q.RY(1), q.RZ(2), q.CNOT(0, 2), q.CZ(1, 2), q.SWAP(0, 1),
              q.StronglyEntanglingLayer(qubits=[0, 1, 2], depth=2)]
    src = q.Circuit(layers=layers + [q.MeasureState()], num_qubits=3)
    back = qasm.circuit_from_qasm(src.to_openqasm2(), q.MeasureState())
    assert back.num_qubits == 3
    assert len(list(back.parameters())) == len(list(src.parameters()))
    st = torch.complex(torch.randn(2, 8), torch.randn(2, 8))
    assert torch.allclose(back(st), src(st), atol=2e-5)  # exported angles are the REMAPPED ones, imported without remapping
    out = back(st)
    out.abs().sum().backward()
    assert all(p.grad is not None for p in back.parameters())


def test_qasm_import_gate_semantics(oracle_backend):
    def state_of(body, n=2):
        c = qasm.circuit_from_qasm(f'OPENQASM 2.0;\ninclude "qelib1.inc";\nqreg q[{n}];\ncreg c[{n}];\n{body}', q.MeasureState())
        return c()

    r = 2**-0.5
    assert torch.allclose(state_of("h q[0]; cx q[0],q[1]; measure q[0] -> c[0];"), torch.tensor([r, 0, 0, r], dtype=torch.complex64), atol=1e-6)
    assert torch.allclose(state_of("x q[1];"), torch.tensor([0, 1, 0, 0], dtype=torch.complex64))
    # y|0> = i|1> on qubit 0 (MSB); s then maps |1> -> i|1>
    assert torch.allclose(state_of("y q[0]; s q[0];"), torch.tensor([0, 0, -1, 0], dtype=torch.complex64), atol=1e-6)
    # u3(theta, phi, lambda)|0> = (cos(theta/2), e^{i phi} sin(theta/2)); u2 / u1 / p / expressions with pi
    got = state_of("u3(pi/2, pi/2, 0.3) q[0];", n=1)
    assert torch.allclose(got, torch.tensor([r, 1j * r], dtype=torch.complex64), atol=1e-6)
    got = state_of("h q[0]; u1(pi/4) q[0]; p(-pi/4) q[0]; t q[0]; tdg q[0]; h q[0];", n=1)
    assert torch.allclose(got, torch.tensor([1, 0], dtype=torch.complex64), atol=1e-6)
    got = state_of("u2(0, pi) q[0];", n=1)  # = H
    assert torch.allclose(got, torch.tensor([r, r], dtype=torch.complex64), atol=1e-6)
    got = state_of("rx(2*pi/4) q[0]; sx q[0]; sdg q[0]; z q[0]; id q[0];", n=1)
    ref = torch.tensor([[1, 0], [0, -1]], dtype=torch.complex64) @ torch.tensor([[1, 0], [0, -1j]], dtype=torch.complex64) @ \
        torch.tensor(qasm._FIXED["sx"], dtype=torch.complex64) @ torch.tensor([r, -1j * r], dtype=torch.complex64)
    assert torch.allclose(got, ref, atol=1e-6)
    # toffoli, two registers laid out in declaration order
    c = qasm.circuit_from_qasm("OPENQASM 2.0; qreg a[2]; qreg b[1]; x a[0]; x a[1]; ccx a[0],a[1],b[0];", q.MeasureState())
    assert c.num_qubits == 3 and torch.allclose(c(), torch.eye(8, dtype=torch.complex64)[7], atol=1e-6)


@pytest.mark.parametrize("bad", [
    "OPENQASM 3.0; qubit[2] q;", "OPENQASM 2.0; qreg q[2]; foo q[0];", "OPENQASM 2.0; qreg q[2]; rx(1,2) q[0];",
    "OPENQASM 2.0; qreg q[2]; h q;", "OPENQASM 2.0; qreg q[2]; h q[2];", "OPENQASM 2.0; h q[0];",
    "OPENQASM 2.0; qreg q[1]; rx(__import__('os')) q[0];", "OPENQASM 2.0; qreg q[1]; gate g a { h a; }",
])
def test_qasm_import_rejects(bad):
    with pytest.raises(qasm.QasmSyntaxError):
        qasm.parse_qasm(bad)


# ---- round-2 host fixes ------------------------------------------------------------------------------------
def test_non_unitary_and_trainable_u_take_the_torch_path(oracle_backend):
    """The reference accepts any 2x2 in U and differentiates through it (operators.py:100-126); the adjoint-state backward needs
    unitary gates, so such a U runs as a torch module between engine segments -- same values, same gradients as the dense algorithm."""
    n = 3
    torch.manual_seed(3)
    m = torch.complex(torch.randn(2, 2), torch.randn(2, 2))  # not unitary
    with pytest.warns(UserWarning, match="not unitary"):
        circ = q.Circuit(layers=[q.RY(0, theta=0.4, remapping=None), q.U(1, m), q.CNOT(1, 2), q.RX(2, theta=-0.8, remapping=None),
                                 q.MeasureJointProbability()], num_qubits=n)
    out = circ()
    th = torch.tensor([0.4, -0.8], requires_grad=True)
    st = O.zero_state(n, 1)
    st = O.dense_apply(st, O.dense_rot(O.OP_RY, th[0:1], 0, n))
    st = O.dense_apply(st, O.dense_u(m, 1, n))
    st = O.dense_apply(st, O.dense_cnot(1, 2, n))
    st = O.dense_apply(st, O.dense_rot(O.OP_RX, th[1:2], 2, n))
    ref = O.measure_joint(st)[0]
    assert torch.allclose(out, ref.detach(), atol=1e-6)
    out.sum().backward()
    ref.sum().backward()
    got = torch.stack([p.grad.reshape(()) for p in circ.parameters()])
    assert torch.allclose(got, th.grad, atol=1e-5)
    # a matrix that requires grad is differentiated through, like the reference's tape does
    mu = torch.linalg.qr(torch.complex(torch.randn(2, 2), torch.randn(2, 2)))[0].clone().requires_grad_(True)
    c2 = q.Circuit(layers=[q.RY(0, theta=0.4, remapping=None), q.U(0, mu), q.MeasureProbability()], num_qubits=1)
    c2().sum().backward()
    assert mu.grad is not None and float(mu.grad.abs().max()) > 0
    with pytest.raises(NotImplementedError):
        q.Circuit(layers=[q.Invert(q.U(0, m))], num_qubits=1)()


def test_lowered_segments_follow_the_layer_list_and_stay_out_of_pickles(oracle_backend):
    import copy
    import pickle

    circ = q.Circuit(layers=[q.RX(0, theta=0.3, remapping=None), q.Invert(q.RY(1, theta=0.2, remapping=None)), q.MeasureProbability()], num_qubits=2)
    a = circ()
    blob = pickle.dumps(circ)  # after a forward: the cache holds closures / plan handles and must not be part of the module state
    b = pickle.loads(blob)()
    assert torch.allclose(a, b)
    assert torch.allclose(copy.deepcopy(circ)(), a)
    assert "_qb_segments" not in circ.circuit.__dict__
    # editing the layer list after the first call is honoured
    circ.circuit.layers.insert(2, q.RX(1, theta=1.1, remapping=None).build(2))
    c = circ()
    ref = q.Circuit(layers=[q.RX(0, theta=0.3, remapping=None), q.Invert(q.RY(1, theta=0.2, remapping=None)), q.RX(1, theta=1.1, remapping=None),
                            q.MeasureProbability()], num_qubits=2)()
    assert torch.allclose(c, ref) and not torch.allclose(c, a)


def test_size_agnostic_measurement_is_planned_per_state_size(oracle_backend):
    m = q.MeasureJointProbability()
    s2 = torch.complex(torch.randn(4), torch.randn(4))
    s3 = torch.complex(torch.randn(3, 8), torch.randn(3, 8))
    assert torch.allclose(m(s2), s2.abs() ** 2, atol=1e-6)
    assert torch.allclose(m(s3), s3.abs() ** 2, atol=1e-6)


def test_measurement_squeeze_applies_before_a_following_torch_module(oracle_backend):
    """reference measurements.py:123: MeasureProbability returns a squeezed tensor to whatever comes next in the layer list."""
    seen = {}

    class Probe(torch.nn.Module):
        def forward(self, x):
            seen["shape"] = tuple(x.shape)
            return x * 2

    circ = q.Circuit(layers=[q.RX(0, theta=0.3, remapping=None), q.RX(1, theta=0.5, remapping=None), q.MeasureProbability(), Probe()], num_qubits=2)
    out = circ()
    assert seen["shape"] == (2,) and tuple(out.shape) == (2,)
    out1 = q.Circuit(layers=[q.RX(0, theta=0.3, remapping=None), q.MeasureProbability(), Probe()], num_qubits=1)()
    assert seen["shape"] == () and tuple(out1.shape) == ()


def test_qasm_angle_expressions_are_parsed_not_evaluated():
    assert abs(qasm._eval_angle("-3*pi/4 + sin(0.5)^2") - (-3 * math.pi / 4 + math.sin(0.5) ** 2)) < 1e-12
    for bad in ("9^9^9^9", "__import__('os')", "pi.real", "1/0", "[1]", "x" * 300):
        with pytest.raises(qasm.QasmSyntaxError):
            qasm._eval_angle(bad)
