"""torch.library contract of the custom ops (last row; run on the B200 box: `pytest -m gpu`): fake implementations,
the registered autograd formula (adjoint-state backward as ONE op call), repeated backward through one forward, and a model that
contains a Circuit under torch.compile."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from qandle_b200 import engine

    engine.load_ops()
    return engine


def _sel(n, depth, measure):
    import qandle_b200 as q

    return q.Circuit(layers=[q.AngleEmbedding(name="x", qubits=list(range(n))),
                             q.StronglyEntanglingLayer(qubits=list(range(n)), depth=depth, remapping=None), measure], num_qubits=n)


@pytest.mark.parametrize("measure", [0, 1, 2])
def test_opcheck_circuit_forward(eng, measure):
    """torch.library.opcheck: schema (no hidden mutation / aliasing), fake implementation vs the real kernel's shapes and dtypes,
    autograd registration, and the AOT-dispatch path torch.compile uses."""
    n, B = 5, 3
    prog = torch.tensor([[1, 0, -1, 0], [2 | 0x100, 1, -1, 0], [5, 0, 1, 0], [3, 2, -1, 1], [6, 1, 4, 0], [2, 3, -1, 2], [5, 4, 2, 0]], dtype=torch.int32)
    plan = eng.Plan(prog, n, eng.C64, (0, 0, 0, 0, 0, 0, 1 if measure == 1 else 0))
    gen = torch.Generator().manual_seed(0)
    shared = torch.rand(3, generator=gen).cuda().requires_grad_(True)
    batch = torch.rand(B, 1, generator=gen).cuda().requires_grad_(True)
    mats = torch.zeros(0, device="cuda")
    init = torch.nn.functional.normalize(torch.complex(torch.randn(B, 2**n, generator=gen), torch.randn(B, 2**n, generator=gen)), dim=1).cuda()
    for ini in (None, init.clone().requires_grad_(True)):
        torch.library.opcheck(torch.ops.qandle_b200.circuit_forward.default, (plan.handle, shared, batch, mats, ini, B, n, measure),
                              test_utils=("test_schema", "test_faketensor", "test_autograd_registration", "test_aot_dispatch_dynamic"))


def test_backward_can_run_twice_and_leaves_the_forward_state_intact(eng):
    import qandle_b200 as q

    n, B = 13, 3
    torch.manual_seed(1)
    circ = _sel(n, 2, q.MeasureProbability()).to("cuda")
    x = torch.rand(B, n, device="cuda", requires_grad=True)
    out = circ(x=x)
    g = torch.randn(B, n, device="cuda")
    params = list(circ.parameters())
    a = torch.autograd.grad(out, params + [x], g, retain_graph=True)
    b = torch.autograd.grad(out, params + [x], g, retain_graph=True)
    c = torch.autograd.grad(out, params + [x], 2 * g)
    for u, v, w in zip(a, b, c):
        assert torch.equal(u, v)
        assert torch.allclose(w, 2 * u, rtol=1e-4, atol=1e-6)


def test_torch_compile_model_with_circuit_matches_eager(eng):
    """A hybrid model (Linear -> Circuit -> Linear) under torch.compile: same outputs and gradients as eager, and the engine call is
    IN a captured graph (the custom op has a fake implementation and a registered autograd formula; nothing falls back to eager
    because an op could not be traced)."""
    import qandle_b200 as q

    n, B = 6, 8

    class Hybrid(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.pre = torch.nn.Linear(4, n)
            self.circ = _sel(n, 2, q.MeasureProbability())
            self.post = torch.nn.Linear(n, 2)

        def forward(self, inp):
            return self.post(self.circ(x=torch.tanh(self.pre(inp))))

    torch.manual_seed(2)
    model = Hybrid().to("cuda")
    inp = torch.randn(B, 4, device="cuda")
    ref = model(inp)
    ref.square().sum().backward()
    ref_grads = [p.grad.clone() for p in model.parameters()]
    for p in model.parameters():
        p.grad = None

    graphs = []

    def recording_backend(gm, example_inputs):
        graphs.append(gm)
        return gm.forward

    torch._dynamo.reset()
    with torch.no_grad():
        got = torch.compile(model, backend=recording_backend)(inp)
    assert torch.allclose(got, ref.detach(), rtol=1e-5, atol=1e-6)
    targets = [str(nd.target) for gm in graphs for nd in gm.graph.nodes if nd.op == "call_function"]
    assert any("qandle_b200.circuit_forward" in t for t in targets), f"engine call was not captured; graphs hold {sorted(set(targets))[:20]}"

    torch._dynamo.reset()
    cm = torch.compile(model, backend="aot_eager")  # AOTAutograd traces the registered backward (circuit_backward) with fake tensors
    out = cm(inp)
    assert torch.allclose(out, ref.detach(), rtol=1e-5, atol=1e-6)
    out.square().sum().backward()
    for p, r in zip(model.parameters(), ref_grads):
        assert p.grad is not None and torch.allclose(p.grad, r, rtol=1e-4, atol=1e-6)


def test_cuda_graph_replay_matches_eager(eng):
    """qandle_b200.cuda_graph: forward + adjoint backward of a small hybrid layer (BASELINE config 1's shape) replayed from CUDA
    graphs -- same outputs, same gradients on the same Parameters and on the input, for fresh input values."""
    import qandle_b200 as q

    n, B = 4, 64
    torch.manual_seed(3)
    layers = [q.AngleEmbedding(name="x", qubits=list(range(n)))] + [q.RX(k) for k in range(n)] + [q.RY(k) for k in range(n)]
    layers += [q.CNOT(k, (k + 1) % n) for k in range(n)] + [q.MeasureProbability()]
    circ = q.Circuit(layers=layers, num_qubits=n).to("cuda")
    x0 = torch.rand(B, n, device="cuda", requires_grad=True)
    fast = q.cuda_graph(circ, x=x0)
    for seed in (5, 6):
        torch.manual_seed(seed)
        x = torch.rand(B, n, device="cuda", requires_grad=True)
        g = torch.randn(B, n, device="cuda")
        for p in circ.parameters():
            p.grad = None
        ref = circ(x=x)
        ref.backward(g)
        ref_grads = [p.grad.clone() for p in circ.parameters()] + [x.grad.clone()]
        for p in circ.parameters():
            p.grad = None
        x.grad = None
        out = fast(x=x)
        out.backward(g)
        assert torch.allclose(out, ref, rtol=1e-6, atol=1e-7)
        for a, b in zip([p.grad for p in circ.parameters()] + [x.grad], ref_grads):
            assert a is not None and torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError):
        fast(x=torch.rand(B + 1, n, device="cuda"))
