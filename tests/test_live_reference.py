"""Random circuits, side by side with the UNMODIFIED reference running in the same process.

The golden vectors (tests/golden/*.npz) pin a fixed list of circuits; this file draws new ones.  `oracle/_ref` is the
reference's Python package as copied by oracle/make_ref.py (git-ignored, travels to the GPU box with the snapshot; never
imported by qandle_b200).  For every seed one spec (tests/golden/specs.py: the same constructor calls for both packages)
is instantiated twice -- `import qandle` (reference: dense 2^n x 2^n matrix per gate, torch autograd) and `qandle_b200` --
the reference's state_dict is loaded into ours, both run forward and backward on the same inputs / cotangent, and output
shape, dtype, values, input / initial-state gradients and every parameter gradient must agree to 1e-5 of the largest
reference entry (complex64; the north_star tolerance).

Backends: `oracle` (CPU, `-m "not gpu"`: oracle interpreter injected below the custom-op boundary, so the host path is
what is compared) and `engine` (`-m gpu`: the CUDA engine through the torch ops; the reference runs on the host cores).
"""
import os
import random

import pytest
import torch

import specs
import qandle_b200 as q
from oracle import make_ref
from oracle import statevec as O
from qandle_b200 import engine, qcircuit

R = make_ref.import_reference()
if R is None:
    pytest.skip("oracle/_ref was never made (python -m oracle.make_ref in the build container)", allow_module_level=True)

TOL = 1e-5
NONE = specs.NONE
SEEDS = list(range(int(os.environ.get("QB_LIVE_SEEDS", "32"))))  # more seeds for a soak run


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


def _unitary(rng):
    g = torch.Generator().manual_seed(rng.randrange(1 << 30))
    u = torch.linalg.qr(torch.complex(torch.randn(2, 2, generator=g), torch.randn(2, 2, generator=g)))[0]
    return specs.T(torch.view_as_real(u).tolist(), "complex64")


def random_case(seed):
    """-> (spec, num_qubits, inputs {name: shape}, state None | 'batched' | 'unbatched', batch)"""
    rng = random.Random(1000 + seed)
    n = rng.randint(2, 8)
    B = rng.choice([1, 2, 3, 5, 8])
    qs = list(range(n))
    spec, inputs = [], {}
    head = rng.choice(["angle", "amp", "state_b", "state_u", "zero"])
    state = None
    if head == "angle":
        spec.append(("AngleEmbedding", {"name": "x", "qubits": qs, "rotation": rng.choice(["rx", "ry", "rz"])}))
        inputs["x"] = [B, n] if rng.random() < 0.8 else [n]
    elif head == "amp":
        pad = rng.random() < 0.4 and n >= 2
        kw = {"name": "a", "qubits": qs, "normalize": True}
        if pad:
            kw["pad_with"] = 0
        spec.append(("AmplitudeEmbedding", kw))
        width = rng.randint(2 ** (n - 1) + 1, 2**n - 1) if pad else 2**n
        inputs["a"] = [B, width] if rng.random() < 0.7 else [width]
    elif head.startswith("state"):
        state = "batched" if head == "state_b" else "unbatched"
    batched = (state == "batched") or any(len(s) == 2 for s in inputs.values())
    rot = lambda: rng.choice(["RX", "RY", "RZ"])  # noqa: E731
    remap = lambda: {} if rng.random() < 0.5 else {"remapping": NONE}  # noqa: E731
    named = no_batch = False
    for _ in range(rng.randint(3, 14)):
        kind = rng.choice(["fixed", "fixed", "train", "train", "named", "cnot", "cnot", "cz", "swap", "u", "sel", "invert", "ctrl", "reset", "twolocal", "su"])
        a = rng.randrange(n)
        b = rng.choice([w for w in qs if w != a])
        if kind == "fixed":
            spec.append((rot(), dict(qubit=a, theta=rng.uniform(-3, 3), **remap())))
        elif kind == "train":
            spec.append((rot(), dict(qubit=a, **remap())))
        elif kind == "named":
            spec.append((rot(), dict(qubit=a, name="phi", **remap())))
            named = True
        elif kind == "cnot":
            spec.append(("CNOT", {"control": a, "target": b}))
        elif kind == "cz":
            spec.append(("CZ", {"control": a, "target": b}))
        elif kind == "swap":
            spec.append(("SWAP", {"a": a, "b": b}))
        elif kind == "u":
            spec.append(("U", {"qubit": a, "matrix": _unitary(rng)}))
        elif kind == "sel":
            sub = sorted(rng.sample(qs, rng.randint(2, n)))
            spec.append(("StronglyEntanglingLayer", dict(qubits=sub, depth=rng.randint(1, 3), **remap())))
        elif kind == "reset":  # non-unitary: a torch module between two engine segments
            spec.append(("Reset", {"qubit": a}))
        elif kind == "twolocal":
            if batched:  # the reference's TwoLocal forward is `matrix @ state` (twolocal.py:102): unbatched states only
                continue
            no_batch = True
            sub = sorted(rng.sample(qs, rng.randint(2, n)))
            spec.append(("TwoLocal", dict(qubits=sub, **remap())))
        elif kind == "su":
            if batched:  # same limit in the reference's SU forward (specialunitary.py:106)
                continue
            no_batch = True
            sub = sorted(rng.sample(qs, rng.randint(2, n)))
            spec.append(("SU", dict(qubits=sub, reps=rng.randint(0, 2), rotations=rng.choice([["ry"], ["rz", "ry"], ["rx", "ry"]]))))
        elif kind == "invert":
            spec.append(("Invert", {"target": {"__op__": [rot(), dict(qubit=a, theta=rng.uniform(-2, 2), **remap())]}}))
        elif kind == "ctrl":
            spec.append(("Controlled", {"control": a, "target": {"__op__": [rot(), dict(qubit=b, theta=rng.uniform(-2, 2), **remap())]}}))
    if named:
        # per-sample values with a batch, else a scalar; an unbatched run with a batched named input is the auto-batching quirk Q7
        inputs["phi"] = [B] if (batched or (rng.random() < 0.3 and not no_batch)) else []
    meas = rng.choice(["MeasureProbability", "MeasureProbability", "MeasureJointProbability", "MeasureState", None])
    if meas:
        spec.append((meas, {}))
    return spec, n, inputs, state, B


def _rel(a, b):
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)


@pytest.mark.parametrize("seed", SEEDS)
def test_random_circuit_matches_the_running_reference(dev, seed):
    spec, n, in_shapes, state_kind, B = random_case(seed)
    ref = specs.build_circuit(R, spec, n)
    own = specs.build_circuit(q, spec, n)
    assert sorted(own.state_dict().keys()) == sorted(ref.state_dict().keys())
    own.load_state_dict(ref.state_dict())
    own = own.to(dev)

    g = torch.Generator().manual_seed(seed)
    host_in = {k: torch.rand(tuple(s), generator=g) * 2 - 0.7 for k, s in in_shapes.items()}
    host_state = None
    if state_kind is not None:
        st = torch.complex(torch.randn(*(([B] if state_kind == "batched" else []) + [2**n]), generator=g),
                           torch.randn(*(([B] if state_kind == "batched" else []) + [2**n]), generator=g))
        host_state = st / torch.linalg.norm(st, dim=-1, keepdim=True)

    def run(circ, device):
        ins = {k: v.clone().to(device).requires_grad_(True) for k, v in host_in.items()}
        st = None if host_state is None else host_state.clone().to(device).requires_grad_(True)
        out = circ(st, **ins)
        return out, ins, st

    r_out, r_in, r_st = run(ref, torch.device("cpu"))
    o_out, o_in, o_st = run(own, dev)
    assert tuple(o_out.shape) == tuple(r_out.shape), (spec, o_out.shape, r_out.shape)
    assert o_out.dtype == r_out.dtype
    assert _rel(o_out.detach().cpu(), r_out.detach()) < TOL, spec

    if not r_out.requires_grad:  # only parameter-free gates on a constant state: nothing to differentiate
        return
    cot = torch.randn(r_out.shape, generator=g)
    if r_out.is_complex():
        cot = torch.complex(cot, torch.randn(r_out.shape, generator=g))
    r_out.backward(cot)
    o_out.backward(cot.to(dev))

    # Reset divides by |a0| + 1e-7 per amplitude pair (reference operators.py:617): gradients through it amplify float32 rounding
    # of either implementation by 1 / |a0| (seed 44: the reference itself is 2e-6 off the float64 oracle there), so 1e-4 for those
    gtol = 1e-4 if any(name == "Reset" for name, _ in spec) else TOL

    def same_grad(own_g, ref_g, what, scale=None):
        if ref_g is not None and bool(torch.isnan(ref_g).any()):
            return  # the reference's own gradient is undefined here (its Reset divides by a norm that can be 0)
        if ref_g is None or not bool(ref_g.abs().max() > 0):
            assert own_g is None or float(own_g.abs().max()) < TOL, what
            return
        assert own_g is not None, what
        scale = float(ref_g.abs().max()) if scale is None else scale
        assert float((own_g.detach().cpu() - ref_g).abs().max()) < gtol * max(1.0, scale), (what, spec)

    for k in host_in:
        same_grad(o_in[k].grad, r_in[k].grad, f"d/d{k}")
    if host_state is not None:
        same_grad(o_st.grad, r_st.grad, "d/dstate")
    # the parameter gradient is ONE vector (the reference keeps a 0-dim Parameter per gate): error relative to its largest entry
    r_params = dict(ref.named_parameters())
    pmax = max([float(p.grad.abs().max()) for p in r_params.values() if p.grad is not None and not torch.isnan(p.grad).any()] + [0.0])
    for k, p in own.named_parameters():
        same_grad(p.grad, r_params[k].grad, k, scale=pmax)


@pytest.mark.parametrize("seed", list(range(64)))
@pytest.mark.parametrize("version", [2, 3])
def test_qasm_export_matches_the_running_reference(seed, version):
    """OpenQASM 2 / 3 text of the same random circuits, statement for statement (reference qasm.py `convert_to_qasm`); the
    reference names embedding inputs and measurement outputs with a random hash, which is masked.  Cases the reference cannot
    export (AmplitudeEmbedding: NotImplementedError; MeasureState / MeasureJointProbability: AttributeError) are not compared."""
    import re

    spec, n, *_ = random_case(seed)
    ref = specs.build_circuit(R, spec, n)
    own = specs.build_circuit(q, spec, n)
    own.load_state_dict(ref.state_dict())
    try:
        want = R.convert_to_qasm(ref, qasm_version=version, include_header=True)
    except (NotImplementedError, AttributeError) as e:
        pytest.skip(f"the reference cannot export this circuit: {type(e).__name__}")
    got = q.convert_to_qasm(own, qasm_version=version, include_header=True)
    mask = lambda t: re.sub(r"(angle|measured)_-?\d+_", r"\1_H_", t)  # noqa: E731
    assert mask(got) == mask(want)


@pytest.mark.parametrize("ci,co,k,pad", [(ci, co, k, pad) for ci in (1, 2, 3) for co in (1, 3, 4) for k in (2, 3) for pad in (0, 1)])
def test_qconv_matches_the_running_reference(monkeypatch, ci, co, k, pad):
    """QConv (reference convolution.py) over a grid of channel / kernel / padding choices: state_dict keys, output shape and dtype,
    values, input gradient and every weight gradient against the running reference (CPU backend; the engine version of this
    check is the reference-generated golden case in test_gpu_parity.py::test_qconv_matches_reference)."""
    def run_circuit(plan, shared, batch, mats, init, B, measure):
        seg, n = plan
        fm = torch.view_as_complex(mats.reshape(-1, 2, 2, 2)) if mats.numel() else None
        return O.run_program(seg.rows, n, shared, batch if batch.numel() else None, fm, init, B, measure)

    monkeypatch.setattr(engine, "require_cuda", lambda: torch.device("cpu"))
    monkeypatch.setattr(qcircuit, "_plan_for", lambda seg, n, real_dtype, dev=None: (seg, n))
    monkeypatch.setattr(engine, "run_circuit", run_circuit)
    torch.manual_seed(ci * 100 + co * 10 + k + pad)
    ref = R.QConv(in_channels=ci, out_channels=co, kernel_size=k, padding=pad)
    own = q.QConv(in_channels=ci, out_channels=co, kernel_size=k, padding=pad)
    assert sorted(own.state_dict()) == sorted(ref.state_dict())
    own.load_state_dict(ref.state_dict())
    x = torch.rand(2, ci, 5, 6)
    xr, xo = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    yr, yo = ref(xr), own(xo)
    assert yo.shape == yr.shape and yo.dtype == yr.dtype
    assert _rel(yo.detach(), yr.detach()) < TOL
    g = torch.randn(yr.shape)
    yr.backward(g)
    yo.backward(g)
    assert _rel(xo.grad, xr.grad) < TOL
    pr = dict(ref.named_parameters())
    pmax = max(float(p.grad.abs().max()) for p in pr.values())
    for name, p in own.named_parameters():
        assert float((p.grad - pr[name].grad).abs().max()) < TOL * max(1.0, pmax), name


@pytest.mark.parametrize("seed", list(range(24)))
def test_sel_variants_match_the_running_reference(monkeypatch, seed):
    """StronglyEntanglingLayer with other rotation sets than the default rz-ry-rz, given `q_params`, sub-sets of the qubits
    in any order and both remappings (reference stronglyentangling.py:23-39, 81-110) -- CPU backend; the engine sees the same
    gate program rows as for the default layer."""
    def run_circuit(plan, shared, batch, mats, init, B, measure):
        seg, n = plan
        fm = torch.view_as_complex(mats.reshape(-1, 2, 2, 2)) if mats.numel() else None
        return O.run_program(seg.rows, n, shared, batch if batch.numel() else None, fm, init, B, measure)

    monkeypatch.setattr(engine, "require_cuda", lambda: torch.device("cpu"))
    monkeypatch.setattr(qcircuit, "_plan_for", lambda seg, n, real_dtype, dev=None: (seg, n))
    monkeypatch.setattr(engine, "run_circuit", run_circuit)
    rng = random.Random(9000 + seed)
    n = rng.randint(2, 7)
    sub = rng.sample(range(n), rng.randint(2, n))
    rots = rng.choice([["rx"], ["ry", "rz"], ["rx", "ry", "rx"], ["rz", "rx", "rz", "ry"], ["rz", "ry", "rz"]])
    depth = rng.randint(1, 4)
    kw = dict(qubits=sub, depth=depth, rotations=rots)
    if rng.random() < 0.5:
        kw["remapping"] = NONE
    if rng.random() < 0.5:
        kw["q_params"] = specs.T([[[rng.uniform(-2, 2) for _ in rots] for _ in sub] for _ in range(depth)])
    spec = [("StronglyEntanglingLayer", kw), (rng.choice(["MeasureProbability", "MeasureState"]), {})]
    torch.manual_seed(seed)
    ref = specs.build_circuit(R, spec, n)
    own = specs.build_circuit(q, spec, n)
    assert sorted(own.state_dict().keys()) == sorted(ref.state_dict().keys())
    own.load_state_dict(ref.state_dict())
    g = torch.Generator().manual_seed(seed)
    st = torch.complex(torch.randn(3, 2**n, generator=g), torch.randn(3, 2**n, generator=g))
    st = st / torch.linalg.norm(st, dim=-1, keepdim=True)
    sr, so = st.clone().requires_grad_(True), st.clone().requires_grad_(True)
    yr, yo = ref(sr), own(so)
    assert yo.shape == yr.shape and yo.dtype == yr.dtype
    assert _rel(yo.detach(), yr.detach()) < TOL
    cot = torch.randn(yr.shape, generator=g)
    if yr.is_complex():
        cot = torch.complex(cot, torch.randn(yr.shape, generator=g))
    yr.backward(cot)
    yo.backward(cot)
    assert _rel(so.grad, sr.grad) < TOL
    pr = dict(ref.named_parameters())
    pmax = max(float(p.grad.abs().max()) for p in pr.values())
    for name, p in own.named_parameters():
        assert float((p.grad - pr[name].grad).abs().max()) < TOL * max(1.0, pmax), name
