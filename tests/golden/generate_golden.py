"""Generate golden vectors by running the REAL reference (gstenzel/qandle, /root/reference/src) on CPU.

Run in the build container only (the reference does not travel to the GPU box):

    python tests/golden/generate_golden.py

It puts oracle/_shim (a stand-in for the un-vendored qW-Map dependency) and /root/reference/src on
sys.path, imports `qandle`, and writes
  tests/golden/api_cases.npz   circuits built through the reference's public API (specs.py)
  tests/golden/ir_cases.npz    circuits built from the engine's gate-program IR, remapping=None,
                               incl. gradients w.r.t. every angle, named input and the initial state
Nothing here is imported by the product.
"""
import json
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_shim"))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import qandle  # noqa: E402  (the reference)
import specs  # noqa: E402
from oracle import statevec as O  # noqa: E402  (only for the IR opcode constants)


def rand_state(shape, gen):
    st = torch.complex(torch.rand(*shape, generator=gen), torch.rand(*shape, generator=gen) - 0.5)
    return st / torch.linalg.norm(st, dim=-1, keepdim=True)


def run_api_case(name, case, idx):
    torch.manual_seed(1000 + idx)
    gen = torch.Generator().manual_seed(2000 + idx)
    circ = specs.build_circuit(qandle, case["spec"], case["num_qubits"])
    n = circ.num_qubits
    inputs = {k: torch.rand(tuple(shape), generator=gen).requires_grad_(True) for k, shape in case["inputs"].items()}
    state = None
    if case["state"] == "batched":
        state = rand_state((case["batch"], 2**n), gen).requires_grad_(True)
    elif case["state"] == "unbatched":
        state = rand_state((2**n,), gen).requires_grad_(True)
    out = circ(state, **inputs)
    g = torch.randn(out.shape, generator=gen, dtype=out.dtype) if out.is_complex() else torch.randn(out.shape, generator=gen)
    out.backward(g)
    rec = {
        "spec": np.array(specs.dumps(case["spec"])),
        "num_qubits": np.array(-1 if case["num_qubits"] is None else case["num_qubits"]),
        "n": np.array(n),
        "out": out.detach().numpy(),
        "g": g.numpy(),
        "has_state": np.array(state is not None),
    }
    if state is not None:
        rec["state"] = state.detach().numpy()
        rec["grad_state"] = state.grad.numpy()
    for k, v in inputs.items():
        rec[f"in.{k}"] = v.detach().numpy()
        rec[f"gin.{k}"] = v.grad.numpy() if v.grad is not None else np.zeros(v.shape, np.float32)
    for k, p in circ.named_parameters():
        rec[f"p.{k}"] = p.detach().numpy()
        rec[f"gp.{k}"] = p.grad.numpy() if p.grad is not None else np.full(p.shape, np.nan, np.float32)
    return rec


def random_program(rng, n, G, n_shared, n_batch, n_mats):
    rows = []
    for _ in range(G):
        r = rng.random()
        if n >= 2 and r < 0.35:
            a, b = rng.sample(range(n), 2)
            rows.append((rng.choice([O.OP_CNOT, O.OP_CZ, O.OP_SWAP]), a, b, 0))
        elif n_mats and r < 0.45:
            rows.append((O.OP_U, rng.randrange(n), -1, rng.randrange(n_mats)))
        elif n_batch and r < 0.6:
            rows.append((rng.choice([O.OP_RX, O.OP_RY, O.OP_RZ]) | O.FLAG_BATCH, rng.randrange(n), -1, rng.randrange(n_batch)))
        else:
            rows.append((rng.choice([O.OP_RX, O.OP_RY, O.OP_RZ]), rng.randrange(n), -1, rng.randrange(n_shared)))
    return rows


def random_unitary(gen):
    a = torch.complex(torch.randn(2, 2, generator=gen), torch.randn(2, 2, generator=gen))
    q, r = torch.linalg.qr(a)
    return (q * (torch.diagonal(r) / torch.diagonal(r).abs())).to(torch.complex64)


def run_ir_case(idx, n, G, B, measure, with_init, n_shared=5, n_batch=2, n_mats=2):
    rng = random.Random(500 + idx)
    gen = torch.Generator().manual_seed(3000 + idx)
    prog = random_program(rng, n, G, n_shared, n_batch, n_mats)
    shared = ((torch.rand(n_shared, generator=gen) - 0.5) * 4 * np.pi).requires_grad_(True)
    batch = ((torch.rand(B, max(n_batch, 1), generator=gen) - 0.5) * 4 * np.pi).requires_grad_(True)
    mats_user = torch.stack([random_unitary(gen) for _ in range(max(n_mats, 1))])  # as the user passes to qandle.U
    init = rand_state((B, 2**n), gen).requires_grad_(True) if with_init else None
    # the same slot may be used by several gates: give each gate its own theta tensor that is a view of
    # the slot so torch sums the gradients, exactly as shared weights would behave.
    layers = []
    kw_inputs = {}
    RCLS = {O.OP_RX: qandle.RX, O.OP_RY: qandle.RY, O.OP_RZ: qandle.RZ}
    state = init
    if state is None:
        state = torch.zeros(2**n, dtype=torch.complex64)
        state[0] = 1
    # apply built gates one at a time so that non-leaf thetas (views of `shared`) keep their graph
    for code, q0, q1, slot in prog:
        kind = code & O.OP_MASK
        if kind in RCLS:
            if code & O.FLAG_BATCH:
                gate = RCLS[kind](qubit=q0, name="a", remapping=None).build(num_qubits=n)
                state = gate(state, a=batch[:, slot])
            else:
                gate = RCLS[kind](qubit=q0, theta=torch.tensor(0.0), remapping=None).build(num_qubits=n)
                # reference math with an externally owned angle: get_matrix path, operators.py:265-275
                t = shared[slot] / 2
                mat = gate._a * gate.a_op(t) + gate._b * gate.b_op(t)
                state = state @ mat
        elif kind == O.OP_U:
            state = qandle.U(qubit=q0, matrix=mats_user[slot]).build(num_qubits=n)(state)
        elif kind == O.OP_CNOT:
            state = qandle.CNOT(q0, q1).build(num_qubits=n)(state)
        elif kind == O.OP_CZ:
            state = qandle.CZ(q0, q1).build(num_qubits=n)(state)
        elif kind == O.OP_SWAP:
            state = qandle.SWAP(q0, q1).build(num_qubits=n)(state)
    if state.dim() == 1:
        state = state.unsqueeze(0).expand(B, -1)
    if measure == O.MEASURE_PROBS:
        out = qandle.MeasureProbability().build(num_qubits=n)(state).reshape(B, n)
    elif measure == O.MEASURE_JOINT:
        out = qandle.MeasureJointProbability()(state)
    else:
        out = state
    g = torch.randn(out.shape, generator=gen, dtype=out.dtype) if out.is_complex() else torch.randn(out.shape, generator=gen)
    out.backward(g)
    z = lambda t: t.grad.numpy() if t.grad is not None else np.zeros(t.shape, np.float32)
    rec = {
        "n": np.array(n), "B": np.array(B), "measure": np.array(measure),
        "program": np.array(prog, dtype=np.int32).reshape(-1, 4),
        "shared": shared.detach().numpy(), "batch": batch.detach().numpy(),
        "mats_user": mats_user.numpy(), "has_init": np.array(with_init),
        "out": out.detach().numpy(), "g": g.numpy(),
        "g_shared": z(shared), "g_batch": z(batch),
    }
    if with_init:
        rec["init"] = init.detach().numpy()
        rec["g_init"] = init.grad.numpy()
    return rec


def main():
    api = {}
    for idx, (name, case) in enumerate(specs.api_specs().items()):
        rec = run_api_case(name, case, idx)
        for k, v in rec.items():
            api[f"{name}/{k}"] = v
        print("api", name, rec["out"].shape)
    np.savez_compressed(os.path.join(HERE, "api_cases.npz"), **api)

    ir = {}
    cases = []
    idx = 0
    for n in (1, 2, 3, 4, 5):
        for measure in (O.MEASURE_STATE, O.MEASURE_PROBS, O.MEASURE_JOINT):
            cases.append((n, 12 + 4 * n, 3, measure, True))
    cases += [(6, 60, 5, O.MEASURE_PROBS, False), (7, 80, 2, O.MEASURE_STATE, True), (8, 90, 1, O.MEASURE_JOINT, False),
              (9, 50, 2, O.MEASURE_PROBS, True), (10, 24, 2, O.MEASURE_PROBS, False)]
    for (n, G, B, measure, with_init) in cases:
        rec = run_ir_case(idx, n, G, B, measure, with_init)
        for k, v in rec.items():
            ir[f"ir{idx}/{k}"] = v
        print("ir", idx, n, G, B, measure, with_init)
        idx += 1
    ir["count"] = np.array(idx)
    np.savez_compressed(os.path.join(HERE, "ir_cases.npz"), **ir)

    # the reference's model-level caller: QConv (reference convolution.py)
    torch.manual_seed(77)
    conv = qandle.QConv(in_channels=3, out_channels=4, kernel_size=3, padding=1, qdepth=2)
    gen = torch.Generator().manual_seed(78)
    x = torch.rand(2, 3, 5, 5, generator=gen).requires_grad_(True)
    out = conv(x)
    g = torch.randn(out.shape, generator=gen)
    out.backward(g)
    rec = {"x": x.detach().numpy(), "out": out.detach().numpy(), "g": g.numpy(), "gx": x.grad.numpy()}
    for k, p in conv.named_parameters():
        rec[f"p.{k}"] = p.detach().numpy()
        rec[f"gp.{k}"] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "qconv_case.npz"), **rec)
    print("qconv", out.shape)


if __name__ == "__main__":
    main()
