"""CPU oracle for the QANDLE state-vector hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module, and only as the checker / the timed CPU baseline -- never on the product path.

It restates, on CPU torch (so gradients come from torch autograd exactly like the reference's own
backward, reference SURVEY 3.3), the algorithm of /root/reference/src/qandle for the path
"apply a gate list to a (batched) state vector, measure, back-propagate".  Two restatements:

* the *dense* one (``dense_*``) follows the reference line by line: every gate is a 2^n x 2^n matrix built
  with torch.kron and applied as ``state @ M`` (operators.py:277-298, 546-558, 581-591, 662-674).  It is
  O(4^n) and only usable for n <= 10; it exists to pin the fast restatement.
* the *fast* one (``apply_*``, ``run_program``) applies the same 2x2 / permutation / sign action on a reshaped
  view in O(2^n) per gate.  It is what parity tests use at n > 10 and for complex128.

Pinning: tests/test_oracle_golden.py checks both against tests/golden/*.npz, which were produced by
importing the real reference in the build container (tests/golden/generate_golden.py).  The weight
remapping function (qw_map.tanh, un-vendored dependency) is NOT pinned: see oracle/_shim/qw_map.py.

Conventions (reference): qubit 0 is the most significant bit of the state index (operators.py:549);
states are (2^n,) or (B, 2^n); MeasureProbability returns P(qubit = 0) (measurements.py:113-123).
"""
from __future__ import annotations

import math
import typing

import torch

# ---------------------------------------------------------------------------------------------
# Gate-program IR shared with the engine (include/qandle_b200.h: QB_OP_*).
# A program is an int32 array [G, 4]: (opcode | flags, q0, q1, slot).
OP_RX, OP_RY, OP_RZ, OP_U, OP_CNOT, OP_CZ, OP_SWAP = 1, 2, 3, 4, 5, 6, 7
FLAG_BATCH = 0x100  # the angle slot indexes batch_angles[:, slot] instead of shared_angles[slot]
OP_MASK = 0xFF
MEASURE_STATE, MEASURE_PROBS, MEASURE_JOINT = 0, 1, 2


def _cdtype(real_dtype):
    return torch.complex128 if real_dtype == torch.float64 else torch.complex64


# ---------------------------------------------------------------------------------------------
# 2x2 matrices, column-vector convention  psi_out = M psi_in   (the reference stores the transposed
# kron products and right-multiplies, operators.py:236-238, 292: state @ M^T == M psi).
def rot_matrix(kind: int, theta: torch.Tensor) -> torch.Tensor:
    """(…,) angles -> (…, 2, 2) complex.  operators.py:368-395 with t = theta/2 (operators.py:267, 271):
    RX = cos(t) I - i sin(t) X;  RY = cos(t) I + sin(t) [[0,-1],[1,0]];  RZ = diag(e^{-it}, e^{+it})."""
    t = theta / 2
    c, s = torch.cos(t), torch.sin(t)
    z = torch.zeros_like(c)
    if kind == OP_RX:
        re = torch.stack([torch.stack([c, z], -1), torch.stack([z, c], -1)], -2)
        im = torch.stack([torch.stack([z, -s], -1), torch.stack([-s, z], -1)], -2)
    elif kind == OP_RY:
        re = torch.stack([torch.stack([c, -s], -1), torch.stack([s, c], -1)], -2)
        im = torch.zeros_like(re)
    elif kind == OP_RZ:
        re = torch.stack([torch.stack([c, z], -1), torch.stack([z, c], -1)], -2)
        im = torch.stack([torch.stack([-s, z], -1), torch.stack([z, s], -1)], -2)
    else:
        raise ValueError(kind)
    return torch.complex(re, im)


# ---------------------------------------------------------------------------------------------
# fast restatement
def apply_1q(state: torch.Tensor, mat: torch.Tensor, qubit: int, n: int) -> torch.Tensor:
    """Apply a 2x2 (shared, shape (2,2)) or per-sample ((B,2,2)) matrix M to `qubit` as M psi.
    state: (B, 2^n).  Same action as operators.py:289-298 (`state @ M_full^T`)."""
    B = state.shape[0]
    hi, lo = 2**qubit, 2 ** (n - qubit - 1)
    v = state.reshape(B, hi, 2, lo)
    if mat.dim() == 2:
        out = torch.einsum("ij,bhjl->bhil", mat, v)
    else:
        if mat.shape[0] != B:
            mat = mat.expand(B, 2, 2)
        out = torch.einsum("bij,bhjl->bhil", mat, v)
    return out.reshape(B, -1)


def _index_bits(n: int, device=None):
    return torch.arange(2**n, device=device)


def apply_cnot(state, c, t, n):
    """operators.py:546-555: bit position of qubit q is n-q-1; |i> -> |i ^ (1<<t2)> when bit c2 is set."""
    idx = _index_bits(n)
    c2, t2 = n - c - 1, n - t - 1
    src = torch.where((idx >> c2) & 1 == 1, idx ^ (1 << t2), idx)
    return state[:, src]


def apply_cz(state, c, t, n):
    """operators.py:581-588: -1 on indices with both bits set."""
    idx = _index_bits(n)
    c2, t2 = n - c - 1, n - t - 1
    sign = 1 - 2 * (((idx >> c2) & 1) & ((idx >> t2) & 1))
    return state * sign.to(state.real.dtype)


def apply_swap(state, a, b, n):
    """operators.py:662-671."""
    idx = _index_bits(n)
    a2, b2 = n - a - 1, n - b - 1
    diff = ((idx >> a2) & 1) != ((idx >> b2) & 1)
    src = torch.where(diff, idx ^ ((1 << a2) | (1 << b2)), idx)
    return state[:, src]


def measure_probability(state, n):
    """measurements.py:113-123: P(qubit q = 0) for every qubit, column q = qubit q.  (B, n) real.
    (The reference additionally .squeeze()s; shape quirks are the host layer's job.)"""
    B = state.shape[0]
    p = (state.real**2 + state.imag**2)
    cols = []
    for q in range(n):
        v = p.reshape(B, 2**q, 2, 2 ** (n - q - 1))
        cols.append(v[:, :, 0, :].sum(dim=(1, 2)))
    return torch.stack(cols, dim=1)


def measure_joint(state):
    """measurements.py:78-79."""
    return state.real**2 + state.imag**2


def zero_state(n, B, real_dtype=torch.float32):
    """qcircuit.py:148-149."""
    s = torch.zeros(B, 2**n, dtype=_cdtype(real_dtype))
    s[:, 0] = 1
    return s


def angle_embedding(x: torch.Tensor, qubits, rotation_kind: int, n: int) -> torch.Tensor:
    """embeddings.py:148-164: psi = (prod_q M_q(x_q)) e_0, i.e. rotation gates on |0...0>; the incoming
    state is ignored and no remapping is applied.  x: (B, d)."""
    B = x.shape[0]
    st = zero_state(n, B, x.dtype)
    for k, q in enumerate(qubits):
        st = apply_1q(st, rot_matrix(rotation_kind, x[:, k]), q, n)
    return st


def amplitude_embedding(x: torch.Tensor, n: int, normalize: bool, pad_with) -> torch.Tensor:
    """embeddings.py:50-61."""
    if pad_with is not None:
        x = torch.nn.functional.pad(x, (0, 2**n - x.shape[-1]), mode="constant", value=pad_with)
    if normalize:
        x = torch.nn.functional.normalize(x, p=2, dim=-1)
    return torch.complex(x, torch.zeros_like(x))


def run_program(
    program,
    n: int,
    shared_angles: torch.Tensor,
    batch_angles: typing.Optional[torch.Tensor],
    fixed_mats: typing.Optional[torch.Tensor],
    init_state: typing.Optional[torch.Tensor],
    batch: int,
    measure: int,
):
    """Execute the engine's gate-program IR with the fast restatement.  Angles are the already
    remapped gate angles (operators.py:271) / raw named inputs (operators.py:267).  fixed_mats[slot]
    is applied as M psi (the host passes the transpose of a `U` gate's matrix: operators.py:103-104,
    125-126 apply `state @ kron(..)` un-transposed)."""
    real_dtype = shared_angles.dtype if shared_angles is not None else torch.float32
    state = zero_state(n, batch, real_dtype) if init_state is None else init_state
    if state.dim() == 1:
        state = state.unsqueeze(0)
    if state.shape[0] != batch:
        state = state.expand(batch, -1)
    for row in program:
        code, q0, q1, slot = (int(v) for v in row)
        kind = code & OP_MASK
        if kind in (OP_RX, OP_RY, OP_RZ):
            th = batch_angles[:, slot] if code & FLAG_BATCH else shared_angles[slot]
            state = apply_1q(state, rot_matrix(kind, th), q0, n)
        elif kind == OP_U:
            state = apply_1q(state, fixed_mats[slot].to(state.dtype), q0, n)
        elif kind == OP_CNOT:
            state = apply_cnot(state, q0, q1, n)
        elif kind == OP_CZ:
            state = apply_cz(state, q0, q1, n)
        elif kind == OP_SWAP:
            state = apply_swap(state, q0, q1, n)
        else:
            raise ValueError(f"bad opcode {code}")
    if measure == MEASURE_STATE:
        return state
    if measure == MEASURE_PROBS:
        return measure_probability(state, n)
    if measure == MEASURE_JOINT:
        return measure_joint(state)
    raise ValueError(measure)


# ---------------------------------------------------------------------------------------------
# dense restatement (reference algorithm, O(4^n)); n <= 10 only.
def _hydrate(special: torch.Tensor, qubit: int, n: int) -> torch.Tensor:
    """operators.py:277-287."""
    m = torch.eye(1, dtype=special.dtype)
    for i in range(n):
        m = torch.kron(m, special if i == qubit else torch.eye(2, dtype=special.dtype))
    return m


def dense_rot(kind: int, theta: torch.Tensor, qubit: int, n: int) -> torch.Tensor:
    """Full matrix M(theta) = _a f_a(t) + _b f_b(t) (operators.py:265-275, 368-395), returned so that
    forward(state) == state @ M (i.e. already transposed as operators.py:237-238)."""
    cd = _cdtype(theta.dtype)
    t = theta / 2
    if t.dim() == 1:
        t = t.unsqueeze(-1).unsqueeze(-1)
    I2 = torch.eye(2, dtype=cd)
    if kind == OP_RX:
        a = -_hydrate(torch.tensor([[0, 1], [1, 0]], dtype=cd), qubit, n) * 1j
        b = _hydrate(I2, qubit, n)
        fa, fb = torch.sin(t), torch.cos(t)
    elif kind == OP_RY:
        a = _hydrate(torch.tensor([[0, -1], [1, 0]], dtype=cd), qubit, n)
        b = _hydrate(I2, qubit, n)
        fa, fb = torch.sin(t), torch.cos(t)
    else:
        a = _hydrate(torch.tensor([[1, 0], [0, 0]], dtype=cd), qubit, n)
        b = _hydrate(torch.tensor([[0, 0], [0, 1]], dtype=cd), qubit, n)
        fa, fb = torch.exp(-1j * t), torch.exp(1j * t)
    return a.T.contiguous() * fa + b.T.contiguous() * fb


def dense_cnot(c, t, n, cd=torch.complex64):
    """operators.py:546-555."""
    M = torch.zeros(2**n, 2**n, dtype=cd)
    c2, t2 = n - c - 1, n - t - 1
    for i in range(2**n):
        M[i, i ^ (1 << t2) if i & (1 << c2) else i] = 1
    return M


def dense_cz(c, t, n, cd=torch.complex64):
    """operators.py:581-588."""
    c2, t2 = n - c - 1, n - t - 1
    idx = torch.arange(2**n)
    diag = torch.ones(2**n, dtype=cd)
    diag[((idx & (1 << c2)) != 0) & ((idx & (1 << t2)) != 0)] = -1
    return torch.diag(diag)


def dense_swap(a, b, n, cd=torch.complex64):
    """operators.py:662-671."""
    M = torch.zeros(2**n, 2**n, dtype=cd)
    a2, b2 = n - a - 1, n - b - 1
    for i in range(2**n):
        j = i ^ ((1 << a2) | (1 << b2)) if ((i >> a2) & 1) != ((i >> b2) & 1) else i
        M[i, j] = 1
    return M


def dense_u(matrix: torch.Tensor, qubit: int, n: int, cd=torch.complex64):
    """operators.py:100-104: kron NOT transposed, applied as state @ M (operators.py:125-126)."""
    m = torch.eye(1)
    for i in range(n):
        m = torch.kron(m, matrix if i == qubit else torch.eye(2))
    return m.to(cd).contiguous()


def dense_apply(state, M):
    """operators.py:291-298."""
    if M.dim() == 2:
        return state @ M
    if state.dim() == 1:
        state = state.unsqueeze(0)
    return (state.unsqueeze(1) @ M).squeeze(1)


def dense_run_program(program, n, shared_angles, batch_angles, fixed_mats_untransposed, init_state, batch, measure):
    """Reference algorithm on the IR (dense matrices).  NOTE: fixed mats here are the user's `U`
    matrices *as given* (the reference's un-transposed application)."""
    real_dtype = shared_angles.dtype
    cd = _cdtype(real_dtype)
    state = zero_state(n, batch, real_dtype) if init_state is None else init_state
    if state.dim() == 1:
        state = state.unsqueeze(0)
    for row in program:
        code, q0, q1, slot = (int(v) for v in row)
        kind = code & OP_MASK
        if kind in (OP_RX, OP_RY, OP_RZ):
            th = batch_angles[:, slot] if code & FLAG_BATCH else shared_angles[slot]
            state = dense_apply(state, dense_rot(kind, th, q0, n))
        elif kind == OP_U:
            state = dense_apply(state, dense_u(fixed_mats_untransposed[slot], q0, n, cd))
        elif kind == OP_CNOT:
            state = dense_apply(state, dense_cnot(q0, q1, n, cd))
        elif kind == OP_CZ:
            state = dense_apply(state, dense_cz(q0, q1, n, cd))
        elif kind == OP_SWAP:
            state = dense_apply(state, dense_swap(q0, q1, n, cd))
    if state.shape[0] != batch:
        state = state.expand(batch, -1)
    if measure == MEASURE_STATE:
        return state
    if measure == MEASURE_PROBS:
        return measure_probability(state, n)
    return measure_joint(state)


# ---------------------------------------------------------------------------------------------
# Ansatz helper: gate list of the reference's StronglyEntanglingLayer
def sel_program(qubits, depth, n_rot=3, rotations=(OP_RZ, OP_RY, OP_RZ), slot0=0):
    """ansaetze/stronglyentangling.py:93-121 (built form): per depth d, for each qubit the rotation list
    with q_params[d, wi, r] (slot = slot0 + (d*nq + wi)*n_rot + r), then CNOT(q[c], q[(c + d%(nq-1) + 1) % nq])."""
    nq = len(qubits)
    rows = []
    for d in range(depth):
        for wi, w in enumerate(qubits):
            for r in range(n_rot):
                rows.append((rotations[r], w, -1, slot0 + (d * nq + wi) * n_rot + r))
        it = d % (nq - 1)
        for ci in range(nq):
            ti = (ci + it + 1) % nq
            rows.append((OP_CNOT, qubits[ci], qubits[ti], 0))
    return rows


def pi_tanh(x):
    """Assumed semantics of qw_map.tanh (un-vendored; parity unpinned)."""
    return math.pi * torch.tanh(x)
