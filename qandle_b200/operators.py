"""Gate operators: same public surface as reference src/qandle/operators.py, no dense matrices.

The reference builds a 2^n x 2^n matrix per gate (operators.py:277-287, 546-555) and applies it with
``state @ M`` (operators.py:289-298).  Here a built gate is only a *description* (qubit(s), Parameter,
remapping, name); applying it -- alone or inside a Circuit -- lowers it to the engine's gate program and
runs the sm_100a sweep kernels.  Constructors, attributes, Parameter names/shapes and error behaviour follow
the reference so existing user code and checkpoints keep working.

Scope: RX, RY, RZ, U/CustomGate, CNOT, CZ, SWAP; from SURVEY 8f rank 3: Invert and Controlled (lowered to
engine ops) and Reset (non-unitary: runs in torch between two engine segments).
"""
from __future__ import annotations

import abc
import cmath
import dataclasses
import math
import typing
import warnings

import torch

from . import config, errors, remap

__all__ = [
    "Operator", "UnbuiltOperator", "BuiltOperator", "RX", "RY", "RZ", "CNOT", "CZ", "SWAP", "U", "CustomGate",
    "BuiltRX", "BuiltRY", "BuiltRZ", "BuiltU", "BuiltCNOT", "BuiltCZ", "BuiltSWAP", "BUILT_CLASS_RELATION",
    "Reset", "BuiltReset", "Invert", "BuiltInvert", "Controlled", "BuiltControlled",
    "QasmRepresentation",
]


@dataclasses.dataclass
class QasmRepresentation:
    """Minimal stand-in for reference qasm.py:6-31 (string export is out of the engine's scope)."""

    gate_str: str
    qubit: typing.Union[int, str, None] = ""
    gate_value: typing.Union[float, None] = None
    qasm3_inputs: str = ""
    qasm3_outputs: str = ""


class Operator(abc.ABC):
    """Everything that can be applied to a state (reference operators.py:35-49)."""

    named = False

    @abc.abstractmethod
    def __str__(self) -> str: ...

    def __repr__(self) -> str:
        return self.__str__()

    def to_qasm(self) -> QasmRepresentation:
        raise NotImplementedError("OpenQASM export is outside the scope of the B200 engine (SURVEY.md 2)")


class UnbuiltOperator(Operator, abc.ABC):
    """Operator specification; ``build(num_qubits)`` turns it into an nn.Module (reference operators.py:52-57)."""

    @abc.abstractmethod
    def build(self, num_qubits, **kwargs) -> "BuiltOperator": ...


class BuiltOperator(Operator, torch.nn.Module, abc.ABC):
    """Built operator.  ``forward`` runs the single gate through the engine (reference operators.py:60-69)."""

    num_qubits: int

    def forward(self, state: torch.Tensor, **kwargs) -> torch.Tensor:
        from . import qcircuit

        return qcircuit.run_modules(self, [self], self.num_qubits, state, kwargs)

    def to_matrix(self, **kwargs) -> torch.Tensor:
        """Dense matrix M with ``forward(state) == state @ M`` (reference operators.py:67-69).  Lazy O(4^n)
        helper for small n only -- it is not on the accelerated path."""
        raise NotImplementedError


def chain_matrices(owner: torch.nn.Module, mods, **kwargs) -> torch.Tensor:
    """Product of the gates' dense matrices in application order (row-vector convention: state @ M_1 @ M_2 ...), on the
    device the owner's parameters live on: parameter-free gates (CNOT / CZ / SWAP) build theirs on the CPU."""
    dev = next((p.device for p in owner.parameters()), None)
    m = None
    for g in mods:
        gm = g.to_matrix(**kwargs)
        if dev is None:
            dev = gm.device
        gm = gm.to(dev)
        m = gm if m is None else m @ gm
    return m


def _dense_from_2x2(m2: torch.Tensor, qubit: int, n: int) -> torch.Tensor:
    """Row-vector-convention full matrix of a one-qubit action psi -> m2 psi:  state @ result == m2 applied."""
    if n > 12:
        raise ValueError("to_matrix() is a dense O(4^n) helper; refusing n > 12")
    full = torch.eye(1, dtype=m2.dtype, device=m2.device)
    eye = torch.eye(2, dtype=m2.dtype, device=m2.device)
    for i in range(n):
        full = torch.kron(full, m2 if i == qubit else eye)
    return full.transpose(-1, -2).contiguous()


# ---------------------------------------------------------------------------------------------------------
class UnbuiltParametrizedOperator(UnbuiltOperator):
    """RX/RY/RZ specification (reference operators.py:150-212)."""

    def __init__(self, qubit: int, theta: typing.Union[float, torch.Tensor, None] = None,
                 name: typing.Union[str, None] = None, **kwargs):
        assert isinstance(qubit, int), "qubit must be an integer"
        assert qubit >= 0, "qubit must be >= 0"
        if isinstance(theta, (float, int)):
            theta = torch.tensor(theta, requires_grad=True, dtype=torch.float)
        remapping = kwargs.get("remapping", config.DEFAULT_MAPPING)
        if remapping is None:
            remapping = remap.none
        self.qubit = qubit
        self.name = name
        self.named = name is not None
        self.theta = theta
        self.remapping = remapping

    def __str__(self) -> str:
        base = f"{self.__class__.__name__}{self.qubit}"
        if self.named:
            return f"{base} ({self.name})"
        if self.theta is None:
            return base
        return f"{base} ({self.remapping(self.theta).item():.2f})"

    def to_qasm(self) -> QasmRepresentation:
        gate = self.__class__.__name__.lower()
        if self.named:
            return QasmRepresentation(gate_str=gate, qubit=self.qubit, qasm3_inputs=self.name)
        if self.theta is not None:
            return QasmRepresentation(gate_str=gate, qubit=self.qubit, gate_value=self.remapping(self.theta).item())
        raise errors.UnbuiltGateError(
            "This gate has no parameter. Set parameter or build or set name before converting to OpenQASM2.")

    def build(self, num_qubits, **kwargs) -> "BuiltParametrizedOperator":
        return BUILT_CLASS_RELATION[self.__class__](
            qubit=self.qubit, initialtheta=self.theta, name=self.name, remapping=self.remapping, num_qubits=num_qubits)


class BuiltParametrizedOperator(BuiltOperator):
    """Built rotation: owns ``theta`` (float32 Parameter, shape (1,) when random, the given tensor's shape
    otherwise -- reference operators.py:226-230, quirk Q13).  A named gate reads ``kwargs[name]`` un-remapped and
    keeps an unused theta (quirk Q6, operators.py:266-271)."""

    engine_opcode: int = 0

    def __init__(self, qubit: int, remapping: typing.Callable, num_qubits: int,
                 initialtheta: typing.Union[torch.Tensor, None] = None, name: typing.Union[str, None] = None):
        super().__init__()
        self.qubit = qubit
        if initialtheta is None:
            initialtheta = torch.rand(1)
        self.name = name
        self.named = name is not None
        self.theta = torch.nn.Parameter(initialtheta, requires_grad=True)
        self.remapping = remapping
        self.num_qubits = num_qubits
        self.unbuilt_class = BUILT_CLASS_RELATION.T[self.__class__]

    def __str__(self) -> str:
        base = f"{self.unbuilt_class.__name__}_{self.qubit}"
        if self.named:
            return f"{base} ({self.name})"
        return f"{base} ({self.remapping(self.theta).item():.2f})"

    def to_qasm(self) -> QasmRepresentation:
        gate = self.unbuilt_class.__name__.lower()
        if self.named:
            return QasmRepresentation(gate_str=gate, qubit=self.qubit, qasm3_inputs=self.name)
        return QasmRepresentation(gate_str=gate, qubit=self.qubit, gate_value=self.remapping(self.theta).item())

    def gate_angle(self, **kwargs) -> torch.Tensor:
        """The angle fed to the engine: kwargs[name] for named gates, remapping(theta) otherwise."""
        if self.named:
            return kwargs[self.name]
        return self.remapping(self.theta)

    def matrix2(self, angle: torch.Tensor) -> torch.Tensor:
        """(…, 2, 2) complex, column-vector convention (reference operators.py:368-395 with t = angle / 2)."""
        t = angle / 2
        c, s = torch.cos(t), torch.sin(t)
        z = torch.zeros_like(c)
        if self.engine_opcode == 1:
            re = torch.stack([torch.stack([c, z], -1), torch.stack([z, c], -1)], -2)
            im = torch.stack([torch.stack([z, -s], -1), torch.stack([-s, z], -1)], -2)
        elif self.engine_opcode == 2:
            re = torch.stack([torch.stack([c, -s], -1), torch.stack([s, c], -1)], -2)
            im = torch.zeros_like(re)
        else:
            re = torch.stack([torch.stack([c, z], -1), torch.stack([z, c], -1)], -2)
            im = torch.stack([torch.stack([-s, z], -1), torch.stack([z, s], -1)], -2)
        return torch.complex(re, im)

    def to_matrix(self, **kwargs) -> torch.Tensor:
        ang = self.gate_angle(**kwargs)
        m2 = self.matrix2(ang.reshape(()) if ang.numel() == 1 else ang)
        if m2.dim() == 2:
            return _dense_from_2x2(m2, self.qubit, self.num_qubits)
        return torch.stack([_dense_from_2x2(m, self.qubit, self.num_qubits) for m in m2])


class RX(UnbuiltParametrizedOperator):
    """Rotation around x: exp(-i theta X / 2).  Same arguments as the reference's RX (operators.py:308-325)."""


class RY(UnbuiltParametrizedOperator):
    """Rotation around y: exp(-i theta Y / 2) (reference operators.py:328-345)."""


class RZ(UnbuiltParametrizedOperator):
    """Rotation around z: exp(-i theta Z / 2) (reference operators.py:348-365)."""


class BuiltRX(BuiltParametrizedOperator):
    engine_opcode = 1


class BuiltRY(BuiltParametrizedOperator):
    engine_opcode = 2


class BuiltRZ(BuiltParametrizedOperator):
    engine_opcode = 3


# ---------------------------------------------------------------------------------------------------------
class U(UnbuiltOperator):
    """User-defined one-qubit gate (reference operators.py:72-84)."""

    def __init__(self, qubit: int, matrix: torch.Tensor):
        self.qubit = qubit
        self.matrix = matrix

    def __str__(self) -> str:
        return f"U_{self.qubit}"

    def to_qasm(self) -> QasmRepresentation:
        return QasmRepresentation(gate_str="U", qubit=self.qubit)

    def build(self, num_qubits, **kwargs) -> "BuiltU":
        return BuiltU(qubit=self.qubit, matrix=self.matrix, num_qubits=num_qubits)


class BuiltU(BuiltOperator):
    """Built custom gate.  Parity quirk Q2: the reference kron-expands ``matrix`` WITHOUT transposing and
    right-multiplies (operators.py:100-104, 125-126), so the gate acts as ``matrix^T psi``; reproduced here by
    handing the engine the transpose."""

    def __init__(self, qubit: int, matrix: torch.Tensor, num_qubits: int, self_description: str = "U"):
        super().__init__()
        self.qubit = qubit
        self.num_qubits = num_qubits
        self.description = self_description
        self.original_matrix = matrix
        m = torch.as_tensor(matrix)
        if m.shape != (2, 2):
            raise ValueError("U expects a 2x2 matrix")
        self.engine_matrix = m.detach().to(torch.complex128).transpose(0, 1).contiguous()  # applied as M psi
        # The engine's backward is the adjoint-state method: it un-computes psi with U^+, which is U^-1 only for a unitary.  The
        # reference accepts ANY 2x2 (its tape differentiates through a dense matmul, operators.py:125-126) and also back-propagates
        # into a matrix that requires grad; those two cases run as a torch module between engine segments instead.
        dev = (self.engine_matrix @ self.engine_matrix.conj().transpose(0, 1) - torch.eye(2, dtype=torch.complex128)).abs().max()
        self.engine_unitary = bool(dev < 1e-6) and not m.requires_grad
        if not self.engine_unitary and not m.requires_grad:
            warnings.warn(f"{self.description}_{qubit}: the matrix is not unitary (|U U^+ - 1| = {float(dev):.1e}); it is applied by a torch "
                          "fallback between engine segments (the adjoint-state backward needs unitary gates)")

    def __str__(self) -> str:
        return f"{self.description}_{self.qubit}"

    def to_qasm(self) -> QasmRepresentation:
        o = self.original_matrix
        u = f"{o[0, 0]:.2f}, {o[0, 1]:.2f}, {o[1, 0]:.2f}, {o[1, 1]:.2f}"
        return QasmRepresentation(gate_str=f"gate {self.description}_{self.qubit} q[{self.qubit}] {{{{ U({u}) }}}}", qubit=self.qubit)

    @property
    def matrix(self) -> torch.Tensor:
        return self.to_matrix()

    def to_matrix(self, **kwargs) -> torch.Tensor:
        return _dense_from_2x2(self.engine_matrix.to(torch.complex64), self.qubit, self.num_qubits)

    def forward(self, state: torch.Tensor, **kwargs) -> torch.Tensor:
        if self.engine_unitary:
            return super().forward(state, **kwargs)
        # non-unitary or trainable matrix: O(2^n) torch path, differentiable by torch's tape like the reference's
        # `state @ kron(.., matrix, ..)` (acts as matrix^T psi, quirk Q2)
        m = torch.as_tensor(self.original_matrix).to(device=state.device)
        m = m.to(state.dtype if state.is_complex() else torch.complex64).transpose(0, 1)
        st = state if state.is_complex() else state.to(m.dtype)
        lead = st.shape[:-1]
        v = st.reshape(-1, 2**self.qubit, 2, 2 ** (self.num_qubits - self.qubit - 1))
        return torch.einsum("ij,bhjl->bhil", m, v).reshape(*lead, -1)


CustomGate = BuiltU


# ---------------------------------------------------------------------------------------------------------
class _TwoQubitUnbuilt(UnbuiltOperator):
    _label = ""
    _qasm = ""

    def __init__(self, control: int, target: int):
        assert control != target, "Control and target must be different"
        self.c = control
        self.t = target

    def __str__(self) -> str:
        return f"{self._label} {self.c}|{self.t}"

    def to_qasm(self) -> QasmRepresentation:
        return QasmRepresentation(gate_str=f"{self._qasm} q[{self.c}], q[{self.t}]")

    def build(self, num_qubits, **kwargs):
        return BUILT_CLASS_RELATION[self.__class__](control=self.c, target=self.t, num_qubits=num_qubits)


class CNOT(_TwoQubitUnbuilt):
    """CNOT(control, target) (reference operators.py:398-413)."""

    _label, _qasm = "CNOT", "cx"


class CZ(_TwoQubitUnbuilt):
    """CZ(control, target) (reference operators.py:518-533)."""

    _label, _qasm = "CZ", "cz"


class _TwoQubitBuilt(BuiltOperator):
    engine_opcode = 0
    _unbuilt: type = CNOT

    def __init__(self, control: int, target: int, num_qubits: int):
        super().__init__()
        self.c = control
        self.t = target
        self.num_qubits = num_qubits

    def __str__(self) -> str:
        return self._unbuilt(self.c, self.t).__str__()

    def to_qasm(self) -> QasmRepresentation:
        return self._unbuilt(self.c, self.t).to_qasm()

    def _index_map(self):
        n = self.num_qubits
        if n > 12:
            raise ValueError("to_matrix() is a dense O(4^n) helper; refusing n > 12")
        idx = torch.arange(2**n)
        return idx, n - self.c - 1, n - self.t - 1


class BuiltCNOT(_TwoQubitBuilt):
    """Bit position of qubit q is n-q-1 (reference operators.py:546-555)."""

    engine_opcode = 5
    _unbuilt = CNOT

    def to_matrix(self, **kwargs) -> torch.Tensor:
        idx, c2, t2 = self._index_map()
        dst = torch.where((idx >> c2) & 1 == 1, idx ^ (1 << t2), idx)
        M = torch.zeros(len(idx), len(idx), dtype=torch.cfloat)
        M[idx, dst] = 1
        return M


class BuiltCZ(_TwoQubitBuilt):
    """-1 on indices with both bits set (reference operators.py:581-588)."""

    engine_opcode = 6
    _unbuilt = CZ

    def to_matrix(self, **kwargs) -> torch.Tensor:
        idx, c2, t2 = self._index_map()
        diag = torch.ones(len(idx), dtype=torch.cfloat)
        diag[(((idx >> c2) & 1) & ((idx >> t2) & 1)) == 1] = -1
        return torch.diag(diag)


class SWAP(UnbuiltOperator):
    """SWAP(a, b) (reference operators.py:686-698)."""

    def __init__(self, a: int, b: int):
        self.a = a
        self.b = b

    def __str__(self) -> str:
        return f"Swap {self.a}|{self.b}"

    def to_qasm(self) -> QasmRepresentation:
        return QasmRepresentation(gate_str=f"swap q[{self.a}], q[{self.b}]")

    def build(self, num_qubits, **kwargs) -> "BuiltSWAP":
        return BuiltSWAP(a=self.a, b=self.b, num_qubits=num_qubits)


class BuiltSWAP(BuiltOperator):
    """reference operators.py:653-683."""

    engine_opcode = 7

    def __init__(self, a: int, b: int, num_qubits: int):
        super().__init__()
        self.a = a
        self.b = b
        self.num_qubits = num_qubits

    def __str__(self) -> str:
        return f"SWAP {self.a}|{self.b}"

    def to_qasm(self) -> QasmRepresentation:
        return QasmRepresentation(gate_str=f"swap q[{self.a}], q[{self.b}]")

    def to_matrix(self, **kwargs) -> torch.Tensor:
        n = self.num_qubits
        if n > 12:
            raise ValueError("to_matrix() is a dense O(4^n) helper; refusing n > 12")
        idx = torch.arange(2**n)
        a2, b2 = n - self.a - 1, n - self.b - 1
        diff = ((idx >> a2) & 1) != ((idx >> b2) & 1)
        dst = torch.where(diff, idx ^ ((1 << a2) | (1 << b2)), idx)
        M = torch.zeros(len(idx), len(idx), dtype=torch.cfloat)
        M[idx, dst] = 1
        return M


# ---------------------------------------------------------------------------------------------------------
# SURVEY.md 8f rank 3 ("next"): operators outside the accelerated gate set, kept for API completeness.
class Reset(UnbuiltOperator):
    """Reset a qubit to |0> while preserving the norm of the state (reference operators.py:603-650).  Non-unitary, so
    it is not an engine op: inside a Circuit it runs in torch between two engine segments."""

    def __init__(self, qubit: int):
        self.qubit = qubit

    def __str__(self) -> str:
        return f"Reset {self.qubit}"

    def to_qasm(self) -> QasmRepresentation:
        return QasmRepresentation(gate_str="reset", qubit=self.qubit)

    def build(self, num_qubits, **kwargs) -> "BuiltReset":
        return BuiltReset(qubit=self.qubit, num_qubits=num_qubits)


class BuiltReset(torch.nn.Module):
    """Projection on qubit = 0 with a per-pair rescale that preserves the norm (reference operators.py:612-623)."""

    named = False

    def __init__(self, qubit: int, num_qubits: int):
        super().__init__()
        self.qubit = qubit
        self.num_qubits = num_qubits

    def __repr__(self) -> str:  # "Reset 3" inside print(circuit), like the reference's built operators
        return str(self)

    def forward(self, state: torch.Tensor) -> torch.Tensor:
        unbatched = state.dim() == 1
        if unbatched:
            state = state.unsqueeze(0)
        B = state.shape[0]
        v = state.reshape(B, 2**self.qubit, 2, 2 ** (self.num_qubits - self.qubit - 1))
        keep = v[:, :, 0, :]
        # every amplitude pair (a0, a1) of the qubit becomes (a0 |pair| / |a0|, 0): the reference rescales per pair
        # (rows of its "(batch rest) (sub)" matrix view, operators.py:616-619), not per state
        # (torch.linalg.norm like the reference, not sqrt / abs: its subgradient at an all-zero pair is 0, theirs NaN)
        scale = torch.linalg.norm(v, dim=2) / (torch.linalg.norm(v[:, :, :1, :], dim=2) + 1e-7)
        out = torch.zeros_like(v, dtype=torch.cfloat if state.dtype != torch.complex128 else torch.complex128)
        out[:, :, 0, :] = keep * scale
        out = out.reshape(B, -1)
        return out.squeeze(0) if unbatched else out

    def __str__(self) -> str:
        return f"Reset {self.qubit}"

    def to_qasm(self) -> QasmRepresentation:
        return QasmRepresentation(gate_str="reset", qubit=self.qubit)

    def to_matrix(self, **kwargs):
        raise ValueError("Reset gate does not have a matrix representation.")


class Invert(UnbuiltOperator):
    """Inverse of an engine gate (reference operators.py:416-456).  Rotations are lowered with the negated angle, U with
    the conjugate transpose; CNOT / CZ / SWAP are their own inverses."""

    def __init__(self, target: Operator):
        self.t = target

    def __str__(self) -> str:
        return f"{self.t}^-1"

    def to_qasm(self) -> QasmRepresentation:
        return QasmRepresentation(gate_str=f"{self.t}^-1")

    def build(self, num_qubits, **kwargs):
        target = self.t.build(num_qubits) if hasattr(self.t, "build") else self.t
        return BuiltInvert(target, num_qubits)


def _neg_remap(fn):
    def neg(x):
        return -fn(x)

    return neg


class BuiltInvert(BuiltOperator):
    def __init__(self, target: Operator, num_qubits: int):
        super().__init__()
        if hasattr(target, "build"):
            target = target.build(num_qubits)
        self.target = target  # registered under the reference's name: `...target.theta`
        self.num_qubits = num_qubits
        if isinstance(target, BuiltParametrizedOperator):
            if target.named:
                raise NotImplementedError("Invert of a named gate is not supported")
        elif not isinstance(target, (BuiltU, BuiltCNOT, BuiltCZ, BuiltSWAP)):
            raise NotImplementedError(f"Invert({type(target).__name__}) is not supported by the engine")

    def engine_lower(self, slot0: int, mat0: int):
        """-> (gate-program rows, shared-angle sources [(module, attr, n_slots, remapping)], fixed matrices)."""
        t = self.target
        if isinstance(t, BuiltParametrizedOperator):  # R(theta)^-1 = R(-theta)
            return [(t.engine_opcode, t.qubit, -1, slot0)], [(t, "theta", 1, _neg_remap(t.remapping))], []
        if isinstance(t, BuiltU):  # the reference inverts the dense matrix (operators.py:416-456: torch.linalg.inv)
            if not t.engine_unitary:
                raise NotImplementedError("Invert(U) of a non-unitary / trainable matrix is not supported by the engine")
            return [(4, t.qubit, -1, mat0)], [], [torch.linalg.inv(t.engine_matrix).contiguous()]
        if isinstance(t, BuiltSWAP):
            return [(t.engine_opcode, t.a, t.b, 0)], [], []
        return [(t.engine_opcode, t.c, t.t, 0)], [], []

    def __str__(self) -> str:
        return f"{self.target}^-1"

    def to_qasm(self) -> QasmRepresentation:
        return QasmRepresentation(gate_str=f"{self.target}^-1")

    def to_matrix(self, **kwargs):
        return torch.linalg.inv(self.target.to_matrix(**kwargs))



# ---------------------------------------------------------------------------------------------------------
def _half_remap(fn, sign):
    def half(x):
        return (0.5 * sign) * fn(x)

    return half


def _rz2(a):
    return torch.tensor([[cmath.exp(-0.5j * a), 0], [0, cmath.exp(0.5j * a)]], dtype=torch.complex128)


def _ry2(a):
    c, s = math.cos(a / 2), math.sin(a / 2)
    return torch.tensor([[c, -s], [s, c]], dtype=torch.complex128)


def _zyz(m: torch.Tensor):
    """m (2x2 unitary, complex128) = exp(i alpha) RZ(beta) RY(gamma) RZ(delta) -> (alpha, beta, gamma, delta)."""
    det = complex(m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0])
    alpha = cmath.phase(det) / 2
    v = m * cmath.exp(-1j * alpha)
    v00, v10, v11 = complex(v[0, 0]), complex(v[1, 0]), complex(v[1, 1])
    gamma = 2 * math.atan2(abs(v10), abs(v00))
    if abs(v10) < 1e-12:
        beta = delta = cmath.phase(v11)
    elif abs(v00) < 1e-12:
        beta = cmath.phase(v10)
        delta = -beta
    else:
        plus, minus = 2 * cmath.phase(v11), 2 * cmath.phase(v10)  # beta + delta, beta - delta
        beta, delta = (plus + minus) / 2, (plus - minus) / 2
    return alpha, beta, gamma, delta


_H2 = torch.tensor([[1, 1], [1, -1]], dtype=torch.complex128) / math.sqrt(2)
_T2 = torch.tensor([[1, 0], [0, cmath.exp(0.25j * math.pi)]], dtype=torch.complex128)


class Controlled(UnbuiltOperator):
    """Apply ``target`` where qubit ``control`` is 1 (reference operators.py:459-475)."""

    def __init__(self, control: int, target: Operator):
        self.c = control
        self.t = target

    def __str__(self) -> str:
        return f"Controlled {self.c}|{self.t}"

    def to_qasm(self) -> QasmRepresentation:
        return QasmRepresentation(gate_str=f"controlled q[{self.c}], {self.t}")

    def build(self, num_qubits, **kwargs) -> "BuiltControlled":
        return BuiltControlled(control=self.c, target=self.t, num_qubits=num_qubits)


class BuiltControlled(BuiltOperator):
    """The reference masks the state to control = 1, multiplies by the target's dense matrix and selects per amplitude
    (operators.py:478-515).  For a target that does not act on the control qubit that is the ordinary controlled gate;
    here it is lowered to engine ops, so it runs inside the fused sweeps and its angle stays differentiable:

    * controlled RY / RZ(theta) = R(theta/2) . CNOT . R(-theta/2) . CNOT, controlled RX the same with CZ
      (X R(a) X = R(-a) for RY / RZ, Z RX(a) Z = RX(-a)); a named target takes the halves of the per-sample input;
    * controlled U: ZYZ decomposition  U = e^{ia} RZ(b) RY(g) RZ(d)  ->  A . CNOT . B . CNOT . C + a phase on the control;
    * controlled CNOT / CZ / SWAP: Toffoli from 6 CNOTs and T gates, conjugated as needed.

    The submodule is registered as ``t`` like the reference's (Parameter name ``...t.theta``)."""

    def __init__(self, control: int, target: Operator, num_qubits: int):
        super().__init__()
        self.c = control
        if hasattr(target, "build"):
            target = target.build(num_qubits)
        self.t = target
        self.named = target.named
        self.num_qubits = num_qubits
        t = target
        if isinstance(t, (BuiltParametrizedOperator, BuiltU)):
            touched = [t.qubit]
        elif isinstance(t, (BuiltCNOT, BuiltCZ)):
            touched = [t.c, t.t]
        elif isinstance(t, BuiltSWAP):
            touched = [t.a, t.b]
        else:
            raise NotImplementedError(f"Controlled({type(t).__name__}) is not supported by the engine")
        if control in touched:
            raise NotImplementedError("Controlled: the target acts on the control qubit (the reference's result is not unitary)")
        if isinstance(t, BuiltU):
            m = t.engine_matrix
            if float((m.conj().T @ m - torch.eye(2, dtype=m.dtype)).abs().max()) > 1e-5:
                raise NotImplementedError("Controlled(U): the matrix must be unitary")

    # -- lowering ------------------------------------------------------------------------------------------------
    @staticmethod
    def _toffoli(c1: int, c2: int, t: int, mat0: int):
        """CCX(c1, c2 -> t) as engine rows; fixed matrices [H, T, T^+] start at index mat0."""
        h, tg, td = mat0, mat0 + 1, mat0 + 2
        u = lambda q, k: (4, q, -1, k)
        cx = lambda a, b: (5, a, b, 0)
        return [u(t, h), cx(c2, t), u(t, td), cx(c1, t), u(t, tg), cx(c2, t), u(t, td), cx(c1, t), u(c2, tg), u(t, tg), u(t, h),
                cx(c1, c2), u(c1, tg), u(c2, td), cx(c1, c2)]

    def engine_lower_into(self, seg) -> None:
        """Append this gate's engine rows to a lowering segment (qcircuit._Segment)."""
        t, c = self.t, self.c
        if isinstance(t, BuiltParametrizedOperator):
            q = t.qubit
            link = (6, c, q, 0) if t.engine_opcode == 1 else (5, c, q, 0)  # CZ for RX, CNOT for RY / RZ
            if t.named:
                plus = seg.batch_col(t.name, -1, 0.5)
                minus = seg.batch_col(t.name, -1, -0.5)
                flag = 0x100
            else:
                plus, minus, flag = seg.n_slots, seg.n_slots + 1, 0
                seg.weight_srcs.append((t, "theta", 1, _half_remap(t.remapping, 1.0)))
                seg.weight_srcs.append((t, "theta", 1, _half_remap(t.remapping, -1.0)))
                seg.n_slots += 2
            seg.rows += [(t.engine_opcode | flag, q, -1, plus), link, (t.engine_opcode | flag, q, -1, minus), link]
        elif isinstance(t, BuiltU):
            q = t.qubit
            al, be, ga, de = _zyz(t.engine_matrix)
            m0 = len(seg.mats)
            seg.mats += [_rz2((de - be) / 2), _ry2(-ga / 2) @ _rz2(-(de + be) / 2), _rz2(be) @ _ry2(ga / 2),
                         torch.tensor([[1, 0], [0, cmath.exp(1j * al)]], dtype=torch.complex128)]
            seg.rows += [(4, q, -1, m0), (5, c, q, 0), (4, q, -1, m0 + 1), (5, c, q, 0), (4, q, -1, m0 + 2), (4, c, -1, m0 + 3)]
        else:
            m0 = len(seg.mats)
            seg.mats += [_H2, _T2, _T2.conj()]
            if isinstance(t, BuiltCNOT):
                seg.rows += self._toffoli(c, t.c, t.t, m0)
            elif isinstance(t, BuiltCZ):  # CCZ = H_t CCX H_t
                seg.rows += [(4, t.t, -1, m0)] + self._toffoli(c, t.c, t.t, m0) + [(4, t.t, -1, m0)]
            else:  # Fredkin: CX(b, a) CCX(c, a -> b) CX(b, a)
                seg.rows += [(5, t.b, t.a, 0)] + self._toffoli(c, t.a, t.b, m0) + [(5, t.b, t.a, 0)]

    def __str__(self) -> str:
        return f"Controlled {self.c}|{self.t}"

    def to_qasm(self) -> QasmRepresentation:
        return QasmRepresentation(gate_str=f"controlled q[{self.c}], {self.t}")

    def to_matrix(self, **kwargs) -> torch.Tensor:
        """Dense controlled matrix, row-vector convention (small n only).  The reference returns the bare target's
        matrix here (operators.py:514-515, quirk Q11): its own forward does not agree with that, this one does."""
        tm = self.t.to_matrix(**kwargs)
        N = 2**self.num_qubits
        on = ((torch.arange(N, device=tm.device) >> (self.num_qubits - self.c - 1)) & 1).bool()
        eye = torch.eye(N, dtype=tm.dtype, device=tm.device)
        return torch.where(on[:, None], tm, eye) if tm.dim() == 2 else torch.where(on[None, :, None], tm, eye[None])


class rdict(dict):
    """dict with an inverse view (reference operators.py:701-706)."""

    @property
    def T(self):
        return {v: k for k, v in self.items()}


BUILT_CLASS_RELATION = rdict({
    UnbuiltOperator: BuiltOperator,
    UnbuiltParametrizedOperator: BuiltParametrizedOperator,
    RX: BuiltRX,
    RY: BuiltRY,
    RZ: BuiltRZ,
    CNOT: BuiltCNOT,
    U: BuiltU,
    SWAP: BuiltSWAP,
    CZ: BuiltCZ,
    Reset: BuiltReset,
    Invert: BuiltInvert,
    Controlled: BuiltControlled,
})
