"""Measurements: same surface as reference src/qandle/measurements.py.

Inside a Circuit a trailing measurement is fused into the engine call (MeasureProbability -> one-pass
reduction kernel instead of the reference's n-fold repeat of the state, measurements.py:120-123).  Called
directly on a state, the same engine path runs with an empty gate program.
"""
import random
import warnings

import torch

from . import operators as op
from .errors import UnbuiltGateError

__all__ = ["BuiltMeasurement", "UnbuiltMeasurement", "MeasureAllAbsolute", "MeasureJointProbability", "MeasureState",
           "MeasureProbability", "MeasureProbabilityBuilt", "MeasureExpectation", "MeasureExpectationBuilt"]


class BuiltMeasurement(op.BuiltOperator):
    """reference measurements.py:12-60."""

    engine_measure = 0
    num_qubits = None

    def __str__(self) -> str:
        return "M"

    def forward(self, state: torch.Tensor) -> torch.Tensor:
        from . import qcircuit

        n = self.num_qubits
        if n is None:
            n = int(state.shape[-1]).bit_length() - 1
        return qcircuit.run_modules(self, [self], n, state, {})

    def to_qasm(self):
        """One ``measure q[w] -> c[w]`` per qubit, each with an OpenQASM 3 output name (reference measurements.py:16-22)."""
        if self.num_qubits is None:  # size-agnostic measurements: the reference raises AttributeError here
            return []
        i = hash(random.random())
        return [op.QasmRepresentation(gate_str=f"measure q[{w}] -> c[{w}]", qasm3_outputs=f"measured_{i}_{w}")
                for w in range(self.num_qubits)]

    def decompose(self):
        warnings.warn("Decomposed measurements a no-op. Consider acting on the resulting statevector directly", RuntimeWarning)
        return []

    def to_matrix(self):
        return torch.eye(2**self.num_qubits)


class UnbuiltMeasurement(op.UnbuiltOperator):
    """reference measurements.py:62-70."""

    def __str__(self) -> str:
        return "M"

    def to_qasm(self):
        raise UnbuiltGateError(f"Unbuilt {self.__class__.__name__} cannot be converted to qasm.")

    def build(self, num_qubits: int, **kwargs) -> BuiltMeasurement:
        return BUILT_CLASS_RELATION[self.__class__](num_qubits=num_qubits, **kwargs)


class MeasureAllAbsolute(BuiltMeasurement):
    """|psi_i|^2 for every basis state (reference measurements.py:73-82)."""

    engine_measure = 2


MeasureJointProbability = MeasureAllAbsolute


class MeasureState(BuiltMeasurement):
    """The state vector itself (reference measurements.py:85-91)."""

    engine_measure = 0


class MeasureProbability(UnbuiltMeasurement):
    r"""P(qubit = |0>) for every qubit; ``1 - p`` is P(|1>) (reference measurements.py:94-98)."""


class MeasureProbabilityBuilt(BuiltMeasurement):
    """Output (B, n) with the reference's ``.squeeze()`` shape quirks: B = 1 and n = 1 are dropped
    (reference measurements.py:101-123, quirk Q5)."""

    engine_measure = 1

    def __init__(self, num_qubits: int):
        super().__init__()
        self.num_qubits = num_qubits
        self.qubits = list(range(num_qubits))


def _z_from_p0(p: torch.Tensor) -> torch.Tensor:
    return 2 * p - 1


class MeasureExpectation(UnbuiltMeasurement):
    r"""Expectation value :math:`\langle P_q \rangle` of one Pauli operator (``"z"``, ``"x"`` or ``"y"``) for every qubit.

    No reference counterpart (BASELINE north_star (c) names expectation-value reductions).  It reuses
    the engine's one-pass probability reduction: :math:`\langle Z_q \rangle = 2 P(q = 0) - 1` after a basis change on
    every qubit (H for X, H S^+ for Y) that is fused into the circuit's last sweep.  Output shape rules are those of
    ``MeasureProbability``."""

    def __init__(self, pauli: str = "z"):
        pauli = pauli.lower()
        if pauli not in ("x", "y", "z"):
            raise ValueError(f"pauli must be 'x', 'y' or 'z', got {pauli!r}")
        self.pauli = pauli

    def build(self, num_qubits: int, **kwargs) -> "MeasureExpectationBuilt":
        return MeasureExpectationBuilt(num_qubits=num_qubits, pauli=self.pauli)


class MeasureExpectationBuilt(BuiltMeasurement):
    engine_measure = 1

    def __init__(self, num_qubits: int, pauli: str = "z"):
        super().__init__()
        self.num_qubits = num_qubits
        self.pauli = pauli
        self.qubits = list(range(num_qubits))

    def engine_lower_into(self, seg) -> None:
        if self.pauli != "z":
            r = 2**-0.5
            h = torch.tensor([[r, r], [r, -r]], dtype=torch.complex128)
            if self.pauli == "y":  # H S^+ maps the Y eigenbasis to the computational basis
                h = h @ torch.tensor([[1, 0], [0, -1j]], dtype=torch.complex128)
            k = len(seg.mats)
            seg.mats.append(h)
            seg.rows += [(4, qb, -1, k) for qb in range(self.num_qubits)]
        seg.measure = self.engine_measure
        seg.post = _z_from_p0


BUILT_CLASS_RELATION = {MeasureProbability: MeasureProbabilityBuilt, MeasureExpectation: MeasureExpectationBuilt}
