"""Mirror of reference src/qandle/errors.py:1-6."""


class UnbuiltGateError(ValueError):
    """Raised when a gate is used before it is built (build it with ``gate.build(num_qubits)`` or put it in a Circuit)."""
