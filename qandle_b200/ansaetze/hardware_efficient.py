"""TwoLocal, SpecialUnitary and StronglyEntanglingLayerBudget: same surface, Parameter names and
FORWARD semantics as reference src/qandle/ansaetze/{twolocal,specialunitary,stronglyentangling_budget}.py.

The decision on quirk Q8: parity is against what the reference's ``forward`` computes, not against its ``decompose()``.
``TwoLocalBuilt.forward`` / ``SpecialUnitaryBuilt.forward`` left-multiply the state by the product C_0 C_1 ... C_{k-2} of
the linear CNOT chain (twolocal.py:83-92, 99-103; specialunitary.py:93-108), i.e. they apply the chain in REVERSED order
(CNOT(q[k-2], q[k-1]) first); that is what the engine program does.  The reference's form only accepts an unbatched
state there; here batched states work as well (same circuit per sample).  ``TwoLocalBuilt`` builds its RY gates on
``range(num_qubits)`` with the default remapping whatever ``qubits`` / ``remapping`` say (twolocal.py:78-85): reproduced.
Each ansatz is one run of engine rows inside the circuit's gate program -- no dense CNOT-chain matrix.
"""
import copy
import typing

import torch

from .. import config, remap, utils
from .. import operators as op

__all__ = ["TwoLocal", "TwoLocalBuilt", "SpecialUnitary", "SpecialUnitaryBuilt", "SU", "StronglyEntanglingLayerBudget"]


def _reversed_chain(qubits: typing.List[int], num_qubits: int) -> typing.List[op.BuiltCNOT]:
    return [op.CNOT(qubits[ci], qubits[ci + 1]).build(num_qubits) for ci in reversed(range(len(qubits) - 1))]


class _BuiltAnsatz(op.BuiltOperator):
    def engine_children(self) -> typing.List[torch.nn.Module]:
        """Built gates in the order the reference's forward applies them (qcircuit._flatten)."""
        raise NotImplementedError

    def to_matrix(self, **kwargs):
        return op.chain_matrices(self, self.engine_children(), **kwargs)


class TwoLocal(op.UnbuiltOperator):
    """RY layer + linear CNOT chain, ``depth`` times (reference twolocal.py:13-60)."""

    def __init__(self, qubits: typing.Union[typing.List[int], None] = None, depth: int = 1,
                 remapping: typing.Union[typing.Callable, None] = config.DEFAULT_MAPPING, q_params=None):
        self.depth = depth
        self.remapping = remapping
        self.q_params = q_params
        self.qubits = qubits

    def build(self, *args, **kwargs) -> "TwoLocalBuilt":
        return TwoLocalBuilt(num_qubits=kwargs["num_qubits"], qubits=self.qubits, depth=self.depth, remapping=self.remapping,
                             q_params=self.q_params)

    def __str__(self) -> str:
        return "TL"

    def to_qasm(self):
        return [g.to_qasm() for g in self.decompose()]

    def decompose(self) -> typing.List[op.Operator]:
        if self.qubits is None:
            raise ValueError("It is not specified on which qubits to apply the ansatz. Build the circuit to specify or pass them explicitly as a list")
        res = []
        qp = self.q_params if self.q_params is not None else torch.rand(self.depth, len(self.qubits))
        for d in range(self.depth):
            for wi in range(len(self.qubits)):
                res.append(op.RY(qubit=self.qubits[wi], theta=qp[d, wi]))
            for wi in range(len(self.qubits) - 1):
                res.append(op.CNOT(self.qubits[wi], self.qubits[wi + 1]))
        return res


class TwoLocalBuilt(_BuiltAnsatz):
    def __init__(self, num_qubits: int, qubits: typing.Union[typing.List[int], None], depth: int = 1,
                 remapping: typing.Union[typing.Callable, None] = config.DEFAULT_MAPPING, q_params=None):
        super().__init__()
        self.num_qubits = num_qubits
        self.depth = depth
        self.remapping = remapping
        self.qubits = qubits or list(range(num_qubits))
        if q_params is None or q_params.shape != (depth, len(self.qubits)):
            q_params = torch.rand(depth, num_qubits)
        layers = [[op.RY(qubit=w, theta=q_params[d, w]).build(num_qubits) for w in range(num_qubits)] for d in range(depth)]
        self.mods = torch.nn.ModuleList([torch.nn.Sequential(*lay) for lay in layers])
        self._chain = _reversed_chain(self.qubits, num_qubits)  # plain list: CNOTs own no state

    def engine_children(self):
        out = []
        for d in range(self.depth):
            out.extend(self.mods[d])
            out.extend(self._chain)
        return out

    def __str__(self) -> str:
        return "TL"

    def to_qasm(self):
        outp = []
        for d in range(self.depth):
            outp.extend([g.to_qasm() for g in self.mods[d]])
            for ci in range(len(self.qubits) - 1):
                outp.append(op.CNOT(self.qubits[ci], self.qubits[ci + 1]).to_qasm())
        return outp

    def decompose(self) -> typing.List[op.Operator]:
        outp = []
        for mod in self.mods:
            for seq in mod:
                outp.append(copy.deepcopy(seq))
            for ci in range(len(self.qubits) - 1):
                outp.append(op.CNOT(self.qubits[ci], self.qubits[ci + 1]))
        return outp


class SpecialUnitary(op.UnbuiltOperator):
    """Hardware-efficient SU(2) 2-local circuit (reference specialunitary.py:11-57)."""

    def __init__(self, qubits: typing.Union[typing.List[int], None] = None, reps: int = 1, rotations=("ry",)):
        self.qubits = qubits
        self.reps = reps
        self.rotations = [utils.parse_rot(r) for r in rotations]

    def build(self, *args, **kwargs) -> "SpecialUnitaryBuilt":
        return SpecialUnitaryBuilt(num_qubits=kwargs["num_qubits"], reps=self.reps, rotations=self.rotations, qubits=self.qubits)

    def __str__(self) -> str:
        return "SU"

    def to_qasm(self):
        return [g.to_qasm() for g in self.decompose()]

    def decompose(self) -> typing.List[op.Operator]:
        if self.qubits is None:
            raise ValueError("It is not specified on which qubits to apply the ansatz. Build the circuit to specify or pass them explicitly as a list")
        outp = []
        for r in self.rotations:
            for w in self.qubits:
                outp.append(r(w))
        for _ in range(self.reps):
            for wi in range(len(self.qubits) - 1):
                outp.append(op.CNOT(self.qubits[wi], self.qubits[wi + 1]))
            for r in self.rotations:
                for w in self.qubits:
                    outp.append(r(w))
        return outp


SU = SpecialUnitary


class SpecialUnitaryBuilt(_BuiltAnsatz):
    def __init__(self, num_qubits: int, qubits: typing.Union[typing.List[int], None], reps: int, rotations):
        super().__init__()
        self.num_qubits = num_qubits
        self.reps = reps
        self.rotations = rotations
        self.qubits = qubits or list(range(num_qubits))
        blocks = [[r(w).build(num_qubits=num_qubits) for w in self.qubits for r in rotations] for _ in range(reps + 1)]
        self.layers = torch.nn.ModuleList([torch.nn.Sequential(*b) for b in blocks])
        self._chain = _reversed_chain(self.qubits, num_qubits)

    def engine_children(self):
        out = list(self.layers[0])
        for layer in list(self.layers)[1:]:
            out.extend(self._chain)
            out.extend(layer)
        return out

    def __str__(self) -> str:
        return "SU"

    def to_qasm(self):
        return [g.to_qasm() for g in self.decompose()]

    def decompose(self) -> typing.List[op.Operator]:
        res = list(self.layers[0])
        for mod in list(self.layers)[1:]:
            res.extend(mod)
            for w in range(len(self.qubits) - 1):
                res.append(op.CNOT(self.qubits[w], self.qubits[w + 1]))
        return res


class StronglyEntanglingLayerBudget(_BuiltAnsatz):
    """Strongly entangling layers sized by a parameter budget instead of a depth; built directly
    (reference stronglyentangling_budget.py:15-94)."""

    def __init__(self, num_qubits_total: int, qubits: typing.Union[typing.List[int], None] = None, param_budget: int = 1,
                 control_gate=op.CNOT, control_gate_spacing=1, rotations=("rz", "ry"),
                 remapping: typing.Union[typing.Callable, None] = config.DEFAULT_MAPPING):
        super().__init__()
        assert control_gate in [op.CNOT, op.CZ], f"Control gate {control_gate} not supported"
        if qubits is None:
            qubits = list(range(num_qubits_total))
        self.qubits = qubits
        self.num_qubits = num_qubits_total
        self.rots = [utils.parse_rot(r) for r in rotations]
        if remapping is None:
            remapping = remap.none
        self.remapping = remapping
        self.param_budget = param_budget
        self.control_gate = control_gate
        self.control_gate_spacing = control_gate_spacing
        self.layers_ub = self._allocate_params()
        self.layers = torch.nn.Sequential(*[layer.build(num_qubits=num_qubits_total) for layer in self.layers_ub])

    def _allocate_params(self):
        """Rotation layers cycle through ``rotations``; a ring of control gates with a growing range goes in front of
        every ``control_gate_spacing``-th layer; stop when the budget is spent (stronglyentangling_budget.py:49-79)."""
        out, nq = [], len(self.qubits)
        params, depth, ring = 0, 0, -1
        while True:
            rot = self.rots[depth % len(self.rots)]
            if depth % self.control_gate_spacing == 0 and depth > 0:
                ring = (ring + 1) % nq
                it = ring % (nq - 1)
                for ci in range(nq):
                    out.append(self.control_gate(control=self.qubits[ci], target=self.qubits[(ci + 1 + it) % nq]))
            depth += 1
            for qb in self.qubits:
                params += 1
                if params > self.param_budget:
                    return out
                out.append(rot(qubit=qb, remapping=self.remapping))

    def engine_children(self):
        return list(self.layers)

    def __str__(self) -> str:
        return "SEL"

    def to_qasm(self):
        return [layer.to_qasm() for layer in self.layers_ub]

    def decompose(self) -> typing.List[op.UnbuiltOperator]:
        return self.layers_ub
