"""StronglyEntanglingLayer: same surface as reference src/qandle/ansaetze/stronglyentangling.py.

Per depth d: the rotation list (default rz, ry, rz) on every listed qubit with q_params[d, wi, r], then a CNOT
ring CNOT(q[c], q[(c + d % (nq-1) + 1) % nq]) (reference stronglyentangling.py:93-121).  One 0-dim Parameter per
gate under ``mods.<j>.theta`` so reference checkpoints load unchanged; the engine fuses each
qubit's rz-ry-rz run into one 2x2 on the device.
"""
import copy
import typing

import torch

from .. import config, remap, utils
from .. import operators as op

__all__ = ["StronglyEntanglingLayer", "StronglyEntanglingLayerBuilt", "StronglyEntanglingLayerPacked",
           "StronglyEntanglingLayerPackedBuilt"]


class StronglyEntanglingLayer(op.UnbuiltOperator):
    def __init__(self, qubits: typing.List[int], num_qubits_total: typing.Union[int, None] = None, depth: int = 1,
                 rotations=("rz", "ry", "rz"), q_params=None,
                 remapping: typing.Union[typing.Callable, None] = config.DEFAULT_MAPPING):
        self.depth = depth
        self.qubits = qubits
        self.num_qubits = num_qubits_total
        self.rots = [utils.parse_rot(r) for r in rotations]
        if q_params is None:
            q_params = torch.rand(depth, len(qubits), len(rotations))
        self.q_params = q_params
        if remapping is None:
            remapping = remap.none
        self.remapping = remapping

    def build(self, *args, **kwargs) -> "StronglyEntanglingLayerBuilt":
        return StronglyEntanglingLayerBuilt(num_qubits=kwargs["num_qubits"], depth=self.depth, rotations=self.rots,
                                            q_params=self.q_params, remapping=self.remapping, qubits=self.qubits)

    def __str__(self) -> str:
        return "SEL"

    def decompose(self) -> typing.List[op.UnbuiltOperator]:
        """Unbuilt form (reference stronglyentangling.py:60-78; needs num_qubits_total, quirk Q12)."""
        layers = []
        for d in range(self.depth):
            for wi, w in enumerate(self.qubits):
                for r in range(len(self.rots)):
                    layers.append(self.rots[r](qubit=w, theta=self.q_params[d, wi, r], remapping=self.remapping))
            it = d % (self.num_qubits - 1)
            for ci in range(len(self.qubits)):
                layers.append(op.CNOT(self.qubits[ci], self.qubits[(ci + it + 1) % len(self.qubits)]))
        return layers


class StronglyEntanglingLayerBuilt(op.BuiltOperator):
    def __init__(self, num_qubits: int, qubits: typing.List[int], depth: int, rotations, q_params: torch.Tensor,
                 remapping: typing.Union[typing.Callable, None]):
        super().__init__()
        self.num_qubits = num_qubits
        self.depth = depth
        self.rots = rotations
        self.qubits = qubits
        layers = []
        for d in range(depth):
            for wi, w in enumerate(qubits):
                for r in range(len(rotations)):
                    layers.append(rotations[r](qubit=w, theta=q_params[d, wi, r], remapping=remapping).build(num_qubits))
            layers.extend(self._get_cnots(qubits, num_qubits, d % (len(qubits) - 1)))
        self.mods = torch.nn.Sequential(*layers)

    @staticmethod
    def _get_cnots(qubits, num_qubits_total: int, iteration: int):
        assert iteration + 1 < num_qubits_total
        nq = len(qubits)
        return [op.CNOT(qubits[ci], qubits[(ci + iteration + 1) % nq]).build(num_qubits_total) for ci in range(nq)]

    def __str__(self) -> str:
        return "SEL"

    def decompose(self):
        return [copy.deepcopy(m) for m in self.mods]

    def to_matrix(self, **kwargs):
        return op.chain_matrices(self, self.mods, **kwargs)


class StronglyEntanglingLayerPacked(StronglyEntanglingLayer):
    """Same circuit as StronglyEntanglingLayer, but the built module owns ONE Parameter ``q_params`` of shape
    (depth, len(qubits), len(rotations)) instead of one 0-dim Parameter per gate: the
    per-step angle gather is a single remap of one tensor and autograd accumulates one gradient tensor.  Not
    state_dict-compatible with the reference's per-gate layout (use StronglyEntanglingLayer for that)."""

    def build(self, *args, **kwargs) -> "StronglyEntanglingLayerPackedBuilt":
        return StronglyEntanglingLayerPackedBuilt(num_qubits=kwargs["num_qubits"], depth=self.depth, rotations=self.rots,
                                                  q_params=self.q_params, remapping=self.remapping, qubits=self.qubits)


class StronglyEntanglingLayerPackedBuilt(op.BuiltOperator):
    def __init__(self, num_qubits: int, qubits: typing.List[int], depth: int, rotations, q_params: torch.Tensor,
                 remapping: typing.Callable):
        super().__init__()
        self.num_qubits = num_qubits
        self.depth = depth
        self.rots = rotations
        self.qubits = list(qubits)
        self.remapping = remapping
        self.q_params = torch.nn.Parameter(q_params.detach().clone().to(torch.float32), requires_grad=True)

    def __str__(self) -> str:
        return "SEL(packed)"

    def engine_lower_packed(self, slot0: int):
        """Gate-program rows (same order as reference stronglyentangling.py:93-121) and the number of slots used."""
        nq, nr = len(self.qubits), len(self.rots)
        rows = []
        for d in range(self.depth):
            for wi, w in enumerate(self.qubits):
                for r in range(nr):
                    rows.append((op.BUILT_CLASS_RELATION[self.rots[r]].engine_opcode, w, -1, slot0 + (d * nq + wi) * nr + r))
            it = d % (nq - 1)
            for ci in range(nq):
                rows.append((op.BuiltCNOT.engine_opcode, self.qubits[ci], self.qubits[(ci + it + 1) % nq], 0))
        return rows, self.depth * nq * nr

    def decompose(self):
        flat = StronglyEntanglingLayerBuilt(self.num_qubits, self.qubits, self.depth, self.rots, self.q_params.detach(), self.remapping)
        return flat.decompose()

    def to_matrix(self, **kwargs):
        return StronglyEntanglingLayerBuilt(self.num_qubits, self.qubits, self.depth, self.rots, self.q_params.detach(),
                                            self.remapping).to_matrix(**kwargs)
