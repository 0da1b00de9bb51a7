"""Input embeddings: same surface as reference src/qandle/embeddings.py.

AngleEmbedding is lowered to per-sample rotation gates on |0...0> (what the reference's O(B d 8^n) matrix
chain, embeddings.py:148-164, computes); AmplitudeEmbedding pads / normalises in torch and enters the
engine as the initial state (O(B 2^n), once per step; gradients flow through torch autograd).
"""
import typing

import torch

from . import operators as op
from . import utils

__all__ = ["AmplitudeEmbedding", "AmplitudeEmbeddingBuilt", "AngleEmbedding", "AngleEmbeddingBuilt"]


class InputOperator(op.UnbuiltOperator):
    pass


class InputOperatorBuilt(op.BuiltOperator):
    def decompose(self):
        raise NotImplementedError(f"Decomposing {self.__class__} is not yet supported")

    def to_matrix(self):
        raise ValueError("Input operators do not have a matrix representation")

    def forward(self, state=None, **kwargs):
        from . import qcircuit

        return qcircuit.run_modules(self, [self], self.num_qubits, state, kwargs)


class AmplitudeEmbeddingBuilt(InputOperatorBuilt):
    """Ignores the incoming state and sets the features as the state (reference embeddings.py:29-61)."""

    def __init__(self, name: str, num_qubits: int, normalize: bool, pad_with: typing.Union[float, None]):
        super().__init__()
        self.num_qubits = num_qubits
        self.normalize = bool(normalize)
        self.pad_with = pad_with
        self.named = True
        self.name = name

    def embed(self, x: torch.Tensor) -> torch.Tensor:
        """pad -> L2-normalise -> complex cast (reference embeddings.py:50-61)."""
        if self.pad_with is not None:
            x = torch.nn.functional.pad(x, (0, 2**self.num_qubits - x.shape[-1]), mode="constant", value=self.pad_with)
        if self.normalize:
            x = torch.nn.functional.normalize(x, p=2, dim=-1)
        return torch.complex(x, torch.zeros_like(x))

    def __str__(self) -> str:
        return "AmplitudeEmbedding"


class AmplitudeEmbedding(InputOperator):
    """reference embeddings.py:70-111."""

    def __init__(self, name: str, qubits: list, normalize: bool = False, pad_with: typing.Union[float, None] = None):
        self.name = name
        self.named = True
        self.qubits = qubits
        self.pad_with = pad_with
        self.normalize = normalize

    def build(self, num_qubits: int, **kwargs) -> AmplitudeEmbeddingBuilt:
        assert num_qubits == len(self.qubits), "Current Implementation requires all qubits to be used."
        return AmplitudeEmbeddingBuilt(name=self.name, num_qubits=num_qubits, normalize=self.normalize, pad_with=self.pad_with)

    def __str__(self) -> str:
        return "AmplitudeEmbedding"


class AngleEmbeddingBuilt(InputOperatorBuilt):
    """psi = (x)_q R_rot(x_q)|0>; ignores the incoming state; inputs are not remapped (reference embeddings.py:114-186)."""

    def __init__(self, name: str, num_qubits: int, qubits: typing.List[int], rotation: str):
        super().__init__()
        self.num_qubits = num_qubits
        self.qubits = list(qubits)
        self.named = True
        self.name = name
        self.rotation = rotation
        self.rots = utils.parse_rot(rotation)
        self.engine_opcode = op.BUILT_CLASS_RELATION[self.rots].engine_opcode

    def __str__(self) -> str:
        return f"AngleEmbedding_{self.rotation}"

    def decompose(self):
        return [self.rots(qubit=w, name=f"angle_{id(self)}_{w}") for w in self.qubits]


class AngleEmbedding(InputOperator):
    """reference embeddings.py:184-214."""

    def __init__(self, name: str, qubits: typing.Union[typing.List[int], None] = None, rotation="rx"):
        self.name = name
        self.named = True
        self.qubits = qubits
        self.rotation = rotation

    def build(self, num_qubits: int, **kwargs) -> AngleEmbeddingBuilt:
        qubits = list(range(num_qubits)) if self.qubits is None else self.qubits
        return AngleEmbeddingBuilt(name=self.name, qubits=qubits, num_qubits=num_qubits, rotation=self.rotation)

    def __str__(self) -> str:
        return f"AngleEmbedding_{self.rotation}"
