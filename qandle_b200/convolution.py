"""QConv: quantum convolution layer -- the reference's one in-tree model-level caller of the hot path
(reference src/qandle/convolution.py:9-61).  Unfold -> AmplitudeEmbedding (pad with 0) ->
StronglyEntanglingLayer -> joint probabilities: a huge batch of tiny circuits, i.e. the engine's batch path."""
import math
import typing

import torch

from . import embeddings, measurements, qcircuit
from .ansaetze import StronglyEntanglingLayer

__all__ = ["QConv"]


class QConv(torch.nn.Module):
    def __init__(self, in_channels: int, out_channels: int, kernel_size: typing.Union[typing.Tuple[int, int], int] = (3, 3),
                 padding: typing.Union[typing.Tuple[int, int], int] = (1, 1), qdepth: int = 1):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.qdepth = qdepth
        self.kernel_size = kernel_size if isinstance(kernel_size, tuple) else (kernel_size, kernel_size)
        self.padding = padding if isinstance(padding, tuple) else (padding, padding)
        self.unfold = torch.nn.Unfold(kernel_size=kernel_size, padding=self.padding)
        qubits_for_inp = math.ceil(math.log2(self.kernel_size[0] * self.kernel_size[1] * in_channels))
        qubits_for_out = math.ceil(math.log2(out_channels))
        self.qubits = max(qubits_for_inp, qubits_for_out, 1)
        amp = embeddings.AmplitudeEmbedding(qubits=list(range(self.qubits)), pad_with=0, name="emb")
        sel = StronglyEntanglingLayer(qubits=list(range(self.qubits)), depth=self.qdepth)
        mes = measurements.MeasureJointProbability()
        self.qcircuit = qcircuit.Circuit(num_qubits=self.qubits, layers=[amp, sel, mes])
        self.almost_zero = 0.001

    def _post_process(self, x):
        x = x[:, : self.out_channels, :, :]  # remove padding
        return x * self.out_channels / 2  # rescale

    def forward(self, x):
        b, c_in, h_in, w_in = x.shape
        h_out = h_in + 2 * self.padding[0] - self.kernel_size[0] + 1
        w_out = w_in + 2 * self.padding[1] - self.kernel_size[1] + 1
        if c_in != self.in_channels:
            raise ValueError(f"Input channels {c_in} does not match in_channels {self.in_channels}")
        x = self.unfold(x)  # (b, c*kh*kw, L)
        x = x.permute(0, 2, 1).reshape(b * x.shape[2], x.shape[1])  # "(batch feat) channel"
        x = x + self.almost_zero  # avoid zero input
        x = self.qcircuit(emb=x)  # (b*L, 2^qubits)
        x = x.reshape(b, h_out, w_out, x.shape[-1]).permute(0, 3, 1, 2)  # "batch channel h_out w_out"
        return self._post_process(x)
