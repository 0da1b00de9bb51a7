"""CUDA-graph replay for small circuits (SURVEY.md 7 "hard parts": per-gate Parameters and a dozen tiny launches per step make a
4-qubit hybrid layer launch-latency bound -- BASELINE config 1 runs 0.4 ms per step eagerly, of which the kernels are ~30 us).

``cuda_graph(circuit, **example_inputs)`` captures the circuit's forward and its adjoint backward (angle gather, remapping, the
engine's kernels, the gradient scatter) into two CUDA graphs with torch.cuda.make_graphed_callables and returns a module that is
called like the circuit -- same named inputs, same outputs, same .grad on the same Parameters -- but replays the graphs: one
graph launch forward, one backward.  Inputs must keep the example's shapes / dtypes / devices (a new shape needs a new capture).
"""
from __future__ import annotations

import typing

import torch

__all__ = ["cuda_graph"]


class _Positional(torch.nn.Module):
    def __init__(self, circuit: torch.nn.Module, names: typing.Sequence[str]):
        super().__init__()
        self.circuit = circuit
        self.names = list(names)

    def forward(self, *args):
        return self.circuit(**dict(zip(self.names, args)))


class GraphedCircuit(torch.nn.Module):
    """The circuit behind CUDA-graph replay; parameters are the wrapped circuit's own (state_dict keys are prefixed `circuit.`)."""

    def __init__(self, circuit: torch.nn.Module, example_inputs: typing.Dict[str, torch.Tensor]):
        super().__init__()
        for k, v in example_inputs.items():
            if not (torch.is_tensor(v) and v.is_cuda):
                raise ValueError(f"cuda_graph: input {k!r} must be a CUDA tensor (graphs replay on fixed device buffers)")
        self.circuit = circuit
        self.names = list(example_inputs)
        self.signature = {k: (tuple(v.shape), v.dtype, v.device) for k, v in example_inputs.items()}
        samples = tuple(example_inputs[k].detach().clone().requires_grad_(example_inputs[k].requires_grad) for k in self.names)
        # warm-up on a side stream (plan creation, workspace sizes, allocator pools), then capture forward and backward
        self._graphed = torch.cuda.make_graphed_callables(_Positional(circuit, self.names), samples)

    def forward(self, **kwargs):
        for k in self.names:
            v = kwargs[k]
            if (tuple(v.shape), v.dtype, v.device) != self.signature[k]:
                raise ValueError(f"cuda_graph: input {k!r} is {tuple(v.shape)} {v.dtype} on {v.device}, captured as {self.signature[k]}")
        return self._graphed(*[kwargs[k] for k in self.names])


def cuda_graph(circuit: torch.nn.Module, **example_inputs: torch.Tensor) -> GraphedCircuit:
    """Capture ``circuit(**example_inputs)`` and its backward into CUDA graphs; call the result like the circuit."""
    return GraphedCircuit(circuit, example_inputs)
