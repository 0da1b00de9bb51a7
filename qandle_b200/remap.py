"""Weight remapping functions.

The reference takes these from the un-vendored PyPI package qW-Map 0.1.2 (reference config.py:1-3,
operators.py:168-171); only ``tanh`` and ``none`` are used.  If ``qw_map`` is installed it is used
verbatim, so the user's remapping is honoured bit for bit; otherwise the semantics assumed from
arXiv 2212.14807 are provided (tanh(x) = pi * tanh(x); parity of this function is unpinned).
The engine never evaluates a remap inside a kernel: ``remapping(theta)`` stays a differentiable torch
call upstream of the custom op (operators.py:271)."""
import math

import torch

try:  # pragma: no cover - depends on the user's environment
    from qw_map import none, tanh  # type: ignore
except Exception:  # noqa: BLE001

    def none(x):
        return x

    def tanh(x):
        return math.pi * torch.tanh(x)
