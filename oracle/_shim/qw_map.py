"""Stand-in for the un-vendored PyPI dependency qW-Map 0.1.2 (reference pyproject.toml:26,
poetry.lock:1887-1888).  TEST INFRASTRUCTURE ONLY: used by tests/golden/generate_golden.py so that
`import qandle` works in the build container.  Only `tanh` and `none` are used by the reference
(config.py:3, operators.py:168-171).  Semantics assumed from arXiv 2212.14807: tanh(x) = pi*tanh(x).
PARITY UNPINNED for the remap itself: no reference test pins a remapped value."""
import math
import torch


def none(x):
    return x


def tanh(x):
    return math.pi * torch.tanh(x)
