"""Multi-GPU modes of the engine (north_star (d)); one process per GPU, torch.distributed for the plumbing.

* batch data-parallel: every rank runs the same circuit on its own batch shard; the only exchange is the
  all-reduce of the weight gradients (``allreduce_gradients``).  Per-sample inputs / outputs never move.
* amplitude sharding: see ``ShardedCircuit`` -- rank r owns amplitudes [r 2^n/P, (r+1) 2^n/P) (qubit 0 is the
  most significant index bit, so the top log2 P qubits are the rank bits); plans contain exchange steps that
  swap the rank bits with the top local bits by one all-to-all.  Replaces the reference's circuit splitting
  (reference src/qandle/splitter/, qcircuit.py:215-314).
"""
from __future__ import annotations

import typing

import torch
import torch.distributed as dist


def allreduce_gradients(params: typing.Iterable[torch.nn.Parameter], group=None, average: bool = False):
    """Sum (or average) the gradients of `params` over the ranks with ONE collective on a flat buffer."""
    ps = [p for p in params if p.grad is not None]
    if not ps or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in ps])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for p in ps:
        k = p.grad.numel()
        p.grad.copy_(flat[off:off + k].reshape(p.grad.shape))
        off += k
