"""world_size > 1 tests on CPU (gloo) of the multi-GPU host logic (qandle_b200/distributed.py):

* ShardedDriver: sweep runs + exchange steps over an all-to-all, forward and adjoint backward, with the numpy
  plan interpreter standing in for the CUDA kernels (the driver is backend-agnostic; on the GPU box the same
  driver runs _CudaBackend).  Result must equal the single-state oracle.
* allreduce_gradients for batch data-parallel.
No engine compute entry point is called here.
"""
import os
import random
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import plan_emulator as E
from oracle import statevec as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class EmulatorBackend:
    """numpy stand-in for the step-level kernels; per-rank K' accumulators like the workspace on the device."""

    def __init__(self, pd, shared, rank):
        self.pd, self.shared, self.rank = pd, shared, rank
        self.ms, self.mb = E.build_mats(pd, 1, shared, None, None)
        self.Ks = np.zeros((max(pd["n_k_shared"], 1), 3))
        self.Kb = np.zeros((1, max(pd["n_k_batch"], 1), 3))

    def _sweeps(self, s0, s1):
        return [self.pd["sweeps"][self.pd["steps"][i]["index"]] for i in range(s0, s1)]

    def apply_forward(self, s0, s1, state, rank):
        a = state.numpy()
        for sw in self._sweeps(s0, s1):
            a = E.sweep_forward(self.pd, sw, a, self.ms, self.mb, rank=rank)
        state.copy_(torch.from_numpy(a))

    def apply_backward(self, s0, s1, state, lam, rank):
        a, l = state.numpy(), lam.numpy()
        for sw in reversed(self._sweeps(s0, s1)):
            a, l = E.sweep_backward(self.pd, sw, a, l, self.ms, self.mb, self.Ks, self.Kb, rank=rank)
        state.copy_(torch.from_numpy(a))
        lam.copy_(torch.from_numpy(l))


def _sharded_worker(rank, world, port, n, prog, shared_np, g_np, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from qandle_b200 import distributed as qd
        from qandle_b200 import engine

        g_bits = world.bit_length() - 1
        n_local = n - g_bits
        plan = engine.Plan(torch.tensor(prog, dtype=torch.int32).reshape(-1, 4), n, engine.C128, (3, 1, 0, n_local, 1, 0, 1, 0))
        pd = engine.parse_plan_dump(plan.dump().tolist())
        be = EmulatorBackend(pd, shared_np, rank)
        drv = qd.ShardedDriver(plan.step_types(), be, rank, world, pieces=2)
        assert drv.n_exchanges >= 1
        state = torch.zeros(1, 2**n_local, dtype=torch.complex128)
        if rank == 0:
            state[0, 0] = 1
        drv.forward(state)
        # local probabilities (what probs_partial/finalize kernels do), then the all-reduce
        p = (state.abs() ** 2).numpy()
        idx = np.arange(2**n_local)
        probs = np.zeros((1, n))
        for q in range(n):
            b = pd["final_pos"][q]
            if b >= n_local:
                probs[:, q] = 0.0 if (rank >> (b - n_local)) & 1 else p.sum(axis=1)
            else:
                probs[:, q] = p[:, ((idx >> b) & 1) == 0].sum(axis=1)
        pt = torch.from_numpy(probs)
        dist.all_reduce(pt)
        # adjoint: local seed, reverse driver, finalize partial gradients, all-reduce
        w = np.zeros(2**n_local)
        for q in range(n):
            b = pd["final_pos"][q]
            if b >= n_local:
                w += g_np[0, q] * (0.0 if (rank >> (b - n_local)) & 1 else 1.0)
            else:
                w += g_np[0, q] * (((idx >> b) & 1) == 0)
        lam = state * torch.from_numpy(w)[None, :]
        drv.backward(state, lam)
        gs, _gb = E.finalize_grads(pd, 1, shared_np, None, None, be.Ks, be.Kb, len(shared_np), 0)
        gt = torch.from_numpy(gs)
        dist.all_reduce(gt)
        if rank == 0:
            ret["probs"] = pt.numpy().copy()
            ret["grads"] = gt.numpy().copy()
            ret["psi0"] = state.numpy().copy()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 6), (4, 8)])
def test_sharded_driver_matches_oracle_over_gloo(world, n):
    from test_planner_emulation import random_program

    rng = random.Random(world * 100 + n)
    prog = random_program(rng, n, 80, 6, 0, 0, p2=0.4)
    gen = torch.Generator().manual_seed(n)
    shared = ((torch.rand(6, generator=gen, dtype=torch.float64) - 0.5) * 6).requires_grad_(True)
    g = torch.randn(1, n, generator=gen, dtype=torch.float64)
    ref = O.run_program(prog, n, shared, None, None, None, 1, O.MEASURE_PROBS)
    ref.backward(g)
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_sharded_worker, args=(world, port, n, prog, shared.detach().numpy(), g.numpy(), ret), nprocs=world, join=True)
    assert np.allclose(ret["probs"], ref.detach().numpy(), atol=1e-12)
    assert np.allclose(ret["grads"], shared.grad.numpy(), atol=1e-10)
    psi0 = ret["psi0"]
    assert abs(psi0[0, 0] - 1) < 1e-12 and np.abs(psi0[0, 1:]).max() < 1e-12  # un-computed back to |0..0> on rank 0


def _dp_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from qandle_b200 import distributed as qd

        ps = [torch.nn.Parameter(torch.zeros(3)), torch.nn.Parameter(torch.zeros(())), torch.nn.Parameter(torch.zeros(2, 2))]
        for i, p in enumerate(ps):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        qd.allreduce_gradients(ps)
        if rank == 0:
            ret["g"] = [p.grad.clone().numpy() for p in ps]
    finally:
        dist.destroy_process_group()


def test_batch_dp_gradient_allreduce_gloo():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_dp_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    for i, g in enumerate(ret["g"]):
        assert np.allclose(g, 3.0 * (i + 1))
