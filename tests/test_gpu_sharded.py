"""Amplitude sharding on >= 2 GPUs (run with `gpurun --gpus 2|4|8`): ShardedCircuit must equal the single-GPU
engine and the CPU oracle (SURVEY 8c tier T3: forced sharding at small n)."""
import os
import random
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import statevec as O

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _layers(q, n, depth, seed):
    rng = random.Random(seed)
    layers, rows, slot = [], [], 0
    for d in range(depth):
        for k in range(n):
            kind = rng.choice(["RX", "RY", "RZ"])
            layers.append(getattr(q, kind)(k, remapping=None))
            rows.append(({"RX": O.OP_RX, "RY": O.OP_RY, "RZ": O.OP_RZ}[kind], k, -1, slot))
            slot += 1
        for k in list(range(0, n - 1, 2)) + list(range(1, n - 1, 2)):
            layers.append(q.CZ(k, k + 1))
            rows.append((O.OP_CZ, k, k + 1, 0))
        if d % 2 == 1:
            layers.append(q.CNOT(n - 1, 0))
            rows.append((O.OP_CNOT, n - 1, 0, 0))
    layers.append(q.MeasureProbability())
    return layers, rows, slot


def _worker(rank, world, port, n, depth, dtype_name, pieces, exchange, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import qandle_b200 as q
        from qandle_b200.distributed import ShardedCircuit

        dtype = getattr(torch, dtype_name)
        torch.manual_seed(0)
        layers, _rows, _ = _layers(q, n, depth, seed=n)
        sc = ShardedCircuit(layers, num_qubits=n, pieces=pieces, tile_bits=6, low_bits=2, exchange=exchange).to(f"cuda:{rank}")
        with torch.no_grad():
            for p in sc.parameters():
                p.mul_(6.0)
        out = sc(dtype=dtype)
        g = torch.linspace(-1, 1, n, device=out.device, dtype=out.dtype)
        out.backward(g)
        if rank == 0:
            ret["out"] = out.detach().cpu().numpy()
            ret["grads"] = np.array([float(p.grad) for p in sc.parameters()])
            ret["thetas"] = np.array([float(p) for p in sc.parameters()])
            ret["n_exchanges"] = sum(1 for t in sc.step_types if t == 1)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,depth,dtype_name,pieces,exchange", [(12, 4, "float32", 1, "nccl"), (13, 3, "float64", 2, "nccl"),
                                                               (13, 4, "float32", 1, "p2p"), (12, 3, "float64", 1, "p2p"),
                                                               (13, 4, "float32", 2, "push"), (12, 3, "float64", 1, "push"),
                                                               (14, 3, "float32", 1, "auto")])
def test_sharded_circuit_matches_oracle(n, depth, dtype_name, pieces, exchange):
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 1 << (world.bit_length() - 1)
    world = min(world, 8)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, depth, dtype_name, pieces, exchange, ret), nprocs=world, join=True)
    import qandle_b200 as q

    _layers_unused, rows, n_slots = _layers(q, n, depth, seed=n)
    thetas = torch.tensor(ret["thetas"], dtype=torch.float64, requires_grad=True)
    ref = O.run_program(rows, n, thetas, None, None, None, 1, O.MEASURE_PROBS)
    ref.backward(torch.linspace(-1, 1, n, dtype=torch.float64).reshape(1, n))
    tol = 1e-5 if dtype_name == "float32" else 1e-12
    assert ret["n_exchanges"] >= 1
    assert np.abs(ret["out"] - ref.detach().numpy().reshape(-1)).max() < tol
    assert np.abs(ret["grads"] - thetas.grad.numpy()).max() < max(tol * 20, 1e-7)  # Parameters (and their .grad) are float32
