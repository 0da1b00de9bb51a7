"""Golden vectors from the REAL reference at 12 and 13 qubits: the sizes where one sample fills one / two full 2^12-amplitude
tiles, so the engine's FULL-tile sweep kernels (the ones bench.py times) are compared with reference-generated numbers and not
only with the oracle port (the other golden cases stop at 10 qubits: partial tiles).

Run in the build container only (minutes and ~10 GB: every reference gate is a dense 2^n x 2^n matrix, 512 MB at 13 qubits):

    python tests/golden/generate_full_tile.py

Writes tests/golden/full_tile_cases.npz (inputs, initial states, parameters, outputs, cotangents and every gradient).
Nothing here is imported by the product.
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_shim"))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, HERE)

import qandle  # noqa: E402  (the reference)
import specs  # noqa: E402


def main():
    out = {}
    for idx, (name, case) in enumerate(specs.full_tile_specs().items()):
        t0 = time.time()
        torch.manual_seed(7000 + idx)
        gen = torch.Generator().manual_seed(8000 + idx)
        n, B = case["num_qubits"], case["batch"]
        circ = specs.build_circuit(qandle, case["spec"], n)
        inputs = {k: (torch.rand(tuple(s), generator=gen) * 2 - 1).requires_grad_(True) for k, s in case["inputs"].items()}
        st = torch.complex(torch.randn(B, 2**n, generator=gen), torch.randn(B, 2**n, generator=gen))
        state = (st / torch.linalg.norm(st, dim=-1, keepdim=True)).requires_grad_(True)
        res = circ(state, **inputs)
        g = torch.randn(res.shape, generator=gen)
        if res.is_complex():
            g = torch.complex(g, torch.randn(res.shape, generator=gen))
        res.backward(g)
        out[f"{name}/state"] = state.detach().numpy()
        out[f"{name}/out"] = res.detach().numpy()
        out[f"{name}/g"] = g.numpy()
        out[f"{name}/grad_state"] = state.grad.numpy()
        for k, v in inputs.items():
            out[f"{name}/in.{k}"] = v.detach().numpy()
            out[f"{name}/gin.{k}"] = v.grad.numpy()
        for k, p in circ.named_parameters():
            out[f"{name}/p.{k}"] = p.detach().numpy()
            out[f"{name}/gp.{k}"] = np.full(p.shape, np.nan, np.float32) if p.grad is None else p.grad.numpy()  # NaN: no gradient
        print(f"{name}: {time.time() - t0:.0f} s", flush=True)
        np.savez_compressed(os.path.join(HERE, "full_tile_cases.npz"), **out)


if __name__ == "__main__":
    main()
