"""Pin the CPU oracle (oracle/statevec.py) against golden vectors produced by the real reference
(tests/golden/generate_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import statevec as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ir_cases.npz")


def _cases():
    z = np.load(GOLD)
    return z, int(z["count"])


def _load(z, i):
    g = lambda k: z[f"ir{i}/{k}"]
    return g


@pytest.mark.parametrize("i", range(20))
@pytest.mark.parametrize("mode", ["fast", "dense"])
def test_oracle_matches_reference(i, mode):
    z, count = _cases()
    assert count == 20
    g = _load(z, i)
    n, B, measure = int(g("n")), int(g("B")), int(g("measure"))
    if mode == "dense" and n > 8:
        pytest.skip("dense restatement is O(4^n)")
    prog = g("program")
    shared = torch.tensor(g("shared"), requires_grad=True)
    batch = torch.tensor(g("batch"), requires_grad=True)
    mats_user = torch.tensor(g("mats_user"))
    init = torch.tensor(g("init"), requires_grad=True) if bool(g("has_init")) else None
    if mode == "fast":
        # engine convention: fixed mats are applied as M psi => pass the transpose of the user's matrix
        out = O.run_program(prog, n, shared, batch, mats_user.transpose(-1, -2), init, B, measure)
    else:
        out = O.dense_run_program(prog, n, shared, batch, mats_user, init, B, measure)
    ref = torch.tensor(g("out"))
    assert out.shape == ref.shape
    assert torch.allclose(out, ref, rtol=1e-5, atol=2e-6), (out - ref).abs().max()
    out.backward(torch.tensor(g("g")))
    gs, gb = torch.tensor(g("g_shared")), torch.tensor(g("g_batch"))
    scale = max(1.0, float(gs.abs().max()))
    assert torch.allclose(shared.grad, gs, rtol=1e-4, atol=2e-5 * scale), (shared.grad - gs).abs().max()
    if batch.grad is not None:
        assert torch.allclose(batch.grad, gb, rtol=1e-4, atol=2e-5 * scale)
    else:
        assert gb.abs().max() == 0
    if init is not None:
        assert torch.allclose(init.grad, torch.tensor(g("g_init")), rtol=1e-4, atol=2e-5)


def test_oracle_complex128_consistent_with_complex64():
    z, _ = _cases()
    g = _load(z, 15)
    n, B, measure = int(g("n")), int(g("B")), int(g("measure"))
    out32 = O.run_program(g("program"), n, torch.tensor(g("shared")), torch.tensor(g("batch")),
                          torch.tensor(g("mats_user")).transpose(-1, -2), None, B, measure)
    out64 = O.run_program(g("program"), n, torch.tensor(g("shared")).double(), torch.tensor(g("batch")).double(),
                          torch.tensor(g("mats_user")).transpose(-1, -2).to(torch.complex128), None, B, measure)
    assert out64.dtype == torch.float64
    assert torch.allclose(out32.double(), out64, atol=1e-5)


def test_sel_program_matches_reference_layout():
    # ansaetze/stronglyentangling.py:93-121 -- 3 rotations per qubit then CNOT ring with range d%(nq-1)+1
    rows = O.sel_program([0, 1, 2, 3], depth=3)
    assert len(rows) == 3 * (12 + 4)
    cn = [r for r in rows if r[0] == O.OP_CNOT]
    assert cn[0][1:3] == (0, 1) and cn[4][1:3] == (0, 2) and cn[8][1:3] == (0, 3) and cn[11][1:3] == (3, 2)
