"""ctypes front end of the C restatement (oracle/statevec_c.c).  TEST INFRASTRUCTURE ONLY: see the header of that file.

``build()`` compiles it with gcc (-O3 -fopenmp) into oracle/libqandle_oracle.so (git-ignored; it travels to the GPU box
with the snapshot).  ``run_program`` mirrors oracle.statevec.run_program for the forward pass in float64.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_DIR, "statevec_c.c")
LIB = os.path.join(_DIR, "libqandle_oracle.so")
_lib = None

MEASURE_STATE, MEASURE_PROBS, MEASURE_JOINT = 0, 1, 2


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.run(["gcc", "-O3", "-march=x86-64-v2", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"], check=True)
    return LIB


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        _lib.qo_run_program.restype = ctypes.c_int
    return _lib


def run_program(program, n: int, shared, batch_angles, fixed_mats, init_state, batch: int, measure: int) -> np.ndarray:
    """Forward of the gate-program IR in complex128 (same arguments as oracle.statevec.run_program; numpy / torch inputs).
    fixed_mats[slot] is applied as M psi.  Returns the state (B, 2^n) complex128, P(q=0) (B, n) or |psi|^2 (B, 2^n)."""
    as_np = lambda a, dt: np.ascontiguousarray(a.detach().cpu().numpy() if hasattr(a, "detach") else a, dtype=dt)
    prog = as_np(program, np.int32).reshape(-1, 4)
    sh = as_np(shared if shared is not None else np.zeros(0), np.float64).reshape(-1)
    ba = as_np(batch_angles if batch_angles is not None else np.zeros((batch, 0)), np.float64).reshape(batch, -1)
    fm = as_np(fixed_mats if fixed_mats is not None else np.zeros((0, 2, 2)), np.complex128).reshape(-1, 2, 2)
    N = 2**n
    if init_state is None:
        st = np.zeros((batch, N), np.complex128)
        st[:, 0] = 1
    else:
        st = as_np(init_state, np.complex128).reshape(-1, N)
        if st.shape[0] != batch:
            st = np.ascontiguousarray(np.broadcast_to(st, (batch, N)))
        st = st.copy()
    probs = np.zeros((batch, n), np.float64) if measure == MEASURE_PROBS else None
    joint = np.zeros((batch, N), np.float64) if measure == MEASURE_JOINT else None
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None and a.size else None
    rc = lib().qo_run_program(P(prog), len(prog), n, P(sh), P(ba), ba.shape[1], P(fm), ctypes.c_int64(batch), P(st), P(probs), P(joint))
    if rc:
        raise ValueError(f"bad opcode in program row {rc - 1}")
    return st if measure == MEASURE_STATE else (probs if measure == MEASURE_PROBS else joint)
