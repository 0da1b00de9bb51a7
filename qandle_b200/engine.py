"""Python side of the torch.library layer: loads the native libraries, wraps plans, registers the
autograd bridge (adjoint-state backward instead of torch's tape, SURVEY.md 7 / north_star (c)).

There is no CPU fallback: if the CUDA extension cannot be loaded or no CUDA device is present, every
compute entry raises.
"""
from __future__ import annotations

import ctypes
import os
import threading
import typing

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
# QB_LIB_DIR: load a second in-tree build of the native libraries (kernel A/B experiments, qandle_b200/csrc/build.py --variant)
_LIB_DIR = os.environ.get("QB_LIB_DIR") or _PKG
LIB_CORE = os.path.join(_LIB_DIR, "libqandle_b200.so")
LIB_TORCH = os.path.join(_LIB_DIR, "libqandle_b200_torch.so")

# opcodes (include/qandle_b200.h)
OP_RX, OP_RY, OP_RZ, OP_U, OP_CNOT, OP_CZ, OP_SWAP = 1, 2, 3, 4, 5, 6, 7
FLAG_BATCH = 0x100
MEASURE_STATE, MEASURE_PROBS, MEASURE_JOINT = 0, 1, 2
C64, C128 = 0, 1
STEP_SWEEP, STEP_EXCHANGE = 0, 1

_lock = threading.Lock()
_core: typing.Optional[ctypes.CDLL] = None
_ops_loaded = False


class EngineUnavailable(RuntimeError):
    pass


def _ensure_built():
    if os.path.exists(LIB_CORE) and os.path.exists(LIB_TORCH):
        return
    try:
        from .csrc import build

        build.build_all()
    except Exception as e:  # noqa: BLE001
        raise EngineUnavailable(
            "qandle_b200: the native CUDA extension is missing and could not be built "
            f"({e}). Run `python -m qandle_b200.csrc.build`. There is no CPU fallback."
        ) from e


def core() -> ctypes.CDLL:
    """The C-ABI library (include/qandle_b200.h) via ctypes."""
    global _core
    with _lock:
        if _core is None:
            _ensure_built()
            lib = ctypes.CDLL(LIB_CORE)
            lib.qb_last_error.restype = ctypes.c_char_p
            lib.qb_version.restype = ctypes.c_char_p
            lib.qb_plan_dump.restype = ctypes.c_int64
            lib.qb_workspace_bytes.restype = ctypes.c_int64
            lib.qb_plan_algorithmic_bytes.restype = ctypes.c_int64
            _core = lib
    return _core


def load_ops():
    """Load the torch.library layer (namespace torch.ops.qandle_b200)."""
    global _ops_loaded
    with _lock:
        if not _ops_loaded:
            _ensure_built()
            torch.ops.load_library(LIB_TORCH)
            _ops_loaded = True
            _register_library()
    return torch.ops.qandle_b200


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise EngineUnavailable(
            "qandle_b200 needs a CUDA device (kernels are built for sm_100a / B200); there is no CPU fallback"
        )
    return torch.device("cuda", torch.cuda.current_device())


class Plan:
    """A compiled gate program (qb_plan).  opts = (tile_bits, low_bits, fuse, n_local, host_only, swap_relabel,
    final_layout, max_ops_per_sweep, staged, packed, flat, narrow_sync, exchange_any_bit, sweep_search) as in qb_plan_opts."""

    def __init__(self, program: torch.Tensor, n_qubits: int, dtype: int, opts: typing.Sequence[int] = ()):
        ops = load_ops()
        self.program = program.to(torch.int32).reshape(-1, 4).contiguous().cpu()
        self.n_qubits = int(n_qubits)
        self.dtype = int(dtype)
        self.opts = [int(o) for o in opts] + [0] * (16 - len(opts))
        self.handle = ops.plan_create(self.program, self.n_qubits, self.dtype, self.opts)
        info = ops.plan_info(self.handle).tolist()
        self.num_steps, self.num_sweeps, self.num_groups, self.launches_fwd, self.launches_bwd = info

    def __del__(self):
        h, self.handle = getattr(self, "handle", 0), 0
        if h:
            try:
                torch.ops.qandle_b200.plan_destroy(h)
            except Exception:  # noqa: BLE001  (interpreter shutdown)
                pass

    def __deepcopy__(self, memo):
        return self  # plans are immutable; share instead of duplicating the native handle

    def __reduce__(self):
        raise TypeError("engine plans hold native handles and cannot be pickled; they are rebuilt on demand")

    def dump(self) -> torch.Tensor:
        return torch.ops.qandle_b200.plan_dump(self.handle)

    def step_types(self):
        lib = core()
        return [lib.qb_plan_step_type(ctypes.c_void_p(self.handle), i) for i in range(self.num_steps)]

    def exchange_bits(self, step: int):
        """Local index bits exchange step `step` swaps with the rank bits (rank bit j <-> bits[j])."""
        lib = core()
        buf = (ctypes.c_int32 * 4)()
        g = lib.qb_plan_exchange_bits(ctypes.c_void_p(self.handle), int(step), buf)
        if g < 0:
            raise ValueError(f"step {step} is not an exchange step")
        return list(buf)[:g]

    def final_pos(self):
        lib = core()
        buf = (ctypes.c_int32 * self.n_qubits)()
        lib.qb_plan_final_pos(ctypes.c_void_p(self.handle), buf)
        return list(buf)

    def algorithmic_bytes(self, batch: int, backward: bool) -> int:
        return int(core().qb_plan_algorithmic_bytes(ctypes.c_void_p(self.handle), ctypes.c_int64(batch), int(backward)))


def parse_plan_dump(words) -> dict:
    """Decode qb_plan_dump (format: qandle_b200/csrc/plan.h)."""
    w = [int(x) for x in words]
    assert w[0] == 0x5142504C414E, "bad plan dump"
    it = iter(w[1:])
    nxt = lambda: next(it)
    d = dict(n_qubits=nxt(), n_local=nxt(), dtype=nxt())
    n_groups, n_members, n_steps, n_sweeps = nxt(), nxt(), nxt(), nxt()
    d.update(n_groups_shared=nxt(), n_groups_batch=nxt(), n_k_shared=nxt(), n_k_batch=nxt())
    keys = ("qubit", "member_begin", "member_count", "batch", "diag", "has_param", "mat_index", "k_index")
    d["groups"] = [dict(zip(keys, [nxt() for _ in keys])) for _ in range(n_groups)]
    d["members"] = [dict(kind=nxt(), slot=nxt(), batch=nxt()) for _ in range(n_members)]
    d["steps"] = [dict(type=nxt(), index=nxt()) for _ in range(n_steps)]
    d["final_pos"] = [nxt() for _ in range(d["n_qubits"])]
    sweeps = []
    for _ in range(n_sweeps):
        m, n_ops, n_ks, ext = nxt(), nxt(), nxt(), nxt()
        tile_bits = [nxt() for _ in range(m)]
        def get_ops(k):
            out = []
            for _ in range(k):
                kind, a, c, mat, ext_mask, ext_bit, kslot, rr = (nxt() for _ in range(8))
                out.append(dict(kind=kind, a=a, c=c, mat=mat, ext_mask=ext_mask, ext_bit=ext_bit, kslot=kslot,
                                r=(rr & 0xFF) - 1, rc=((rr >> 8) & 0xFF) - 1))
            return out

        def get_stages():
            out = []
            for _ in range(nxt()):
                low, r0, r1, r2, r3, ob, oe, pe, sb, flat, la, de, u0, u1, u2, u3, shape, nsg, nph, xth = (nxt() for _ in range(20))
                out.append(dict(low=low, regbits=[r for r in (r0, r1, r2, r3) if r >= 0], op_begin=ob, op_end=oe,
                                pre_end=pe, suf_begin=sb, flat=flat, la_end=la, d_end=de, u_op=[u0, u1, u2, u3],
                                shape=shape, n_sign=nsg, n_phase=nph, xthread=xth & 1,
                                narrow_end=(xth >> 4) & 3, narrow_x=(xth >> 8) & 3))
            return out

        ops = get_ops(n_ops)
        kslots = [dict(batch=nxt(), k_index=nxt()) for _ in range(n_ks)]
        stages = get_stages()
        ops_bwd = get_ops(nxt())
        stages_bwd = get_stages()
        sweeps.append(dict(tile_bits=tile_bits, ops=ops, kslots=kslots, has_ext_diag_param=ext, stages=stages,
                           ops_bwd=ops_bwd, stages_bwd=stages_bwd))
    d["sweeps"] = sweeps
    d["exchanges"] = [[nxt() for _ in range(nxt())] for _ in range(nxt())]  # per exchange step: local bit of rank bit j
    return d


# ---------------------------------------------------------------------------------------------------------
# torch.library registrations on top of the C++ CUDA kernels: fake (meta) implementations, so that FakeTensor tracing /
# torch.compile see shapes and dtypes without running a kernel, and the autograd formula of circuit_forward -- ONE custom-op call
# (circuit_backward: adjoint-state method) instead of torch's tape over the gate loop.
_registered = False


def _real_dtype(t: torch.Tensor) -> torch.dtype:
    return torch.float64 if t.dtype in (torch.float64, torch.complex128) else torch.float32


def _register_library():
    global _registered
    if _registered:
        return
    _registered = True

    @torch.library.register_fake("qandle_b200::circuit_forward")
    def _(plan, shared_angles, batch_angles, fixed_mats, init_state, batch, n_qubits, measure):
        real = _real_dtype(shared_angles)
        cd = torch.complex128 if real == torch.float64 else torch.complex64
        N = 1 << n_qubits
        state = shared_angles.new_empty((batch, N), dtype=cd)
        if measure == MEASURE_STATE:
            return state, shared_angles.new_empty((0,), dtype=cd)
        out = shared_angles.new_empty((batch, n_qubits if measure == MEASURE_PROBS else N), dtype=real)
        return out, state

    @torch.library.register_fake("qandle_b200::circuit_backward")
    def _(plan, shared_angles, batch_angles, fixed_mats, state, grad_out, measure, want_init_grad):
        g_init = torch.empty_like(state) if want_init_grad else state.new_empty((0,))
        return torch.empty_like(shared_angles), torch.empty_like(batch_angles), g_init

    def setup_context(ctx, inputs, output):
        plan, shared, batch, mats, init_state, _batch_size, _n, measure = inputs
        out, state = output
        ctx.plan_handle, ctx.measure = plan, measure
        ctx.has_init = init_state is not None
        ctx.has_batch = batch.numel() > 0
        ctx.mark_non_differentiable(state)
        ctx.set_materialize_grads(False)  # (else autograd zero-fills a [B, 2^n] "gradient" of the state output every backward)
        # the final state is only READ by circuit_backward: saving it through autograd is safe for repeated backward calls
        ctx.save_for_backward(shared, batch, mats, out if measure == MEASURE_STATE else state)

    def backward(ctx, grad_out, _grad_state):
        if grad_out is None:
            return (None,) * 8
        shared, batch, mats, state = ctx.saved_tensors
        want_init = ctx.has_init and ctx.needs_input_grad[4]
        g = grad_out.contiguous()
        if ctx.measure == MEASURE_STATE and not g.is_complex():
            g = g.to(state.dtype)
        g_shared, g_batch, g_init = torch.ops.qandle_b200.circuit_backward(ctx.plan_handle, shared, batch, mats, state, g, ctx.measure, want_init)
        return (None, g_shared if ctx.needs_input_grad[1] else None, g_batch if (ctx.needs_input_grad[2] and ctx.has_batch) else None,
                None, g_init if want_init else None, None, None, None)

    torch.library.register_autograd("qandle_b200::circuit_forward", backward, setup_context=setup_context)


def run_circuit(plan: Plan, shared: torch.Tensor, batch: torch.Tensor, mats: torch.Tensor,
                init_state: typing.Optional[torch.Tensor], batch_size: int, measure: int) -> torch.Tensor:
    """Differentiable engine call: out = measure(G_K ... G_1 psi_0), backward = adjoint-state method in one custom op.
    All tensors must already be on the CUDA device and of the plan's dtype."""
    out, _state = torch.ops.qandle_b200.circuit_forward(plan.handle, shared, batch, mats, init_state, batch_size, plan.n_qubits, measure)
    return out
