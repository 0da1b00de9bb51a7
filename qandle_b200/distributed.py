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


# =========================================================================================================
# amplitude sharding
class _CudaBackend:
    """Production backend of the sharded driver: step-level torch.library ops (sm_100a kernels)."""

    def __init__(self, plan, batch: int, device):
        from . import engine

        self.ops = engine.load_ops()
        self.plan, self.B, self.device = plan, batch, device
        self.ws = torch.empty(self.ops.workspace_bytes(plan.handle, batch) + 256, dtype=torch.uint8, device=device)

    def new_state(self, n_local, cdtype):
        return torch.empty(self.B, 2**n_local, dtype=cdtype, device=self.device)

    def prepare(self, shared, batch_angles, mats):
        self.ops.prepare(self.plan.handle, self.B, shared, batch_angles, mats, self.ws)

    def init_zero(self, state, rank):
        self.ops.init_zero(self.plan.handle, self.B, state, rank)

    def apply_forward(self, s0, s1, state, rank):
        self.ops.apply_forward(self.plan.handle, s0, s1, self.B, state, self.ws, rank)

    def apply_backward(self, s0, s1, state, lam, rank):
        self.ops.apply_backward(self.plan.handle, s0, s1, self.B, state, lam, self.ws, rank)

    def measure_probs(self, state, rank):
        return self.ops.measure_probs(self.plan.handle, self.B, self.plan.n_qubits, state, self.ws, rank)

    def seed_probs(self, state, grad, lam, rank):
        self.ops.seed_probs(self.plan.handle, self.B, state, grad, lam, rank)

    def backward_begin(self):
        self.ops.backward_begin(self.plan.handle, self.B, self.ws)

    def finalize(self, shared, batch_angles, mats):
        return self.ops.finalize_grads(self.plan.handle, self.B, shared, batch_angles, mats, self.ws)


def exchange_inplace(state: torch.Tensor, world: int, group=None, pieces: int = 1, staging=None):
    """Swap the top log2(world) local index bits with the rank bits: rank r sends its c-th contiguous chunk to
    rank c and stores what it receives from c in the same place.  Because qubit 0 is the most significant index
    bit, this is exactly an all-to-all over contiguous equal chunks -- no pack/unpack kernels.  With pieces > 1
    the exchange is done in place through two staging buffers of 1/pieces of the shard (36-qubit states: the
    state and its adjoint leave no room for a full-size receive buffer)."""
    B = state.shape[0]
    v = state.view(B, world, -1)
    chunk = v.shape[2]
    assert chunk % pieces == 0
    ps = chunk // pieces
    for p in range(pieces):
        sl = v[:, :, p * ps:(p + 1) * ps]
        if staging is None:
            send = sl.permute(1, 0, 2).contiguous()  # [world][B][ps]: dim 0 = destination rank
            recv = torch.empty_like(send)
        else:
            send, recv = staging
            send = send[: world * B * ps].view(world, B, ps)
            recv = recv[: world * B * ps].view(world, B, ps)
            send.copy_(sl.permute(1, 0, 2))
        if torch.is_complex(send):
            dist.all_to_all_single(torch.view_as_real(recv), torch.view_as_real(send), group=group)
        else:
            dist.all_to_all_single(recv, send, group=group)
        sl.copy_(recv.permute(1, 0, 2))
    return state


class ShardedDriver:
    """Runs a plan that contains exchange steps.  Backend-agnostic so the step/exchange logic is testable on CPU
    (gloo) with the numpy plan interpreter as backend; production uses _CudaBackend."""

    def __init__(self, step_types: typing.Sequence[int], backend, rank: int, world: int, group=None, pieces: int = 1,
                 exchange_fn=None):
        self.steps = list(step_types)
        self.be, self.rank, self.world, self.group, self.pieces = backend, rank, world, group, pieces
        # exchange_fn(tensor, step): in-place global<->local swap of exchange step `step` (its bit positions: Plan.exchange_bits);
        # default = NCCL all-to-all through staging buffers, for plans whose exchanges swap the TOP local bits
        self.exchange_fn = exchange_fn or (lambda t, step: exchange_inplace(t, self.world, self.group, self.pieces))
        # maximal runs of sweep steps between exchanges
        self.runs = []
        i = 0
        while i < len(self.steps):
            if self.steps[i] == 0:
                j = i
                while j < len(self.steps) and self.steps[j] == 0:
                    j += 1
                self.runs.append(("sweeps", i, j))
                i = j
            else:
                self.runs.append(("exchange", i, i + 1))
                i += 1
        self.n_exchanges = sum(1 for r in self.runs if r[0] == "exchange")

    def forward(self, state):
        for kind, s0, s1 in self.runs:
            if kind == "sweeps":
                self.be.apply_forward(s0, s1, state, self.rank)
            else:
                self.exchange_fn(state, s0)
        return state

    def backward(self, state, lam):
        for kind, s0, s1 in reversed(self.runs):
            if kind == "sweeps":
                self.be.apply_backward(s0, s1, state, lam, self.rank)
            else:  # the exchange is an involution
                self.exchange_fn(state, s0)
                self.exchange_fn(lam, s0)
        return state, lam


class _ShardedFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sc: "ShardedCircuit", shared, batch_angles, B: int):
        be = _CudaBackend(sc.plan, B, shared.device)
        mats = sc._mats(shared.device, shared.dtype)
        be.prepare(shared, batch_angles, mats)
        cd = torch.complex128 if shared.dtype == torch.float64 else torch.complex64
        state = sc._buffer("psi", B, cd, shared.device) if sc.exchange == "p2p" else be.new_state(sc.n_local, cd)
        be.init_zero(state, sc.rank)
        drv = ShardedDriver(sc.step_types, be, sc.rank, sc.world, sc.group, sc.pieces, exchange_fn=sc._exchange_fn(B))
        drv.forward(state)
        probs = be.measure_probs(state, sc.rank)
        dist.all_reduce(probs, group=sc.group)
        ctx.sc, ctx.be, ctx.drv, ctx.state, ctx.mats = sc, be, drv, state, mats
        ctx.save_for_backward(shared, batch_angles)
        if sc.keep_state:
            sc.last_state = state
        return probs

    @staticmethod
    def backward(ctx, grad):
        sc, be, drv, state = ctx.sc, ctx.be, ctx.drv, ctx.state
        shared, batch_angles = ctx.saved_tensors
        lam = sc._buffer("lam", state.shape[0], state.dtype, state.device) if sc.exchange == "p2p" else torch.empty_like(state)
        be.seed_probs(state, grad.contiguous(), lam, sc.rank)
        be.backward_begin()
        drv.backward(state, lam)
        g_shared, g_batch = be.finalize(shared, batch_angles, ctx.mats)
        # every rank holds the partial <lambda|dG|psi> of its amplitudes
        if g_shared.numel():
            dist.all_reduce(g_shared, group=sc.group)
        if g_batch.numel():
            dist.all_reduce(g_batch, group=sc.group)
        ctx.state = None
        return None, g_shared, (g_batch if g_batch.numel() else None), None


_SYMM_OK: typing.Dict[int, bool] = {}
_SYMM_POOL: typing.Dict[typing.Tuple, typing.Tuple] = {}  # symmetric-memory state buffers, shared by circuits of one size


def _symmetric_memory_available(group=None) -> bool:
    """Can torch.distributed._symmetric_memory map a buffer of every rank of `group` into this process?  Probed once per group with
    a tiny allocation; every rank must call it (the rendezvous is collective) and all ranks agree on the answer."""
    key = id(group)
    if key not in _SYMM_OK:
        ok = torch.cuda.is_available() and dist.get_backend(group) == "nccl"
        if ok:
            try:
                import torch.distributed._symmetric_memory as symm_mem

                raw = symm_mem.empty(64, dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
                hdl = symm_mem.rendezvous(raw, group if group is not None else dist.group.WORLD)
                ok = len(hdl.buffer_ptrs) == dist.get_world_size(group)
            except Exception:  # noqa: BLE001
                ok = False
        flag = torch.tensor([1 if ok else 0], device="cuda" if torch.cuda.is_available() else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        _SYMM_OK[key] = bool(flag.item())
    return _SYMM_OK[key]


class ShardedCircuit(torch.nn.Module):
    """A circuit whose single state vector is sharded over the ranks of a process group (amplitude sharding).

    Same layer list as ``Circuit`` (gates, AngleEmbedding, StronglyEntanglingLayer, trailing
    MeasureProbability); the state starts from |0...0>; the result -- P(q = 0) for every qubit, shape (B, n)
    then ``.squeeze()`` like the reference -- is replicated on every rank.  Gradients of weights and named
    inputs are all-reduced, so every rank ends with the full gradient.
    """

    def __init__(self, layers, num_qubits: int, group=None, pieces: int = 1, tile_bits: int = 0, low_bits: int = 0,
                 keep_state: bool = False, exchange: str = "auto"):
        super().__init__()
        from . import config, engine, qcircuit

        assert dist.is_initialized(), "ShardedCircuit needs torch.distributed (one process per GPU)"
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        g = self.world.bit_length() - 1
        assert 1 << g == self.world, "world size must be a power of two"
        self.num_qubits = num_qubits
        self.n_local = num_qubits - g
        self.pieces = pieces
        assert exchange in ("auto", "nccl", "p2p", "push")
        # "push": every rank WRITES its outgoing chunks into its peers' staging buffers over NVLink (TMA bulk copies,
        #         qb_exchange_push_dev), then unpacks its own staging buffer locally; the chunk is cut into `pieces`, the staging
        #         buffer (symmetric memory) is 1 / pieces of the shard.  Posted writes run NVLink near its copy bandwidth.
        # "p2p":  the state lives in symmetric memory and an exchange is ONE in-place kernel that pulls half of every chunk pair
        #         through the peer mapping (qb_exchange_p2p_dev): no staging at all, but remote reads are latency-bound.
        # "nccl": all_to_all_single through staging (no peer mappings needed).
        # "auto": p2p when torch's symmetric memory can map the peers (one NVLink / NVSwitch box), else nccl.  Measured on 8 B200s
        #         (30 qubits complex128 / 36 qubits complex64, profiles/r2_multi_gpu.md): p2p 2.9 ms = 647 GB/s per direction per
        #         exchange, push 3.4 ms / 107 ms = 560 GB/s including its unpack pass and second barrier.
        if exchange == "auto":
            exchange = "p2p" if _symmetric_memory_available(group) else "nccl"
        self.exchange = exchange
        self._symm = {}
        self.keep_state = keep_state
        self.last_state = None
        self.circuit = qcircuit.UnsplittedCircuit(num_qubits, list(layers))
        segs = qcircuit.lower_modules(self.circuit.layers, num_qubits)
        assert len(segs) == 1 and segs[0].foreign is None, "ShardedCircuit supports one engine segment"
        self.seg = segs[0]
        assert self.seg.init in ("zero", "inherit"), "amplitude-sharded circuits start from |0...0>"
        assert self.seg.measure == engine.MEASURE_PROBS, "amplitude-sharded circuits end in MeasureProbability"
        # (the peer-memory exchange kernel handles any bit positions: let the planner pick the bits to send out per exchange)
        self._opts = (tile_bits, low_bits, 0, self.n_local, 0, 0, 1, config.ENGINE_MAX_OPS_PER_SWEEP,
                      0 if config.ENGINE_STAGED else -1, 0 if config.ENGINE_PACKED else -1, 0 if config.ENGINE_FLAT else -1, 0,
                      1 if exchange == "p2p" else 0)
        self._plans = {}
        self.plan = None
        self.step_types = None

    def _buffer(self, name, B, cdtype, device):
        """Persistent symmetric-memory buffer [B, 2^n_local] (peer-mapped on every rank).  Reused across calls: in
        p2p mode a new forward overwrites the state a not-yet-run backward of the previous call would need.  The buffers live in a
        process-wide pool keyed by (group, name, shape, dtype): two circuits of the same size share them (a 36-qubit shard and its
        adjoint are 2 x 64 GiB -- a second pair would not fit)."""
        import torch.distributed._symmetric_memory as symm_mem

        key = (name, B, cdtype)
        if key not in self._symm:
            pool_key = (id(self.group), name, B, self.n_local, cdtype, device)
            if pool_key not in _SYMM_POOL:
                real = torch.float64 if cdtype == torch.complex128 else torch.float32
                raw = symm_mem.empty(B * 2 ** self.n_local * 2, dtype=real, device=device)
                hdl = symm_mem.rendezvous(raw, self.group if self.group is not None else dist.group.WORLD)
                view = torch.view_as_complex(raw.view(B, 2 ** self.n_local, 2))
                _SYMM_POOL[pool_key] = (view, hdl, raw)
            self._symm[key] = _SYMM_POOL[pool_key]
        return self._symm[key][0]

    def _staging(self, B, cdtype, device):
        """Peer-mapped staging buffer of the push exchange: world slots of one piece of a chunk per sample."""
        import torch.distributed._symmetric_memory as symm_mem

        real = torch.float64 if cdtype == torch.complex128 else torch.float32
        numel = self.world * B * (2 ** self.n_local // self.world // self.pieces) * 2
        pool_key = (id(self.group), "stage", numel, real, device)
        if pool_key not in _SYMM_POOL:
            raw = symm_mem.empty(numel, dtype=real, device=device)
            hdl = symm_mem.rendezvous(raw, self.group if self.group is not None else dist.group.WORLD)
            _SYMM_POOL[pool_key] = (raw, hdl, raw)
        return _SYMM_POOL[pool_key][:2]

    def _exchange_push(self, t, B):
        from . import engine

        ops = engine.load_ops()
        _raw, hdl = self._staging(B, t.dtype, t.device)
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        for piece in range(self.pieces):
            hdl.barrier(channel=0)  # every peer has unpacked the previous piece / finished the sweep before
            ops.exchange_push(self.plan.handle, B, t, ptrs, self.rank, piece, self.pieces, 0)
            hdl.barrier(channel=0)  # every push has landed
            ops.exchange_push(self.plan.handle, B, t, ptrs, self.rank, piece, self.pieces, 1)

    def _exchange_fn(self, B):
        if self.exchange == "p2p":
            return lambda t, step: self._exchange_p2p(t, B, step)
        if self.exchange == "push":
            return lambda t, step: self._exchange_push(t, B)
        return None  # ShardedDriver's default: NCCL all-to-all

    def _exchange_p2p(self, t, B, step):
        from . import engine

        for view, hdl, _raw in self._symm.values():
            if view.data_ptr() == t.data_ptr():
                hdl.barrier(channel=0)
                engine.load_ops().exchange_p2p(self.plan.handle, B, t, [int(p) for p in hdl.buffer_ptrs], self.rank, step)
                hdl.barrier(channel=0)
                return
        raise RuntimeError("exchange_p2p: tensor is not one of this circuit's symmetric buffers")

    def exchange_time_ms(self, dtype: torch.dtype = torch.float32, reps: int = 2) -> float:
        """Milliseconds of ONE exchange step on a one-state shard (CUDA events, max over ranks) -- what bench.py divides the
        bytes on the wire by.  The exchange is an involution: 2 * reps calls leave the buffer as it was."""
        from . import engine

        dev = engine.require_cuda()
        self._ensure_plan(dtype)
        cd = torch.complex128 if dtype == torch.float64 else torch.complex64
        if self.exchange == "p2p":
            t = self._buffer("psi", 1, cd, dev)
            first = self.step_types.index(1)
            fn = lambda: self._exchange_p2p(t, 1, first)  # noqa: E731
        elif self.exchange == "push":
            t = torch.zeros(1, 2 ** self.n_local, dtype=cd, device=dev)
            fn = lambda: self._exchange_push(t, 1)  # noqa: E731
        else:
            t = torch.zeros(1, 2 ** self.n_local, dtype=cd, device=dev)
            fn = lambda: exchange_inplace(t, self.world, self.group, self.pieces)  # noqa: E731
        fn()
        fn()
        dist.barrier(self.group)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2 * reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / (2 * reps)], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX, group=self.group)
        return float(ms)

    def _mats(self, device, real_dtype):
        if not self.seg.mats:
            return torch.zeros(0, device=device, dtype=real_dtype)
        cd = torch.complex128 if real_dtype == torch.float64 else torch.complex64
        return torch.view_as_real(torch.stack(self.seg.mats).to(device=device, dtype=cd)).contiguous()

    def _ensure_plan(self, real_dtype):
        from . import engine

        if real_dtype not in self._plans:
            prog = torch.tensor(self.seg.rows, dtype=torch.int32).reshape(-1, 4)
            plan = engine.Plan(prog, self.num_qubits, engine.C128 if real_dtype == torch.float64 else engine.C64, self._opts)
            self._plans[real_dtype] = (plan, plan.step_types())
        self.plan, self.step_types = self._plans[real_dtype]

    def forward(self, dtype: torch.dtype = torch.float32, **kwargs):
        from . import engine, qcircuit

        dev = engine.require_cuda()
        self._ensure_plan(dtype)
        shared = qcircuit._gather_weights(self.seg, dev, dtype).contiguous()
        cols, B = [], 1
        for name, col, *scale in self.seg.batch_cols:
            v = torch.as_tensor(kwargs[name])
            v = (v[..., col] if col >= 0 else v).reshape(-1)
            if scale:
                v = v * scale[0]
            B = max(B, v.shape[0])
            cols.append(v)
        batch = (torch.stack([c.expand(B) for c in cols], dim=1).to(device=dev, dtype=dtype).contiguous()
                 if cols else torch.zeros(0, device=dev, dtype=dtype))
        out = _ShardedFunction.apply(self, shared, batch, B)
        return out.squeeze()
