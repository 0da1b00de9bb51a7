"""Circuit container: same surface as reference src/qandle/qcircuit.py, new execution model.

The reference runs a python loop over built gates, one dense matmul each (qcircuit.py:163-174), and lets
torch's tape do the backward.  Here the layer list is lowered ONCE to the engine's gate program
(``lower_modules``); per call the host only gathers the angles -- ``remapping(theta)`` for weights, the raw
named inputs (operators.py:266-271) -- with ordinary differentiable torch ops and makes ONE custom-op call
(forward) whose autograd node makes ONE call (adjoint backward).  Output-shape rules follow the reference.

Circuit splitting (qcircuit.py:215-314, splitter/) is out of scope: amplitude sharding replaces it
(qandle_b200/distributed.py).
"""
from __future__ import annotations

import contextlib
import typing
import warnings
import weakref

import torch

from . import config, embeddings, engine, measurements, operators

__all__ = ["Circuit", "UnsplittedCircuit"]


# ---------------------------------------------------------------------------------------------------------
# lowering: nn.Module list -> segments of engine gate programs
class _Segment:
    """One engine call: optional state reset, a gate program, optional trailing measurement."""

    def __init__(self):
        self.init = "inherit"  # 'inherit' (incoming state), 'zero' (AngleEmbedding), ('amp', module)
        self.rows: typing.List[typing.Tuple[int, int, int, int]] = []
        self.weight_mods: typing.List[operators.BuiltParametrizedOperator] = []  # per-gate Parameters, in slot order
        # shared-angle sources in slot order: (module, attribute name, number of slots, remapping)
        self.weight_srcs: typing.List[typing.Tuple[torch.nn.Module, str, int, typing.Callable]] = []
        self.n_slots = 0
        # batch column j -> (input name, column or -1) or (input name, column or -1, scale) when the input enters scaled
        self.batch_cols: typing.List[typing.Tuple] = []
        self.mats: typing.List[torch.Tensor] = []
        self.measure = engine.MEASURE_STATE
        self.post = None  # optional torch epilogue on the measured values (MeasureExpectation: 2 p - 1)
        self.foreign = None  # a non-engine nn.Module applied after this segment
        self.plans: typing.Dict[typing.Tuple, engine.Plan] = {}
        self._remap_groups = None
        self._whole = False  # not yet determined
        self._mats_cache: typing.Dict[typing.Tuple, torch.Tensor] = {}

    def __getstate__(self):
        d = dict(self.__dict__)
        d["plans"] = {}  # native handles are never copied / pickled; plans are rebuilt lazily
        d["_mats_cache"] = {}
        return d

    def whole_input(self):
        """(name, d) when the per-sample angle columns are exactly columns 0 .. d-1 of ONE input, in order, unscaled."""
        if self._whole is False:
            bc = self.batch_cols
            ok = bool(bc) and all(len(c) == 2 and c[0] == bc[0][0] and c[1] == k for k, c in enumerate(bc))
            self._whole = (bc[0][0], len(bc)) if ok else None
        return self._whole

    def fixed_mats(self, dev, real_dtype) -> torch.Tensor:
        """[F][2][2][2] real view of the U gates' matrices on `dev` (cached: they are constants of the segment)."""
        key = (dev, real_dtype)
        m = self._mats_cache.get(key)
        if m is None:
            if self.mats:
                cd = torch.complex128 if real_dtype == torch.float64 else torch.complex64
                m = torch.view_as_real(torch.stack(self.mats).to(device=dev, dtype=cd)).contiguous()
            else:
                m = _empty(dev, real_dtype)
            self._mats_cache[key] = m
        return m

    def batch_col(self, name: str, col: int, scale: float = 1.0) -> int:
        key = (name, col) if scale == 1.0 else (name, col, scale)
        if key not in self.batch_cols:
            self.batch_cols.append(key)
        return self.batch_cols.index(key)


_EMPTY: typing.Dict[typing.Tuple, torch.Tensor] = {}


def _empty(dev, dtype) -> torch.Tensor:
    """A shared zero-length tensor per (device, dtype): the engine call takes tensors for absent angle / matrix arguments."""
    t = _EMPTY.get((dev, dtype))
    if t is None:
        t = _EMPTY[(dev, dtype)] = torch.zeros(0, device=dev, dtype=dtype)
    return t


def _flatten(mods) -> typing.List[torch.nn.Module]:
    out = []
    for m in mods:
        if isinstance(m, Circuit):
            out.extend(_flatten(m.circuit.layers))
        elif isinstance(m, UnsplittedCircuit):
            out.extend(_flatten(m.layers))
        elif hasattr(m, "engine_children"):  # ansatz: its built gates in application order
            out.extend(_flatten(m.engine_children()))
        elif hasattr(m, "mods") and isinstance(getattr(m, "mods"), torch.nn.Sequential):
            out.extend(_flatten(m.mods))
        else:
            out.append(m)
    return out


def lower_modules(mods, num_qubits: int) -> typing.List[_Segment]:
    segs = [_Segment()]
    flat = _flatten(mods)
    for m in flat:
        seg = segs[-1]
        if seg.measure != engine.MEASURE_STATE or seg.foreign is not None:
            seg = _Segment()
            segs.append(seg)
        if isinstance(m, operators.BuiltParametrizedOperator):
            if m.named:
                seg.rows.append((m.engine_opcode | engine.FLAG_BATCH, m.qubit, -1, seg.batch_col(m.name, -1)))
            else:
                seg.rows.append((m.engine_opcode, m.qubit, -1, seg.n_slots))
                seg.weight_mods.append(m)
                seg.weight_srcs.append((m, "theta", 1, m.remapping))
                seg.n_slots += 1
        elif hasattr(m, "engine_lower_into"):
            m.engine_lower_into(seg)  # Controlled: rows, halved angle sources / scaled input columns, fixed matrices
        elif hasattr(m, "engine_lower"):
            rows, srcs, mats = m.engine_lower(seg.n_slots, len(seg.mats))
            seg.rows.extend(rows)
            for src in srcs:
                seg.weight_srcs.append(src)
                seg.n_slots += src[2]
            seg.mats.extend(mats)
        elif hasattr(m, "engine_lower_packed"):
            # ansatz with ONE packed weight tensor: rows reference consecutive slots of it
            rows, n_new = m.engine_lower_packed(seg.n_slots)
            seg.rows.extend(rows)
            seg.weight_srcs.append((m, "q_params", n_new, m.remapping))
            seg.n_slots += n_new
        elif isinstance(m, operators.BuiltU) and m.engine_unitary:
            seg.rows.append((engine.OP_U, m.qubit, -1, len(seg.mats)))
            seg.mats.append(m.engine_matrix)
        elif isinstance(m, (operators.BuiltCNOT, operators.BuiltCZ)):
            seg.rows.append((m.engine_opcode, m.c, m.t, 0))
        elif isinstance(m, operators.BuiltSWAP):
            seg.rows.append((m.engine_opcode, m.a, m.b, 0))
        elif isinstance(m, embeddings.AngleEmbeddingBuilt):
            # ignores the incoming state (reference embeddings.py:148-154): everything before it is dead
            seg = _Segment()
            segs[-1] = seg
            seg.init = "zero"
            for k, q in enumerate(m.qubits):
                seg.rows.append((m.engine_opcode | engine.FLAG_BATCH, q, -1, seg.batch_col(m.name, k)))
        elif isinstance(m, embeddings.AmplitudeEmbeddingBuilt):
            seg = _Segment()
            segs[-1] = seg
            seg.init = ("amp", m)
        elif isinstance(m, measurements.BuiltMeasurement):
            seg.measure = m.engine_measure
            if m.engine_measure == engine.MEASURE_STATE:
                continue  # identity
        elif isinstance(m, operators.UnbuiltOperator):
            raise TypeError(f"{m} is not built; call .build(num_qubits) or put it in a Circuit")
        else:
            seg.foreign = m  # any other nn.Module: run it in torch on the state, like the reference's loop
    return segs


def _dtype_code(real_dtype: torch.dtype) -> int:
    return engine.C128 if real_dtype == torch.float64 else engine.C64


def _plan_for(seg: _Segment, num_qubits: int, real_dtype: torch.dtype, dev: torch.device) -> engine.Plan:
    """The segment's plan for (state size, dtype, device): a plan owns device-resident tables and per-device kernel attributes, and a
    size-agnostic measurement instance may meet states of different sizes (reference measurements.py:30-33)."""
    final_layout = 1 if seg.measure == engine.MEASURE_PROBS else 0
    key = (num_qubits, dev.index, real_dtype, config.ENGINE_TILE_BITS, config.ENGINE_LOW_BITS, bool(config.ENGINE_FUSE), bool(config.ENGINE_STAGED), bool(config.ENGINE_PACKED), bool(config.ENGINE_FLAT), config.ENGINE_MAX_OPS_PER_SWEEP, final_layout,
           bool(config.ENGINE_SWEEP_SEARCH))
    plan = seg.plans.get(key)
    if plan is None:
        prog = torch.tensor(seg.rows, dtype=torch.int32).reshape(-1, 4)
        opts = (config.ENGINE_TILE_BITS, config.ENGINE_LOW_BITS, 0 if config.ENGINE_FUSE else -1, 0, 0, 0, final_layout, config.ENGINE_MAX_OPS_PER_SWEEP,
                0 if config.ENGINE_STAGED else -1, 0 if config.ENGINE_PACKED else -1, 0 if config.ENGINE_FLAT else -1, 0, 0,
                0 if config.ENGINE_SWEEP_SEARCH else -1)
        plan = engine.Plan(prog, num_qubits, _dtype_code(real_dtype), opts)
        seg.plans[key] = plan
    return plan


def _gather_weights(seg: _Segment, device, real_dtype) -> torch.Tensor:
    """shared_angles[i] = remapping(theta_i) (reference operators.py:271).  Sources are grouped by remapping callable
    so the remap runs once per group instead of once per gate; packed ansatz weights enter as whole tensors."""
    if not seg.weight_srcs:
        return _empty(device, real_dtype)
    if seg._remap_groups is None:
        groups: typing.Dict[int, typing.Tuple[typing.Callable, typing.List[int]]] = {}
        starts, off = [], 0
        for i, (_m, _a, k, fn) in enumerate(seg.weight_srcs):
            groups.setdefault(id(fn), (fn, []))[1].append(i)
            starts.append(off)
            off += k
        order = []
        for _fn, idxs in groups.values():
            for i in idxs:
                order.extend(range(starts[i], starts[i] + seg.weight_srcs[i][2]))
        inv = torch.empty(len(order), dtype=torch.long)
        inv[torch.tensor(order)] = torch.arange(len(order))
        seg._remap_groups = ([(fn, idxs) for fn, idxs in groups.values()], inv, order == list(range(len(order))))
    groups, inv, identity = seg._remap_groups
    parts = []
    for fn, idxs in groups:
        th = torch.cat([getattr(seg.weight_srcs[i][0], seg.weight_srcs[i][1]).reshape(-1) for i in idxs])
        parts.append(fn(th))
    ang = parts[0] if len(parts) == 1 else torch.cat(parts)
    if not identity:
        ang = ang[inv.to(ang.device)]
    return ang.to(device=device, dtype=real_dtype)


def _engine_device(owner, state, kwargs) -> torch.device:
    """The CUDA device the call runs on: where the inputs live, else where the parameters live, else the current device
    (host tensors are copied there and the result is copied back)."""
    cur = engine.require_cuda()
    if cur.type != "cuda":  # a test backend injected in place of the engine
        return cur
    for t in [state, *kwargs.values()]:
        if torch.is_tensor(t) and t.is_cuda:
            return t.device
    for p in owner.parameters():
        if p.is_cuda:
            return p.device
    return cur


def _run_segment(seg: _Segment, num_qubits: int, state, kwargs, batched_flag: typing.List[bool], dev: torch.device):
    """state: None | (N,) | (B,N) complex on any device.  Returns the segment's raw output on the engine device `dev` (the caller
    has made it the current device)."""
    N = 2**num_qubits
    # ---- initial state ----------------------------------------------------------------------------------
    init = None
    if seg.init == "inherit":
        if state is not None:
            init = state
    elif seg.init != "zero":
        x = kwargs[seg.init[1].name]
        init = seg.init[1].embed(x)
    if init is not None:
        if not torch.is_tensor(init):
            init = torch.as_tensor(init)
        if not init.is_complex():
            init = torch.complex(init, torch.zeros_like(init))
        if init.dim() == 2:
            batched_flag[0] = True
        elif init.dim() != 1:
            lead = init.shape[:-1]
            init = init.reshape(-1, init.shape[-1])
            batched_flag[0] = True
            batched_flag.append(lead)
        if init.shape[-1] != N:
            raise RuntimeError(f"state has {init.shape[-1]} amplitudes, expected 2**{num_qubits}")
    # ---- dtype: complex128 only if the state or an input asks for it (reference is complex64 only, Q4) ------
    real_dtype = torch.float32
    if init is not None and init.dtype == torch.complex128:
        real_dtype = torch.float64
    # ---- per-sample angles -------------------------------------------------------------------------------
    B = init.shape[0] if (init is not None and init.dim() == 2) else 1
    cols = []
    whole = seg.whole_input()
    if whole is not None and torch.is_tensor(kwargs.get(whole[0])) and kwargs[whole[0]].dim() >= 1 and kwargs[whole[0]].shape[-1] == whole[1]:
        # the per-sample angle matrix IS one input tensor (an AngleEmbedding over all its columns, in order): no per-column
        # select / stack on the way in, no scatter of the gradient on the way back
        v = kwargs[whole[0]]
        if v.dtype == torch.float64 and init is None:
            real_dtype = torch.float64
        if v.dim() == 1:
            v = v.reshape(1, -1)
        else:
            if v.dim() > 2:
                if len(batched_flag) == 1:
                    batched_flag.append(v.shape[:-1])
                v = v.reshape(-1, v.shape[-1])
            batched_flag[0] = True
        if B != 1 and v.shape[0] != 1 and v.shape[0] != B:
            raise RuntimeError(f"batch size mismatch: {v.shape[0]} vs {B}")
        B = max(B, v.shape[0])
        cols = None
        batch = (v.expand(B, -1) if v.shape[0] != B else v).to(device=dev, dtype=real_dtype).contiguous()
    for name, col, *scale in (seg.batch_cols if cols is not None else ()):
        v = kwargs[name]
        if not torch.is_tensor(v):
            v = torch.as_tensor(v, dtype=torch.float32)
        if scale:
            v = v * scale[0]
        if v.dtype == torch.float64 and init is None:
            real_dtype = torch.float64
        if col >= 0:  # AngleEmbedding input (…, d)
            if v.dim() == 1:
                v = v[col].reshape(1)
            else:
                if v.dim() > 2:
                    if len(batched_flag) == 1:
                        batched_flag.append(v.shape[:-1])
                    v = v.reshape(-1, v.shape[-1])
                v = v[:, col]
                batched_flag[0] = True
        else:  # named gate: 0-dim shared or 1-dim per-sample (reference operators.py:266-269)
            if v.dim() == 0:
                v = v.reshape(1)
            elif v.dim() == 1:
                batched_flag[0] = True
            else:
                raise RuntimeError("named gate inputs must be 0- or 1-dimensional (reference operators.py:268)")
        cols.append(v)
    for v in cols or ():
        if v.shape[0] != 1:
            if B != 1 and v.shape[0] != B:
                raise RuntimeError(f"batch size mismatch: {v.shape[0]} vs {B}")
            B = v.shape[0]
    if cols:
        batch = torch.stack([c.expand(B) if c.shape[0] == 1 else c for c in cols], dim=1).to(device=dev, dtype=real_dtype).contiguous()
    elif cols is not None:
        batch = _empty(dev, real_dtype)
    shared = _gather_weights(seg, dev, real_dtype).contiguous()
    cdtype = torch.complex128 if real_dtype == torch.float64 else torch.complex64
    mats = seg.fixed_mats(dev, real_dtype)
    if init is not None:
        init = init.to(device=dev, dtype=cdtype)
        if init.dim() == 1:
            init = init.unsqueeze(0)
        if init.shape[0] != B:
            init = init.expand(B, -1)  # unbatched state + batched named input (quirk Q7)
        init = init.contiguous()
    plan = _plan_for(seg, num_qubits, real_dtype, dev)
    return engine.run_circuit(plan, shared, batch, mats, init, B, seg.measure)


# Lowered segments per owner module.  Kept OUT of the module's __dict__ (weak keys): the cache holds native plan handles and
# closures, so it must not travel with copy.deepcopy / pickle / torch.save(model) and must die with the module.
_SEGMENTS: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def _segments_for(owner, mods, num_qubits: int) -> typing.List[_Segment]:
    sig = (num_qubits, tuple(id(m) for m in mods))  # a layer list edited after the first call is lowered again
    hit = _SEGMENTS.get(owner)
    if hit is None or hit[0] != sig:
        hit = (sig, lower_modules(mods, num_qubits))
        _SEGMENTS[owner] = hit
    return hit[1]


def segments_of(owner) -> typing.List[_Segment]:
    """The lowered segments of a module that has run at least once (introspection: bench.py, tools/)."""
    hit = _SEGMENTS.get(owner)
    if hit is None:
        raise RuntimeError("the module has not been run yet: nothing is lowered")
    return hit[1]


def run_modules(owner, mods, num_qubits: int, state, kwargs):
    """Apply built modules to `state` through the engine; shared by Circuit, single gates and measurements."""
    segs = _segments_for(owner, mods, num_qubits)
    origin = None
    for t in [state, *kwargs.values()]:
        if torch.is_tensor(t):
            origin = t.device
            break
    if origin is None:
        for p in owner.parameters():
            origin = p.device
            break
    dev = _engine_device(owner, state, kwargs)
    batched = [False]
    out = state
    measure = engine.MEASURE_STATE
    with (torch.cuda.device(dev) if dev.type == "cuda" else contextlib.nullcontext()):
        for seg in segs:
            no_work = not seg.rows and seg.init == "inherit" and seg.measure == engine.MEASURE_STATE
            if not no_work:
                out = _run_segment(seg, num_qubits, out, kwargs, batched, dev)
                if seg.post is not None:
                    out = seg.post(out)
                measure = seg.measure
            if seg.foreign is not None:
                if out is None:
                    out = torch.zeros(2**num_qubits, dtype=torch.complex64)
                    out[0] = 1
                elif measure == engine.MEASURE_PROBS:
                    out = out.squeeze()  # the measurement's own .squeeze() (measurements.py:123) comes before the next module
                elif not no_work and measure == engine.MEASURE_STATE and out.dim() == 2 and not batched[0]:
                    out = out.squeeze(0)  # the engine works on [B, N]; an unbatched state stays unbatched for the torch module
                out = seg.foreign(out, **kwargs) if getattr(seg.foreign, "named", False) else seg.foreign(out)
                measure = None
    if out is None:  # empty circuit on the default state
        out = torch.zeros(2**num_qubits, dtype=torch.complex64)
        out[0] = 1
        return out
    # ---- the reference's output-shape rules ------------------------------------------------------------------
    if measure == engine.MEASURE_PROBS:
        out = out.squeeze()  # measurements.py:123 (quirk Q5)
    elif measure is not None and torch.is_tensor(out) and out.dim() == 2:
        if not batched[0]:
            out = out.squeeze(0)
        elif len(batched) > 1:
            out = out.reshape(*batched[1], out.shape[-1])
    if origin is not None and torch.is_tensor(out) and out.device != origin:
        out = out.to(origin)
    return out


# ---------------------------------------------------------------------------------------------------------
class Circuit(torch.nn.Module):
    """Drop-in for reference qcircuit.py:13-134 (``split_max_qubits`` is accepted and ignored with a warning:
    amplitude sharding replaces circuit splitting)."""

    def __init__(self, layers, num_qubits=None, split_max_qubits=0, circuit=None):
        super().__init__()
        self.name = ""
        self.named = True
        if not hasattr(layers, "__iter__"):
            layers = [layers]
        layers = list(layers)
        if num_qubits is not None:
            self.num_qubits = num_qubits
        else:
            self.num_qubits = len(self._used_qubits(layers))
        if circuit is not None:
            self.circuit = circuit
        else:
            if split_max_qubits > 0 and self.num_qubits > split_max_qubits:
                warnings.warn("qandle_b200 does not split circuits: the engine runs the full state on the GPU "
                              "(use qandle_b200.distributed for states larger than one GPU)")
            self.circuit = UnsplittedCircuit(self.num_qubits, layers)

    def forward(self, state=None, **kwargs):
        return self.circuit.forward(state, **kwargs)

    def to_matrix(self, **kwargs):
        return self.circuit.to_matrix(**kwargs)

    def to_matrix_engine(self, dtype=torch.complex64, **kwargs):
        return self.circuit.to_matrix_engine(dtype=dtype, **kwargs)

    def sample(self, shots: int, state=None, generator=None, **kwargs):
        return self.circuit.sample(shots, state, generator=generator, **kwargs)

    def to_qasm(self):
        return self.circuit.to_qasm()

    def to_openqasm2(self) -> str:
        from . import qasm

        return qasm.convert_to_qasm(self, qasm_version=2, include_header=True)

    def to_openqasm3(self) -> str:
        from . import qasm

        return qasm.convert_to_qasm(self, qasm_version=3, include_header=True)

    def __matmul__(self, x):
        return self.circuit.forward(x)

    @staticmethod
    def _used_qubits(layers) -> typing.Set[int]:
        """reference qcircuit.py:69-106."""
        qubits: typing.Set[int] = set()
        for layer in layers:
            if isinstance(layer, (operators.CNOT, operators.CZ, operators.BuiltCNOT, operators.BuiltCZ)):
                qubits.add(layer.c)
                qubits.add(layer.t)
            elif isinstance(layer, (operators.SWAP, operators.BuiltSWAP)):
                qubits.add(layer.a)
                qubits.add(layer.b)
            elif hasattr(layer, "qubit"):
                qubits.add(layer.qubit)
            elif hasattr(layer, "num_qubits") and layer.num_qubits is not None:
                qubits.update(range(layer.num_qubits))
            elif hasattr(layer, "qubits") and layer.qubits is not None:
                qubits.update(layer.qubits)
            elif isinstance(layer, (measurements.UnbuiltMeasurement, measurements.BuiltMeasurement)):
                pass
            else:
                raise ValueError(
                    f"Unknown layer type {type(layer)}, number of qubits could not be inferred. Pass :code:`num_qubits` to the circuit.")
        if len(qubits) == 0:
            raise ValueError("Number of qubits could not be inferred from layers. Please provide num_qubits to the circuit directly.")
        return qubits

    def decompose(self):
        return Circuit(layers=[], num_qubits=self.num_qubits, circuit=self.circuit.decompose())

    def split(self, max_qubits):
        warnings.warn("qandle_b200 does not split circuits; returning the circuit unchanged")
        return self


class UnsplittedCircuit(torch.nn.Module):
    """reference qcircuit.py:137-212."""

    def __init__(self, num_qubits: int, layers: list):
        super().__init__()
        self.num_qubits = num_qubits
        self.layers = torch.nn.ModuleList(self._build_layers(layers, num_qubits))

    @staticmethod
    def _build_layers(layers: list, num_qubits: int) -> typing.List[torch.nn.Module]:
        return [la.build(num_qubits=num_qubits) if hasattr(la, "build") else la for la in layers]

    @property
    def state(self) -> torch.Tensor:
        """|0...0> complex64 (reference qcircuit.py:148-149)."""
        s = torch.zeros(2**self.num_qubits, dtype=torch.complex64)
        s[0] = 1
        return s

    def forward(self, state=None, **kwargs):
        """Run the circuit.  ``state=None`` starts from |0...0>; named inputs go to the gates / embeddings that carry
        that name (reference qcircuit.py:163-174)."""
        return run_modules(self, self.layers, self.num_qubits, state, kwargs)

    def decompose(self) -> "UnsplittedCircuit":
        new_layers = []
        for layer in self.layers:
            if isinstance(layer, Circuit):
                new_layers.extend(layer.decompose().circuit.layers)
            elif hasattr(layer, "decompose") and not isinstance(layer, (measurements.BuiltMeasurement, embeddings.InputOperatorBuilt)):
                new_layers.extend(layer.decompose())
            else:
                new_layers.append(layer)
        return UnsplittedCircuit(self.num_qubits, new_layers)

    def to_matrix(self, **kwargs):
        return operators.chain_matrices(self, self.layers, **kwargs)

    def _unitary_layers(self) -> typing.List[torch.nn.Module]:
        """The layers that act linearly on the state: everything except measurements; embeddings are rejected."""
        mods = []
        for layer in self.layers:
            if isinstance(layer, measurements.BuiltMeasurement):
                continue
            if isinstance(layer, embeddings.InputOperatorBuilt):
                raise ValueError("Input operators do not have a matrix representation")
            mods.append(layer)
        return mods

    def to_matrix_engine(self, dtype=torch.complex64, **kwargs) -> torch.Tensor:
        """The circuit's matrix M (row-vector convention of ``to_matrix``: ``forward(s) == s @ M``, reference
        operators.py:67-69) evaluated ON THE ENGINE: the identity is pushed through the sweep kernels as a batch of 2^n
        basis states, so M costs O(G 4^n) amplitude updates and 4^n amplitudes of memory instead of the dense helper's
        chain of G products of 4^n-entry matrices.  Named inputs must be 0-dim (one matrix)."""
        owner = _ModuleGroup(self._unitary_layers())
        N = 2**self.num_qubits
        dev = next((p.device for p in self.parameters()), torch.device("cpu"))
        eye = torch.eye(N, dtype=dtype, device=dev)
        out = run_modules(owner, owner.mods, self.num_qubits, eye, kwargs)
        return out.reshape(N, N)

    def sample(self, shots: int, state=None, generator=None, **kwargs) -> torch.Tensor:
        """Draw ``shots`` computational-basis outcomes per batch element from |psi|^2 of the state in front of the
        trailing measurement (qubit 0 is the most significant bit of the returned integers, reference operators.py:549).
        The joint distribution comes from the engine (``MeasureJointProbability`` reduction); the draw is
        ``torch.multinomial`` where the inputs live (at most 2^24 outcomes, i.e. 24 qubits).  Not differentiable.  Returns int64 ``(shots,)`` or ``(B, shots)``."""
        mods = [la for la in self.layers if not isinstance(la, measurements.BuiltMeasurement)]
        owner = _ModuleGroup(mods + [measurements.MeasureJointProbability()])
        with torch.no_grad():
            p = run_modules(owner, owner.mods, self.num_qubits, state, kwargs)
            flat = p.reshape(-1, p.shape[-1]).clamp_min(0)
            idx = torch.multinomial(flat, int(shots), replacement=True, generator=generator)
        return idx.reshape(*p.shape[:-1], int(shots))

    def to_qasm(self):
        """Flat list of QasmRepresentation (reference qcircuit.py:176-181); composite layers without their own
        representation are exported through ``decompose()``."""
        reps = []

        def emit(layer):
            if isinstance(layer, Circuit):
                reps.extend(layer.circuit.to_qasm())
                return
            try:
                r = layer.to_qasm()
            except NotImplementedError:
                for d in layer.decompose():
                    emit(d)
                return
            reps.extend(r if isinstance(r, (list, tuple)) else [r])

        for layer in self.layers:
            emit(layer)  # (measurements: `measure q[w] -> c[w]` lines, reference measurements.py:16-22)
        return reps


class _ModuleGroup(torch.nn.Module):
    """Throw-away owner for a sub-list of built layers (its lowered segments are cached on it, not on the circuit)."""

    def __init__(self, mods):
        super().__init__()
        self.mods = list(mods)  # plain list: the layers stay owned by their circuit

    def parameters(self, recurse: bool = True):
        for m in self.mods:
            yield from m.parameters(recurse)
