"""OpenQASM text export and import for engine circuits.

Export keeps the reference's surface (reference src/qandle/qasm.py:34-70 ``convert_to_qasm``; ``Circuit.to_openqasm2`` /
``to_openqasm3``, reference qcircuit.py:127-134): one statement per gate, named gates become ``input float`` declarations
in OpenQASM 3.  The reference has NO importer; ``circuit_from_qasm`` is the counterpart for it: it reads the OpenQASM 2
subset that the exporter (and most transpilers) emit -- one quantum register, ``rx/ry/rz/p/u1/u2/u3/u``, the fixed one-qubit
gates of qelib1.inc, ``cx/cz/swap/ccx`` -- into the engine's operators, so a circuit written elsewhere runs on the B200
kernels.  Host-only code: no amplitudes are touched here.
"""
from __future__ import annotations

import ast
import cmath
import math
import re
import typing

import torch

from . import operators as op

__all__ = ["convert_to_qasm", "circuit_from_qasm", "QasmSyntaxError"]


class QasmSyntaxError(ValueError):
    pass


# ---------------------------------------------------------------------------------------------------------
# export
def _flatten_reps(reps) -> typing.Iterator[op.QasmRepresentation]:
    if isinstance(reps, op.QasmRepresentation):
        yield reps
        return
    for r in reps:
        yield from _flatten_reps(r)


def _statement(rep: op.QasmRepresentation) -> str:
    s = rep.gate_str
    if rep.qasm3_inputs:
        s += f"({rep.qasm3_inputs})"
    elif rep.gate_value is not None and rep.gate_value != "":
        s += f"({rep.gate_value})"
    if rep.qubit is not None and rep.qubit != "":
        s += f" q[{rep.qubit}]"
    return s + ";"


def convert_to_qasm(circuit, qasm_version: int = 2, include_header: bool = True) -> str:
    """Text of ``circuit`` in OpenQASM 2 or 3 (same statement forms as the reference's exporter)."""
    n = circuit.num_qubits
    reps = list(_flatten_reps(circuit.to_qasm()))
    lines = []
    if include_header:
        if qasm_version == 2:
            lines += ["OPENQASM 2.0;", 'include "qelib1.inc";', f"qreg q[{n}];", f"creg c[{n}];"]
        else:
            lines += ["OPENQASM 3.0;", f"qubit[{n}] q;", f"bit[{n}] c;"]
    if qasm_version == 3:
        for r in reps:
            if r.qasm3_inputs:
                lines.append(f"input float {r.qasm3_inputs};")
            if r.qasm3_outputs:
                lines.append(f"output float {r.qasm3_outputs};")
    lines += [_statement(r) for r in reps]
    return "\n".join(lines) + "\n"


# ---------------------------------------------------------------------------------------------------------
# import
_SAFE_FUNCS = {"pi": math.pi, "sin": math.sin, "cos": math.cos, "tan": math.tan, "exp": math.exp, "ln": math.log, "sqrt": math.sqrt}
def _eval_angle(expr: str) -> float:
    """Constant angle expression of OpenQASM 2 (numbers, pi, + - * / ^, the unary functions of the spec), evaluated over FLOATS by
    walking the parsed expression: no eval(), no integer towers (`9^9^9^9` in an untrusted file is an overflow error, not a hang)."""
    if len(expr) > 256:
        raise QasmSyntaxError("angle expression too long")
    try:
        tree = ast.parse(expr.strip().replace("^", "**"), mode="eval")
    except SyntaxError as exc:
        raise QasmSyntaxError(f"unsupported angle expression {expr!r}") from exc

    def ev(node) -> float:
        if isinstance(node, ast.Expression):
            return ev(node.body)
        if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)) and not isinstance(node.value, bool):
            return float(node.value)
        if isinstance(node, ast.Name) and node.id == "pi":
            return math.pi
        if isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.UAdd, ast.USub)):
            v = ev(node.operand)
            return -v if isinstance(node.op, ast.USub) else v
        if isinstance(node, ast.BinOp) and isinstance(node.op, (ast.Add, ast.Sub, ast.Mult, ast.Div, ast.Pow)):
            a, b = ev(node.left), ev(node.right)
            if isinstance(node.op, ast.Add):
                return a + b
            if isinstance(node.op, ast.Sub):
                return a - b
            if isinstance(node.op, ast.Mult):
                return a * b
            if isinstance(node.op, ast.Div):
                return a / b
            return math.pow(a, b)
        if (isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id in _SAFE_FUNCS and node.func.id != "pi"
                and len(node.args) == 1 and not node.keywords):
            return float(_SAFE_FUNCS[node.func.id](ev(node.args[0])))
        raise QasmSyntaxError(f"unsupported angle expression {expr!r}")

    try:
        v = ev(tree)
    except QasmSyntaxError:
        raise
    except Exception as exc:  # noqa: BLE001  (overflow, division by zero, domain errors)
        raise QasmSyntaxError(f"cannot evaluate angle expression {expr!r}: {exc}") from exc
    if not math.isfinite(v):
        raise QasmSyntaxError(f"angle expression {expr!r} is not finite")
    return v


def _u3(theta: float, phi: float, lam: float) -> torch.Tensor:
    c, s = math.cos(theta / 2), math.sin(theta / 2)
    return torch.tensor([[c, -cmath.exp(1j * lam) * s], [cmath.exp(1j * phi) * s, cmath.exp(1j * (phi + lam)) * c]], dtype=torch.complex64)


_R2 = 2**-0.5
_FIXED = {
    "id": [[1, 0], [0, 1]], "x": [[0, 1], [1, 0]], "y": [[0, -1j], [1j, 0]], "z": [[1, 0], [0, -1]],
    "h": [[_R2, _R2], [_R2, -_R2]], "s": [[1, 0], [0, 1j]], "sdg": [[1, 0], [0, -1j]],
    "t": [[1, 0], [0, cmath.exp(0.25j * math.pi)]], "tdg": [[1, 0], [0, cmath.exp(-0.25j * math.pi)]],
    "sx": [[0.5 + 0.5j, 0.5 - 0.5j], [0.5 - 0.5j, 0.5 + 0.5j]],
}


def _fixed_gate(qubit: int, m) -> op.U:
    # U applies the TRANSPOSE of its matrix (reference operators.py:103-104/125-126, quirk Q2): hand it m^T so the
    # gate acts as m on column vectors.
    return op.U(qubit, torch.as_tensor(m, dtype=torch.complex64).T.contiguous())


_STMT = re.compile(r"^([A-Za-z_][A-Za-z0-9_]*)\s*(?:\((.*)\))?\s*(.*)$", re.S)
_QARG = re.compile(r"^([A-Za-z_][A-Za-z0-9_]*)\s*\[\s*(\d+)\s*\]$")


def _split_args(s: str) -> typing.List[str]:
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def parse_qasm(text: str) -> typing.Tuple[int, typing.List[op.UnbuiltOperator]]:
    """(number of qubits, engine operators) of an OpenQASM 2 program.  Registers are laid out in declaration order;
    ``measure`` / ``barrier`` / ``creg`` are ignored (measurement is chosen by the caller), ``reset`` maps to ``Reset``."""
    text = re.sub(r"//[^\n]*", "", text)
    offsets: typing.Dict[str, typing.Tuple[int, int]] = {}
    n = 0
    layers: typing.List[op.UnbuiltOperator] = []

    def qubit(tok: str) -> int:
        m = _QARG.match(tok.strip())
        if not m or m.group(1) not in offsets:
            raise QasmSyntaxError(f"bad qubit argument {tok!r} (whole-register broadcasts are not supported)")
        off, size = offsets[m.group(1)]
        k = int(m.group(2))
        if k >= size:
            raise QasmSyntaxError(f"qubit index out of range in {tok!r}")
        return off + k

    for raw in text.split(";"):
        st = raw.strip()
        if not st:
            continue
        low = st.lower()
        if low.startswith("openqasm"):
            if not re.match(r"openqasm\s+2(\.\d+)?$", low):
                raise QasmSyntaxError(f"only OpenQASM 2 is imported, got {st!r}")
            continue
        if low.startswith("include") or low.startswith("creg") or low.startswith("barrier") or low.startswith("measure"):
            continue
        if low.startswith("qreg"):
            m = re.match(r"qreg\s+([A-Za-z_][A-Za-z0-9_]*)\s*\[\s*(\d+)\s*\]$", st)
            if not m:
                raise QasmSyntaxError(f"bad register declaration {st!r}")
            offsets[m.group(1)] = (n, int(m.group(2)))
            n += int(m.group(2))
            continue
        if low.startswith("gate ") or low.startswith("opaque ") or low.startswith("if"):
            raise QasmSyntaxError(f"unsupported statement {st.split()[0]!r} (gate definitions and classical control are not imported)")
        m = _STMT.match(st)
        if not m:
            raise QasmSyntaxError(f"cannot parse {st!r}")
        name, params, qargs = m.group(1).lower(), m.group(2), m.group(3)
        ang = [_eval_angle(a) for a in _split_args(params)] if params is not None else []
        qs = [qubit(t) for t in _split_args(qargs)]

        def need(n_ang, n_q):
            if len(ang) != n_ang or len(qs) != n_q:
                raise QasmSyntaxError(f"{name} takes {n_ang} parameter(s) and {n_q} qubit(s): {st!r}")

        if name in ("rx", "ry", "rz"):
            need(1, 1)
            cls = {"rx": op.RX, "ry": op.RY, "rz": op.RZ}[name]
            layers.append(cls(qs[0], theta=torch.tensor(ang[0]), remapping=None))  # angles are literal: no remapping
        elif name in ("p", "u1"):
            need(1, 1)
            layers.append(_fixed_gate(qs[0], [[1, 0], [0, cmath.exp(1j * ang[0])]]))
        elif name == "u2":
            need(2, 1)
            layers.append(_fixed_gate(qs[0], _u3(math.pi / 2, ang[0], ang[1])))
        elif name in ("u3", "u"):
            need(3, 1)
            layers.append(_fixed_gate(qs[0], _u3(*ang)))
        elif name in _FIXED:
            need(0, 1)
            layers.append(_fixed_gate(qs[0], _FIXED[name]))
        elif name in ("cx", "cnot"):
            need(0, 2)
            layers.append(op.CNOT(qs[0], qs[1]))
        elif name == "cz":
            need(0, 2)
            layers.append(op.CZ(qs[0], qs[1]))
        elif name == "swap":
            need(0, 2)
            layers.append(op.SWAP(qs[0], qs[1]))
        elif name == "ccx":
            need(0, 3)
            layers.append(op.Controlled(qs[0], op.CNOT(qs[1], qs[2])))
        elif name == "reset":
            need(0, 1)
            layers.append(op.Reset(qs[0]))
        else:
            raise QasmSyntaxError(f"unsupported gate {name!r}")
    if n == 0:
        raise QasmSyntaxError("no qreg declaration")
    return n, layers


def circuit_from_qasm(text: str, measurement=None):
    """Build a ``Circuit`` from OpenQASM 2 text.  Rotation angles become trainable ``theta`` Parameters initialised to
    the literal values (``remapping=None``); ``measurement`` (e.g. ``MeasureProbability()``) is appended if given."""
    from . import qcircuit

    n, layers = parse_qasm(text)
    if measurement is not None:
        layers = layers + [measurement]
    return qcircuit.Circuit(layers=layers, num_qubits=n)
