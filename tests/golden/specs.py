"""Circuit specs shared by the golden generator (which instantiates them with the REAL reference,
`import qandle`) and by the parity tests (which instantiate them with `qandle_b200`).  Because the
engine mirrors the reference's constructors, one builder serves both modules.

A spec is a list of (class_name, kwargs) tuples; kwargs values are plain python / lists so the spec
can be stored as JSON next to the golden arrays.  Special kwarg values:
  {"__tensor__": [...], "dtype": "float32"|"complex64"}  -> torch.tensor
  "__none__"                                              -> None   (e.g. remapping=None)
"""
import json
import math

import torch


def _decode(v, mod=None):
    if isinstance(v, dict) and "__op__" in v:  # nested operator, e.g. Invert(target=RY(...))
        name, kwargs = v["__op__"]
        return getattr(mod, name)(**{k: _decode(x, mod) for k, x in kwargs.items()})
    if isinstance(v, dict) and "__tensor__" in v:
        if v.get("dtype") == "complex64":
            arr = torch.tensor(v["__tensor__"], dtype=torch.float32)
            return torch.complex(arr[..., 0], arr[..., 1])
        return torch.tensor(v["__tensor__"], dtype=getattr(torch, v.get("dtype", "float32")))
    if v == "__none__":
        return None
    if isinstance(v, str) and v.startswith("__cls__:"):  # a class of the module, e.g. control_gate=CZ
        return getattr(mod, v[len("__cls__:"):])
    return v


def build_layers(mod, spec):
    """Instantiate spec entries with module `mod` (qandle or qandle_b200)."""
    layers = []
    for name, kwargs in spec:
        cls = getattr(mod, name)
        kw = {k: _decode(v, mod) for k, v in kwargs.items()}
        layers.append(cls(**kw))
    return layers


def build_circuit(mod, spec, num_qubits=None):
    return mod.Circuit(layers=build_layers(mod, spec), num_qubits=num_qubits)


def T(values, dtype="float32"):
    return {"__tensor__": values, "dtype": dtype}


NONE = "__none__"


def api_specs():
    """name -> dict(spec, num_qubits, inputs {name: shape}, state: None|'batched'|'unbatched', batch)"""
    S = {}
    # BASELINE config 1 (docs/tutorial/01helloquantum.ipynb cell 5 form), default tanh remap
    c1 = [("AngleEmbedding", {"name": "x", "qubits": [0, 1, 2, 3]})]
    c1 += [("RX", {"qubit": q}) for q in range(4)]
    c1 += [("RY", {"qubit": q}) for q in range(4)]
    c1 += [("CNOT", {"control": q, "target": (q + 1) % 4}) for q in range(4)]
    c1 += [("MeasureProbability", {})]
    S["c1_tanh"] = dict(spec=c1, num_qubits=None, inputs={"x": [64, 4]}, state=None)
    c1n = [(n, dict(kw, remapping=NONE) if n in ("RX", "RY") else kw) for n, kw in c1]
    S["c1_none"] = dict(spec=c1n, num_qubits=None, inputs={"x": [64, 4]}, state=None)
    S["c1_unbatched"] = dict(spec=c1n, num_qubits=None, inputs={"x": [4]}, state=None)
    # every rotation kind with explicit thetas, CZ, SWAP, U (non-symmetric matrix: pins quirk Q2)
    # U = RZ(0.5) RY(0.7): unitary and NOT symmetric, so the reference's un-transposed application shows
    c, s, pr, pi_ = math.cos(0.35), math.sin(0.35), math.cos(0.25), math.sin(0.25)
    u = T([[[pr * c, -pi_ * c], [-pr * s, pi_ * s]], [[pr * s, pi_ * s], [pr * c, pi_ * c]]], "complex64")
    mix = [
        ("RX", {"qubit": 0, "theta": 0.3, "remapping": NONE}),
        ("RY", {"qubit": 1, "theta": -1.1, "remapping": NONE}),
        ("RZ", {"qubit": 2, "theta": 2.2, "remapping": NONE}),
        ("CNOT", {"control": 0, "target": 2}),
        ("U", {"qubit": 1, "matrix": u}),
        ("CZ", {"control": 2, "target": 1}),
        ("RY", {"qubit": 2, "name": "phi", "remapping": NONE}),
        ("SWAP", {"a": 0, "b": 2}),
        ("RX", {"qubit": 1, "theta": 0.77}),  # default (tanh) remapping
        ("RZ", {"qubit": 0, "name": "phi"}),
        ("CNOT", {"control": 2, "target": 0}),
    ]
    S["mix_state"] = dict(spec=mix + [("MeasureState", {})], num_qubits=3, inputs={"phi": [7]}, state="batched", batch=7)
    S["mix_state_scalar_named"] = dict(spec=mix + [("MeasureState", {})], num_qubits=3, inputs={"phi": []}, state="unbatched")
    S["mix_autobatch"] = dict(spec=mix + [("MeasureProbability", {})], num_qubits=3, inputs={"phi": [5]}, state="unbatched")
    S["mix_joint"] = dict(spec=mix + [("MeasureJointProbability", {})], num_qubits=3, inputs={"phi": [4]}, state="batched", batch=4)
    S["mix_nomeasure"] = dict(spec=mix, num_qubits=3, inputs={"phi": [2]}, state="batched", batch=2)
    # embeddings (test_embeddings.py shapes)
    for rot in ("rx", "ry", "rz"):
        for n in (3, 4, 5):
            sp = [("AngleEmbedding", {"name": "amp", "qubits": list(range(n)), "rotation": rot})]
            S[f"angle_{rot}{n}"] = dict(spec=sp, num_qubits=n, inputs={"amp": [n]}, state=None)
            S[f"angle_{rot}{n}_b"] = dict(spec=sp, num_qubits=n, inputs={"amp": [10, n]}, state=None)
    S["amp_norm"] = dict(
        spec=[("AmplitudeEmbedding", {"name": "amp", "qubits": [0, 1, 2, 3], "normalize": True}),
              ("RY", {"qubit": 2, "theta": 0.4, "remapping": NONE}), ("CNOT", {"control": 2, "target": 0}),
              ("MeasureProbability", {})],
        num_qubits=4, inputs={"amp": [6, 16]}, state=None)
    S["amp_pad"] = dict(
        spec=[("AmplitudeEmbedding", {"name": "amp", "qubits": [0, 1, 2, 3], "normalize": True, "pad_with": 0}),
              ("RX", {"qubit": 1, "theta": 1.4, "remapping": NONE}), ("MeasureJointProbability", {})],
        num_qubits=4, inputs={"amp": [11]}, state=None)
    # measurement shape quirks (measurements.py:123 .squeeze())
    S["probs_b1"] = dict(spec=[("RY", {"qubit": 0, "theta": 0.9, "remapping": NONE}), ("MeasureProbability", {})],
                         num_qubits=3, inputs={}, state="batched", batch=1)
    S["probs_n1"] = dict(spec=[("RY", {"qubit": 0, "theta": 0.9, "remapping": NONE}), ("MeasureProbability", {})],
                         num_qubits=1, inputs={}, state="batched", batch=3)
    S["probs_b17"] = dict(spec=[("MeasureProbability", {})], num_qubits=3, inputs={}, state="batched", batch=17)
    # StronglyEntanglingLayer (test_ansaetze.py:6-26, 63-83), tanh default + None
    S["sel4_d10"] = dict(spec=[("StronglyEntanglingLayer", {"qubits": [0, 1, 2, 3], "depth": 10, "remapping": NONE})],
                         num_qubits=4, inputs={}, state="unbatched")
    S["sel5_d7_b"] = dict(spec=[("StronglyEntanglingLayer", {"qubits": [0, 1, 2, 3, 4], "depth": 7, "remapping": NONE}),
                                ("MeasureProbability", {})],
                          num_qubits=5, inputs={}, state="batched", batch=17)
    S["sel_sub_tanh"] = dict(spec=[("StronglyEntanglingLayer", {"qubits": [0, 1, 3, 4], "depth": 3}),
                                   ("MeasureProbability", {})],
                             num_qubits=5, inputs={}, state="batched", batch=3)
    # hybrid: embedding + SEL + data re-uploading + probs at 6 qubits
    hyb = [("AngleEmbedding", {"name": "x", "qubits": list(range(6)), "rotation": "ry"}),
           ("StronglyEntanglingLayer", {"qubits": list(range(6)), "depth": 2, "remapping": NONE}),
           ("RX", {"qubit": 3, "name": "y", "remapping": NONE}),
           ("StronglyEntanglingLayer", {"qubits": list(range(6)), "depth": 2}),
           ("MeasureProbability", {})]
    S["hybrid6"] = dict(spec=hyb, num_qubits=6, inputs={"x": [9, 6], "y": [9]}, state=None)
    # Reset (non-unitary, runs in torch between engine segments) and Invert
    OP = lambda cls_, **kw: {"__op__": [cls_, kw]}
    ri = [("RX", {"qubit": 0, "theta": 0.3, "remapping": NONE}), ("RY", {"qubit": 1, "theta": 1.1}),
          ("CNOT", {"control": 0, "target": 2}),
          ("Invert", {"target": OP("RY", qubit=2, theta=0.7)}),
          ("Reset", {"qubit": 1}),
          ("Invert", {"target": OP("CNOT", control=2, target=1)}),
          ("Invert", {"target": OP("RZ", qubit=0, theta=-0.4, remapping=NONE)}),
          ("RX", {"qubit": 1, "theta": 0.9, "remapping": NONE}),
          ("MeasureProbability", {})]
    S["reset_invert"] = dict(spec=ri, num_qubits=3, inputs={}, state="batched", batch=5)
    S["reset_invert_unbatched"] = dict(spec=ri[:-1] + [("MeasureState", {})], num_qubits=3, inputs={}, state="unbatched")
    # Controlled (reference operators.py:459-515) with every supported target kind
    ctl = [("RY", {"qubit": 0, "theta": 0.8, "remapping": NONE}), ("RX", {"qubit": 1, "theta": -0.6, "remapping": NONE}),
           ("RY", {"qubit": 2, "theta": 1.9, "remapping": NONE}), ("RX", {"qubit": 3, "theta": 0.5, "remapping": NONE}),
           ("Controlled", {"control": 0, "target": OP("RY", qubit=2, theta=0.7)}),  # default (tanh) remapping
           ("Controlled", {"control": 3, "target": OP("RX", qubit=1, theta=-1.3, remapping=NONE)}),
           ("Controlled", {"control": 1, "target": OP("RZ", qubit=0, theta=2.1, remapping=NONE)}),
           ("Controlled", {"control": 2, "target": OP("U", qubit=3, matrix=u)}),
           ("Controlled", {"control": 0, "target": OP("RY", qubit=3, name="phi", remapping=NONE)}),
           ("Controlled", {"control": 2, "target": OP("CNOT", control=0, target=1)}),
           ("Controlled", {"control": 1, "target": OP("CZ", control=3, target=2)}),
           ("Controlled", {"control": 3, "target": OP("SWAP", a=0, b=2)}),
           ("Controlled", {"control": 1, "target": OP("RX", qubit=2, name="phi")})]
    S["controlled_probs"] = dict(spec=ctl + [("MeasureProbability", {})], num_qubits=4, inputs={"phi": [6]}, state="batched", batch=6)
    S["controlled_state_autobatch"] = dict(spec=ctl + [("MeasureState", {})], num_qubits=4, inputs={"phi": [3]}, state="unbatched")
    S["controlled_scalar_named"] = dict(spec=ctl, num_qubits=4, inputs={"phi": []}, state="unbatched")
    # the other ansaetze, with the reference's FORWARD semantics (reversed CNOT chain, quirk Q8); the
    # reference's TwoLocal / SpecialUnitary forward only takes an unbatched state
    S["twolocal_d3"] = dict(spec=[("RX", {"qubit": 1, "theta": 0.4, "remapping": NONE}), ("TwoLocal", {"qubits": [0, 1, 2, 3], "depth": 3}),
                                  ("MeasureProbability", {})], num_qubits=4, inputs={}, state="unbatched")
    S["twolocal_sub"] = dict(spec=[("TwoLocal", {"qubits": [2, 0, 3], "depth": 2}), ("MeasureState", {})],
                             num_qubits=4, inputs={}, state="unbatched")
    S["su_ry_rz"] = dict(spec=[("SpecialUnitary", {"qubits": [0, 1, 2, 3], "reps": 2, "rotations": ["ry", "rz"]}),
                               ("MeasureProbability", {})], num_qubits=4, inputs={}, state="unbatched")
    S["su_default_sub"] = dict(spec=[("RY", {"qubit": 0, "theta": 1.0, "remapping": NONE}), ("SU", {"qubits": [3, 1, 0], "reps": 1}),
                                     ("MeasureState", {})], num_qubits=4, inputs={}, state="unbatched")
    S["sel_budget_cnot"] = dict(spec=[("StronglyEntanglingLayerBudget", {"num_qubits_total": 4, "param_budget": 13}),
                                      ("MeasureProbability", {})], num_qubits=4, inputs={}, state="batched", batch=5)
    S["sel_budget_cz_sub"] = dict(
        spec=[("StronglyEntanglingLayerBudget", {"num_qubits_total": 5, "qubits": [0, 2, 3, 4], "param_budget": 19,
                                                 "control_gate": "__cls__:CZ", "control_gate_spacing": 2,
                                                 "rotations": ["rx", "rz", "ry"], "remapping": NONE}),
              ("MeasureJointProbability", {})], num_qubits=5, inputs={}, state="batched", batch=3)
    return S


def dumps(spec):
    return json.dumps(spec)


def loads(s):
    return [(n, kw) for n, kw in json.loads(s)]


def full_tile_specs():
    """name -> dict(spec, num_qubits, batch, inputs).  Few gates (each costs the reference seconds), chosen to touch the lowest and
    highest qubits, both CNOT directions across the tile boundary, a CZ, a per-sample named input and the default remapping."""
    S = {}
    for name, n, batch, meas in (("ft12_joint", 12, 3, "MeasureJointProbability"), ("ft13_probs", 13, 2, "MeasureProbability")):
        h = n - 1
        spec = [
            ("RY", {"qubit": 0, "remapping": NONE}), ("RX", {"qubit": h, "remapping": NONE}), ("CNOT", {"control": 0, "target": h}),
            ("RZ", {"qubit": n // 2}), ("RY", {"qubit": h, "name": "phi", "remapping": NONE}), ("CNOT", {"control": h, "target": 5}),
            ("RX", {"qubit": 0, "theta": 0.37}), ("CZ", {"control": 3, "target": h - 1}), ("RY", {"qubit": 7, "remapping": NONE}),
            ("SWAP", {"a": 1, "b": h}), ("RZ", {"qubit": 1, "remapping": NONE}),
            (meas, {}),
        ]
        S[name] = dict(spec=spec, num_qubits=n, batch=batch, inputs={"phi": [batch]})
    # a strongly-entangling block over qubits on both sides of the tile boundary (the BASELINE ansatz' structure), full state out
    # (12 qubits: the reference keeps every gate's dense matrix for its backward -- the same 46 gates at 13 qubits need > 64 GB)
    S["ft12_sel_state"] = dict(
        spec=[("StronglyEntanglingLayer", {"qubits": [0, 3, 6, 9, 11], "depth": 2}), ("RY", {"qubit": 10, "name": "phi"}),
              ("StronglyEntanglingLayer", {"qubits": [11, 1, 10, 2], "depth": 1, "remapping": NONE}), ("MeasureState", {})],
        num_qubits=12, batch=2, inputs={"phi": [2]})
    return S
