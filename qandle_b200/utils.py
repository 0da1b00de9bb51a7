"""Helpers on the hot path (reference src/qandle/utils.py:26-37 parse_rot; the einops/splitter helpers are
out of scope)."""


def parse_rot(rot: str):
    """'rx' | 'x' | 'RX' ... -> the rotation operator class (reference utils.py:26-37)."""
    from . import operators as op

    r = rot.lower().replace("r", "")
    if r == "x":
        return op.RX
    if r == "y":
        return op.RY
    if r == "z":
        return op.RZ
    raise ValueError(f"Unknown rotation {rot}")
