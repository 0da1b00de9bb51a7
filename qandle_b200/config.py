"""Mirror of reference src/qandle/config.py (only the values the hot path reads)."""
from . import remap

DEFAULT_MAPPING = remap.tanh
"""Default weight remapping, read at gate-construction time (reference operators.py:168)."""

# presentation-only settings kept for source compatibility (reference config.py:10-80)
DRAW_SPLITTED_PAD = 7
DRAW_DASH = "─"
DRAW_SHOW_VALUES = True
DRAW_SHIFT_LEFT = False
DRAW_CROSS_BETWEEN_CNOT = True

# ---- engine options (no reference counterpart) ----
import os as _os

ENGINE_TILE_BITS = int(_os.environ.get("QB_TILE_BITS", "0"))  # 0 = library default (12 for complex64, 11 for complex128)
ENGINE_LOW_BITS = int(_os.environ.get("QB_LOW_BITS", "0"))  # 0 = library default
ENGINE_FUSE = _os.environ.get("QB_FUSE", "1") != "0"
ENGINE_STAGED = _os.environ.get("QB_STAGED", "1") != "0"  # register-blocked staged sweep kernels
ENGINE_MAX_OPS_PER_SWEEP = int(_os.environ.get("QB_MAX_OPS", "0"))  # 0 = unlimited (maximal fusion)
ENGINE_PACKED = _os.environ.get("QB_PACKED", "1") != "0"  # complex64: packed FFMA2 kernel (planar shared memory)
ENGINE_FLAT = _os.environ.get("QB_FLAT", "1") != "0"  # complex64 packed kernel: straight-line (flat) stage bodies
ENGINE_SWEEP_SEARCH = _os.environ.get("QB_SWEEP_SEARCH", "1") != "0"  # planner: search over where sweeps end / 128-byte chunks (plan.cpp)
