"""Ansaetze on the hot path (StronglyEntanglingLayer; TwoLocal / SpecialUnitary / SEL-Budget are out of scope)."""
from .stronglyentangling import StronglyEntanglingLayer, StronglyEntanglingLayerBuilt  # noqa: F401
