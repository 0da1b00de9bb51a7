"""Ansaetze on the hot path (StronglyEntanglingLayer; TwoLocal / SpecialUnitary / SEL-Budget are out of scope)."""
from .stronglyentangling import (  # noqa: F401
    StronglyEntanglingLayer,
    StronglyEntanglingLayerBuilt,
    StronglyEntanglingLayerPacked,
    StronglyEntanglingLayerPackedBuilt,
)

__all__ = ["StronglyEntanglingLayer", "StronglyEntanglingLayerBuilt", "StronglyEntanglingLayerPacked", "StronglyEntanglingLayerPackedBuilt"]
