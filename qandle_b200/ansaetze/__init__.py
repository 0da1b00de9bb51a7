"""Ansaetze: StronglyEntanglingLayer (hot path) and, from SURVEY 8f rank 1, the packed-weight SEL plus
TwoLocal / SpecialUnitary / StronglyEntanglingLayerBudget with the reference's forward semantics."""
from .hardware_efficient import (  # noqa: F401
    SU,
    SpecialUnitary,
    SpecialUnitaryBuilt,
    StronglyEntanglingLayerBudget,
    TwoLocal,
    TwoLocalBuilt,
)
from .stronglyentangling import (  # noqa: F401
    StronglyEntanglingLayer,
    StronglyEntanglingLayerBuilt,
    StronglyEntanglingLayerPacked,
    StronglyEntanglingLayerPackedBuilt,
)

__all__ = ["StronglyEntanglingLayer", "StronglyEntanglingLayerBuilt", "StronglyEntanglingLayerPacked", "StronglyEntanglingLayerPackedBuilt",
           "TwoLocal", "TwoLocalBuilt", "SpecialUnitary", "SpecialUnitaryBuilt", "SU", "StronglyEntanglingLayerBudget"]
