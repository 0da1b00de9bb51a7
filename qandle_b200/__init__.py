"""qandle_b200 -- B200-native state-vector engine behind the QANDLE operator API.

Same public names as the reference package (reference src/qandle/__init__.py:3-14 re-exports) for the hot
path named in BASELINE.json: Circuit, RX/RY/RZ/U, CNOT/CZ/SWAP, AngleEmbedding/AmplitudeEmbedding,
MeasureProbability/MeasureState/MeasureJointProbability, StronglyEntanglingLayer, named inputs and weight
remapping.  Execution goes through hand-written sm_100a kernels (qandle_b200/csrc); there is no CPU fallback.
"""
# ruff: noqa: F401 F403
from . import config
from .ansaetze import *
from .convolution import *
from .embeddings import *
from .errors import *
from .graphs import *
from .measurements import *
from .operators import *
from .qasm import *
from .qcircuit import *
from .utils import parse_rot

__version__ = "0.1.0"
