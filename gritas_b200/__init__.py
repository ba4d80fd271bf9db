"""gritas_b200 -- B200-native (sm_100a CUDA) implementation of the GrITAS gridding hot path.

Package contents: csrc/ (CUDA kernels + the C ABI of include/gritas_b200.h), cabi (ctypes binding),
m_gritas_grids (host-side mirror of the reference's Fortran module interface), synth (seeded
ODS-shaped synthetic columns for the BASELINE.json configs).  Nothing here imports oracle/.
"""
from . import cabi, m_gritas_grids, synth  # noqa: F401
from .m_gritas_grids import AMISS, Grids, GrITAS_KxKt, GritasPerror  # noqa: F401

__version__ = "0.1.0"
