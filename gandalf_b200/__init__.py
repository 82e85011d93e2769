"""gandalf_b200 -- B200-native (sm_100a) grad-h SPH neighbour path behind GANDALF's
SphNeighbourSearch operator surface.  The compute lives in libgsph.so (CUDA, C ABI: include/gsph.h);
this package is the host-side mirror of the reference interface used by the tests and bench."""
from ._capi import GsphError, LIB_PATH  # noqa: F401
from .params import Parameters, config_from_params  # noqa: F401
from .sphtree import GpuSphTree, ParticleSet  # noqa: F401

__all__ = ["GpuSphTree", "ParticleSet", "Parameters", "config_from_params", "GsphError", "LIB_PATH"]
