"""junction_mpc: B200-native batched MPC step for JunctionSim's path-tracking controller.

Host side is Python (as the reference is); all arithmetic of the hot path runs in hand-written sm_100a CUDA
kernels behind the C ABI declared in include/jmpc.h.  There is no CPU fallback: importing the pieces that
need the library raises when `libjmpc.so` is missing.
"""
from .config import MPCConfig, NPARAM, PARAM_INDEX, PARAM_NAMES  # noqa: F401

__all__ = ["MPCConfig", "NPARAM", "PARAM_INDEX", "PARAM_NAMES"]
