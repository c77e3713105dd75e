"""B200-native batched solver for the GAB1-SHP2 reaction-diffusion model's hot path.

`host` mirrors the reference's Julia call surface on top of the C ABI in include/gab1pde.h, `abi` is the
ctypes binding of libgab1pde.so (hand-written sm_100a CUDA, csrc/), `params` supplies grids, time steps and
synthetic ensembles.  There is no CPU implementation in this package.
"""
from . import abi, params, host  # noqa: F401
from .host import *  # noqa: F401,F403

__all__ = ["abi", "params", "host"]
