"""sdvar_b200 -- B200-native (sm_100a) implementation of SDVAR's speculative draft-then-verify
next-scale generation loop, behind the reference's Python API.

Device work is done by hand-written CUDA kernels in ``libsdvar_b200.so`` (C ABI in
``include/sdvar_b200.h``), reached through ``sdvar_b200._cabi``; there is no CPU fallback.
"""
__version__ = "0.1.0"
