"""b200msm -- host-side mirror of the reference's MSM interface over the B200 engine's C ABI.

Layout: _lib (ctypes binding, fails loudly without the .so), engine (context owner),
protoboard (the reference's test-rig API for this path), sharded (point-range sharding across GPUs).
"""
from ._lib import (lib, B200MsmError, Stats, N8, BLS12_381_G1, BN254_G1, BLS12_381_G2, BN254_G2, EXPORTS, LIB_PATH, constants, strerror)
from .engine import Engine
from .protoboard import Protoboard
from .ffjs import G1, G2

__all__ = ["Engine", "Protoboard", "G1", "G2", "B200MsmError", "Stats", "N8", "BLS12_381_G1", "BN254_G1", "BLS12_381_G2", "BN254_G2", "EXPORTS", "LIB_PATH", "constants", "strerror", "lib"]
