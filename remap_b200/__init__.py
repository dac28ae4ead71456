"""remap_b200 -- B200-native registration hot path of kataklinger/remap (kpe + kpm, fde mask).

Product code: csrc/ (sm_100a kernels + the C ABI of include/remap_b200.h), api.py (ctypes face),
shard.py (multi-GPU frame-range sharding, position scan, fragment map extents); the frc::collector-shaped
C++ front is include/frc_b200.hpp.
synth.py is the synthetic workload generator used by tests and bench.py.
"""
from .api import (BIN_DTYPE, KEYPOINT_DTYPE, OFFSET_DTYPE, PLACEMENT_DTYPE, VOTE_DTYPE, FRAME_DIGEST_DTYPE, Group, Registrar, RemapError, Snippet, CELL_MATCH_DTYPE)  # noqa: F401
from ._lib import RB_OFFSET_TIE_SENSITIVE, RB_OFFSET_VALID  # noqa: F401
