"""bgreat_b200: B200-native read-to-de-Bruijn-graph mapping (the hot path of Malfoy/BGREAT).

Host side mirrors the reference's `class Aligner` (reference aligner.h:58); the work is done by the
sm_100a kernels behind the C ABI in include/bgreat_b200.h. There is no CPU fallback.
"""
from .aligner import Aligner, MappingResult  # noqa: F401
from ._abi import (MODE_GREEDY, MODE_EXHAUSTIVE, MODE_GREEDY_PATH, MODE_EXHAUSTIVE_PATH, BgrError)  # noqa: F401

__all__ = ["Aligner", "MappingResult", "MODE_GREEDY", "MODE_EXHAUSTIVE", "MODE_GREEDY_PATH", "MODE_EXHAUSTIVE_PATH",
           "BgrError"]
