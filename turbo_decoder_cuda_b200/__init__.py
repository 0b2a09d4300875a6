"""B200-native batched LTE turbo decoder: hand-written sm_100a CUDA kernels behind the C ABI in
include/tdb200.h.  `TurboDecoder` is a ctypes mirror of that ABI."""
from .decoder import (ALGO_LOGMAP_F32, ALGO_LOGMAP_F64, ALGO_LOGMAP_S16, ALGO_MAXLOG_F32, ALGO_MAXLOG_S16,  # noqa: F401
                      LIB_PATH, TdbError, TurboDecoder, load_library, lte_qpp_params)

__all__ = ["TurboDecoder", "TdbError", "load_library", "lte_qpp_params", "LIB_PATH",
           "ALGO_LOGMAP_F64", "ALGO_MAXLOG_S16", "ALGO_LOGMAP_S16", "ALGO_LOGMAP_F32", "ALGO_MAXLOG_F32"]
