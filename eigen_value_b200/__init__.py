"""eigen_value_b200 -- B200-native similarity_transform() (max eigenpair of a positive matrix).

Only the hot path of itzmeanjan/eigen_value lives here: the CUDA round loop (csrc/), its C ABI
(include/similarity_transform.h -> libsimilarity_transform.so) and the host-side mirror of the
reference's Python interface.  Importing the package does not touch the GPU; the shared
library is loaded on first use and there is no CPU fallback.
"""
from .similarity_transform import (EPS, MAX_ITR, FORM_INPLACE, FORM_READONLY, STOP_ABSOLUTE, STOP_RELATIVE, ACC_F32, ACC_F64,
                                   EigenValue, Solver, SolveInfo, DeviceBuffer)

__all__ = ["EPS", "MAX_ITR", "FORM_INPLACE", "FORM_READONLY", "STOP_ABSOLUTE", "STOP_RELATIVE", "ACC_F32", "ACC_F64", "EigenValue",
           "Solver", "SolveInfo", "DeviceBuffer"]
