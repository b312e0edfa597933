"""dang_b200: B200-native implementation of hermda02/dang's per-Gibbs-iteration hot path.

The compute path is the C-ABI CUDA library `libdang_gpu.so` (include/dang_gpu.h, built from
dang_b200/csrc); this package is the thin host-side mirror of the reference's operator
interface (`sample_cg_groups`, `sample_spectral_parameters`, `compute_chisq`, ...).
"""
from .config import (Band, CGGroup, Component, IndexSpec, RunConfig, flag_to_map_n,
                     return_poltype_flag)

__all__ = ["Band", "CGGroup", "Component", "IndexSpec", "RunConfig", "flag_to_map_n",
           "return_poltype_flag"]
