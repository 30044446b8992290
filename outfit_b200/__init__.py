"""outfit_b200 -- B200-native (sm_100a, scalar FP64 CUDA) batched initial orbit determination.

The package is a thin host-side mirror of the reference's `FitIOD` / `propagate_universal`
interface over the C-ABI in include/outfit_b200.h (liboutfit_b200.so, built in-tree from
outfit_b200/csrc).  There is no CPU fallback: importing works anywhere, but every compute call
needs the compiled library AND a CUDA device and fails loudly otherwise.
"""
from .api import (  # noqa: F401
    IODParams, IodResult, OutfitB200, OutfitError, SolverType, library_path, load_library,
    RESULT_DTYPE, STATUS_NAMES, DifferentialCorrectionConfig, LSQ_RESULT_DTYPE, OBS_FIT_DTYPE,
    OutfitGroup, pinned_empty, shard_ranges, EphemerisConfig, NBodyConfig, planet_gm, orbits_of_results, ephemeris_mode_epochs,
)
