"""components.flux_calculator_b200 -- B200-native (sm_100a) flux calculator hot path.

Product = libfluxcalc_b200.so (hand-written CUDA behind the C ABI of include/fluxcalc.h).
This package is the thin host-side mirror of the reference's interface for tests and benchmarks:
  flux_library               <- module flux_library            (flux_lib/flux_library.F90)
  flux_calculator_calculate  <- MODULE flux_calculator_calculate (flux_calculator_calculate.F90)
Importing it fails loudly if the CUDA library has not been built; there is no CPU fallback.
"""
from ._lib import lib, LIB_PATH, FluxCalcError, SIGNATURES        # noqa: F401
from .fields import IDX, VARNAMES, METHODS                         # noqa: F401
from .memory import DeviceArray, pinned_empty, pinned_free         # noqa: F401
from . import flux_library                                         # noqa: F401
from .flux_calculator_calculate import (FluxCalculator, NamelistCalculator, comm_get_unique_id, current_month,  # noqa: F401
                                        namelist_get, namelist_registry, nc_read_var, shard_range)
