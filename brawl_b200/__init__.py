"""brawl_b200 -- B200-native atom-swap Monte-Carlo hot path behind BraWl's operator API.

Only what the path needs: csrc/ (CUDA kernels + the C ABI of include/brawl_cuda.h), the ctypes
binding and the host-side mirror of the reference's operator table.  No CPU fallback.
"""
from ._lib import BrawlCudaError, EXPORTS, LIB_PATH, load  # noqa: F401
from .engine import Device, RunParams, K_B_IN_RY, RY_TO_EV, LATTICES  # noqa: F401
from . import wang_landau  # noqa: F401,E402
from . import nested_sampling  # noqa: F401,E402
from . import replica_annealing  # noqa: F401,E402
from . import inputs, netcdf3, text_io  # noqa: F401,E402
