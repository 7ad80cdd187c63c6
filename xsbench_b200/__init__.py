"""xsbench_b200 -- B200-native (sm_100a) implementation of XSBench's macroscopic
cross-section lookup path.

Layout:
  csrc/   hand-written CUDA kernels + the C ABI of ``include/xs_gpu.h`` -> libxsb200.so
  host/   C host driver (CLI, data model, generator, report)           -> libxsb200_host.so, xsbench
  driver  Python mirror of the driver flow over the two libraries (tests, bench.py)

The package has no CPU implementation of the lookup and no fallback path.
"""
from . import _abi
from .driver import *  # noqa: F401,F403
from .driver import __all__ as _driver_all

__all__ = list(_driver_all) + ["_abi"]
__version__ = "0.1.0"
