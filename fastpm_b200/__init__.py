"""fastpm_b200 -- B200 (sm_100a) particle-mesh force step behind FastPM's C API.

The package holds the CUDA library (csrc/, built to libfastpm_b200.so), its ctypes loader (_lib)
and numpy-facing wrappers (device, solver).  There is no CPU implementation in this package: the
CPU oracle lives under oracle/ and is only used by tests and the bench's CPU baseline.
"""
from ._lib import FastPMB200Error, LIB_PATH, load, require_device  # noqa: F401

__version__ = "0.1"
