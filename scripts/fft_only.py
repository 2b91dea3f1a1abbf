"""r2c + one c2r with the fused gravity kernel on an N^3 mesh (2 buffers), for ncu captures of the FFT passes."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastpm_b200 import device, _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
m = device.Mesh(n, float(n) / 2)
lib = m.lib
if os.environ.get("SWAP"):
    ck, real = device.DeviceBuffer(m.alloc_floats * 4), device.DeviceBuffer(m.alloc_floats * 4)
else:
    real, ck = device.DeviceBuffer(m.alloc_floats * 4), device.DeviceBuffer(m.alloc_floats * 4)
print("real %x ck %x" % (real.ptr, ck.ptr))
_lib.check(lib.fpm_fill_whitenoise(m.h, real.ptr, 1), "noise")
kern = m.transfer_for_kernel("1_4", 0, 1)
lib.fpm_prof_enable(1)
for _ in range(reps):
    m.r2c(real, ck)
    m.c2r(ck, real, kern)
_lib.check(lib.fpm_sync(), "sync")
cls = (C.c_int32 * 256)()
ms = (C.c_double * 256)()
nl = lib.fpm_prof_get_launches(cls, ms, 256)
names = ["paint", "readout", "fft_tile", "fft_z", "kick", "drift", "kspace", "pk", "summary", "other"]
S = m.alloc_floats * 4.0
for i in range(min(nl, 256)):
    print("%-9s %8.3f ms  %7.1f GB/s" % (names[cls[i]], ms[i], 2 * S / ms[i] / 1e6))
