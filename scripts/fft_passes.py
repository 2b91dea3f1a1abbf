"""Times every FFT pass in four buffer / direction combinations (diagnostic for the transposing passes)."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastpm_b200 import _lib
if len(sys.argv) > 2:                       # A/B timing of another build of the library (development only)
    _lib.LIB_PATH = os.path.abspath(sys.argv[2])
from fastpm_b200 import device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
m = device.Mesh(n, float(n) / 2)
lib = m.lib
a, b = device.DeviceBuffer(m.alloc_floats * 4), device.DeviceBuffer(m.alloc_floats * 4)
_lib.check(lib.fpm_fill_whitenoise(m.h, a.ptr, 1), "noise")
names = ["paint", "readout", "fft_tile", "fft_z", "kick", "drift", "kspace", "pk", "summary", "other"]
S = m.alloc_floats * 4.0
lib.fpm_prof_enable(1)


def run(tag, fn):
    lib.fpm_prof_reset()
    fn()
    _lib.check(lib.fpm_sync(), "sync")
    cls = (C.c_int32 * 64)()
    ms = (C.c_double * 64)()
    nl = lib.fpm_prof_get_launches(cls, ms, 64)
    print("%-28s" % tag, "  ".join("%s %.2f" % (names[cls[i]][4:], ms[i]) for i in range(nl)))


kerns = [m.transfer_for_kernel("1_4", 0, d) for d in range(3)]
for rep in range(2):
    run("r2c a->b", lambda: m.r2c(a, b))
    run("c2r b->a plain", lambda: m.c2r(b, a))
    for d in range(3):
        run("c2r b->a force d=%d" % d, lambda: m.c2r(b, a, kerns[d]))
    run("r2c b->a", lambda: m.r2c(b, a))
    run("c2r a->b plain", lambda: m.c2r(a, b))
    _lib.check(lib.fpm_fill_whitenoise(m.h, a.ptr, 2), "noise")
