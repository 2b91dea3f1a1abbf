"""ORACLE / TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/_ref/liboracle_port.so: the plain-C
restatement of the path under oracle/port/ plus the CPU FFT of oracle/shims/src/cpufft.c.
Unlike oracle/_ref/libfastpm_ref.so this library needs no reference tree to build."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "liboracle_port.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        _lib.oracle_fft3_plan_new.restype = C.c_void_p
        _lib.oracle_fft3_plan_new.argtypes = [C.c_ssize_t, C.c_ssize_t, C.c_ssize_t, C.c_int]
        _lib.oracle_fft3_plan_free.argtypes = [C.c_void_p]
        _lib.oracle_fft3_exec_r2c.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oracle_fft3_exec_c2r.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


def fft3_r2c(field):
    """Unnormalised forward r2c of a real [n,n,n] float32 array -> complex64 [kx,ky,kz<=n/2]."""
    n = field.shape[0]
    buf = np.zeros((n, n, n + 2), dtype=np.float32)
    buf[:, :, :n] = field
    p = lib().oracle_fft3_plan_new(n, n, n, 0)
    lib().oracle_fft3_exec_r2c(p, buf.ctypes.data, buf.ctypes.data)
    lib().oracle_fft3_plan_free(p)
    return buf.reshape(n, n, n // 2 + 1, 2).view(np.complex64)[..., 0].copy()


def fft3_c2r(cfield):
    """Unnormalised inverse of fft3_r2c."""
    n = cfield.shape[0]
    buf = np.zeros((n, n, n // 2 + 1), dtype=np.complex64)
    buf[...] = cfield
    fb = buf.view(np.float32).reshape(n, n, n + 2)
    p = lib().oracle_fft3_plan_new(n, n, n, 0)
    lib().oracle_fft3_exec_c2r(p, fb.ctypes.data, fb.ctypes.data)
    lib().oracle_fft3_plan_free(p)
    return fb[:, :, :n].copy()
