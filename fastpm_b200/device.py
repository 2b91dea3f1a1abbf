"""Thin numpy-facing wrappers over the C ABI: device buffers, the PM mesh object and layout helpers.

Mirrors the reference's mesh API (api/fastpm/pmapi.h: pm_alloc, pm_free, pm_r2c, pm_c2r, ...)
and painter API (api/fastpm/painter.h) for the hot path.  All arithmetic happens in the CUDA
library; numpy is only used to marshal host data.
"""
import ctypes as C
import numpy as np

from . import _lib
from ._lib import FpmTransfer, check

FORCE_MODES = {"fastpm": 0, "pm": 1, "cola": 2, "2lpt": 3, "za": 4}            # api/fastpm/libfastpm.h:38-43
KERNELS = {"3_4": 0, "3_2": 1, "5_4": 2, "1_4": 3, "1_4_diff0": 4, "gadget": 5, "eastwood": 6, "naive": 7}


class DeviceBuffer:
    """A block of device memory (fpm_malloc); freed when dropped."""

    def __init__(self, nbytes):
        self.lib = _lib.require_device()
        self.nbytes = int(nbytes)
        self.ptr = self.lib.fpm_malloc(self.nbytes)
        if not self.ptr:
            raise _lib.FastPMB200Error("fpm_malloc(%d) failed: %s" % (self.nbytes, self.lib.fpm_last_error().decode()))

    @classmethod
    def from_host(cls, arr):
        arr = np.ascontiguousarray(arr)
        b = cls(arr.nbytes)
        b.upload(arr)
        return b

    def upload(self, arr, offset=0):
        arr = np.ascontiguousarray(arr)
        assert offset + arr.nbytes <= self.nbytes
        check(self.lib.fpm_memcpy_h2d(self.ptr + offset, arr.ctypes.data, arr.nbytes), "h2d")

    def download(self, dtype, count=None, offset=0):
        dtype = np.dtype(dtype)
        if count is None:
            count = (self.nbytes - offset) // dtype.itemsize
        out = np.empty(count, dtype=dtype)
        check(self.lib.fpm_memcpy_d2h(out.ctypes.data, self.ptr + offset, out.nbytes), "d2h")
        return out

    def zero(self):
        check(self.lib.fpm_memset(self.ptr, 0, self.nbytes), "memset")

    def free(self):
        if self.ptr:
            self.lib.fpm_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Mesh:
    """The PM mesh (struct PM of the reference, pmpfft.h:43-70), one x-slab of it per rank."""

    def __init__(self, nmesh, boxsize, nranks=1, rank=0):
        self.lib = _lib.require_device()
        self.h = self.lib.fpm_mesh_create(int(nmesh), float(boxsize), int(nranks), int(rank))
        if not self.h:
            raise _lib.FastPMB200Error("fpm_mesh_create: " + self.lib.fpm_last_error().decode())
        info = np.zeros(16, dtype=np.int64)
        check(self.lib.fpm_mesh_info(self.h, info.ctypes.data), "mesh_info")
        (self.n, self.alloc_floats, self.pitch_r, self.pitch_c, self.nxl, self.x0, self.nyl, self.y0,
         self.nranks, self.rank, self.halo) = [int(v) for v in info[:11]]
        self.boxsize = float(boxsize)

    def close(self):
        if self.h:
            self.lib.fpm_mesh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- pm_alloc (pmapi.c:11): zero-filled mesh buffer
    def alloc(self):
        b = DeviceBuffer(self.alloc_floats * 4)
        b.zero()
        return b

    def ktables(self):
        out = np.zeros((5, self.n), dtype=np.float32)
        check(self.lib.fpm_mesh_ktables_host(self.h, out.ctypes.data), "ktables")
        return dict(k=out[0], kk=out[1], k_finite=out[2], kk_finite=out[3], kk_finite2=out[4])

    # -- host <-> device layout helpers (single rank views)
    def upload_real(self, buf, arr):
        """arr[nxl, N, N] float32 -> device real layout."""
        n = self.n
        full = np.zeros((self.nxl + self.halo, n, self.pitch_r), dtype=np.float32)
        full[:self.nxl, :, :n] = arr
        buf.upload(full)

    def download_real(self, buf, halo=False):
        n = self.n
        planes = self.nxl + (self.halo if halo else 0)
        full = buf.download(np.float32, planes * n * self.pitch_r).reshape(planes, n, self.pitch_r)
        return full[:, :, :n].copy()

    def upload_complex(self, buf, arr):
        """arr[kx, ky_local, kz] complex64 (kz = 0..N/2) -> device [ky][kx][pitch_c]."""
        n = self.n
        full = np.zeros((self.nyl, n, self.pitch_c), dtype=np.complex64)
        full[:, :, :n // 2 + 1] = np.transpose(arr, (1, 0, 2))
        buf.upload(full)

    def download_complex(self, buf):
        """device k-space buffer -> complex64 [kx, ky_local, kz]."""
        n = self.n
        full = buf.download(np.complex64, self.nyl * n * self.pitch_c).reshape(self.nyl, n, self.pitch_c)
        return np.transpose(full[:, :, :n // 2 + 1], (1, 0, 2)).copy()

    # -- hot path
    def paint(self, canvas, x_dev, np_, M0=1.0, mass=None, field=None, field_stride=1):
        check(self.lib.fpm_paint(self.h, canvas.ptr, x_dev.ptr, int(np_), float(M0),
                                 mass.ptr if mass else None, field.ptr if field else None, int(field_stride)), "fpm_paint")

    def readout(self, canvas, x_dev, np_, out_dev, out_stride=1, out_offset_bytes=0, prescale=1.0):
        check(self.lib.fpm_readout(self.h, canvas.ptr, x_dev.ptr, int(np_), out_dev.ptr + out_offset_bytes,
                                   int(out_stride), float(prescale)), "fpm_readout")

    def readout3(self, canvases, x_dev, np_, out_dev):
        """out[i][d] = readout(canvases[d]) for d = 0, 1, 2 in one pass over the particles."""
        check(self.lib.fpm_readout3(self.h, canvases[0].ptr, canvases[1].ptr, canvases[2].ptr, x_dev.ptr, int(np_), out_dev.ptr), "fpm_readout3")

    def r2c(self, real, cplx, scale=None):
        if scale is None:
            scale = 1.0 / float(self.n) ** 3          # pm_r2c carries 1/Norm, pmpfft.c:382-385
        check(self.lib.fpm_r2c(self.h, real.ptr, cplx.ptr, float(scale)), "fpm_r2c")

    def c2r(self, cplx, real, kernel=None):
        check(self.lib.fpm_c2r(self.h, cplx.ptr, real.ptr, C.byref(kernel) if kernel is not None else None), "fpm_c2r")

    def transfer_for_kernel(self, kernel_type, attr, memb):
        t = FpmTransfer()
        kt = KERNELS[kernel_type] if isinstance(kernel_type, str) else int(kernel_type)
        check(self.lib.fpm_transfer_for_kernel(kt, int(attr), int(memb), C.byref(t)), "fpm_transfer_for_kernel")
        return t

    def apply_transfer(self, src, dst, kernel):
        check(self.lib.fpm_apply_transfer(self.h, src.ptr, dst.ptr, C.byref(kernel)), "fpm_apply_transfer")

    def decic(self, src, dst):
        check(self.lib.fpm_apply_decic(self.h, src.ptr, dst.ptr), "fpm_apply_decic")

    def powerspectrum(self, cplx, decic=False):
        nb = self.n // 2
        k, p, nm = np.zeros(nb), np.zeros(nb), np.zeros(nb)
        check(self.lib.fpm_powerspectrum(self.h, cplx.ptr, int(decic), k.ctypes.data, p.ctypes.data, nm.ctypes.data), "fpm_powerspectrum")
        return k, p, nm

    def cross_powerspectrum(self, cplx1, cplx2):
        """sum of w Re(d1 conj d2) per shell, normalised like powerspectrum() (powerspectrum.c:87-123)"""
        nb = self.n // 2
        sums = np.zeros(3 * nb + 1)
        check(self.lib.fpm_cross_powerspectrum_sums(self.h, cplx1.ptr, cplx2.ptr, sums.ctypes.data), "fpm_cross_powerspectrum_sums")
        nm = sums[:nb].copy()
        ok = nm > 0
        k, p = np.zeros(nb), np.zeros(nb)
        k[ok] = sums[2 * nb:3 * nb][ok] / nm[ok]
        p[ok] = sums[nb:2 * nb][ok] / nm[ok] * self.boxsize ** 3
        return k, p, nm

    def fill_gaussian_gadget(self, cplx, seed):
        """fastpm_ic_fill_gaussiank(..., FASTPM_DELTAK_GADGET): RANLUX white noise with the Gadget seeding scheme."""
        check(self.lib.fpm_fill_gaussian_gadget(self.h, cplx.ptr, int(seed)), "fpm_fill_gaussian_gadget")

    def induce_correlation(self, cplx, k, p):
        k = np.ascontiguousarray(k, dtype=np.float64)
        p = np.ascontiguousarray(p, dtype=np.float64)
        check(self.lib.fpm_induce_correlation(self.h, cplx.ptr, k.ctypes.data, p.ctypes.data, len(k)), "fpm_induce_correlation")


def ic_transfer(potorder=0, dirs=(), gradorder=1, scale=1.0):
    """laplace + diff of the IC / 2LPT path (pm2lpt.c:64-133), no negation.

    fastpm_apply_diff_transfer lacks an `else` after zeroing the self-conjugate modes (transfer.c:133-148),
    but every call in pm2lpt.c is IN PLACE (from == to), so the zero it just stored is what it reads back:
    the self-conjugate modes do end up zero, exactly as on the force path."""
    t = FpmTransfer()
    t.active, t.potorder, t.negate, t.ngrad = 1, potorder, 0, len(dirs)
    for i, d in enumerate(dirs):
        t.graddir[i] = d
    t.gradorder, t.zero_selfconj, t.scale = gradorder, 1, scale
    return t


def summary(buf, dtype, ncomp, np_):
    """fastpm_store_summary (store.c:808): returns dict of min, max, mean, std per component."""
    lib = _lib.require_device()
    out = np.zeros((ncomp, 4))
    check(lib.fpm_summary(buf.ptr, 8 if np.dtype(dtype) == np.float64 else 4, ncomp, int(np_), out.ctypes.data), "fpm_summary")
    n = float(np_)
    mean = out[:, 2] / n
    return dict(min=out[:, 0], max=out[:, 1], mean=mean, std=np.sqrt(out[:, 3] / n - mean ** 2))
