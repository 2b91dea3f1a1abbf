"""Python binding of the host C layer (the libfastpm API mirror in include/fastpm_b200_api.h).

`Solver` drives fastpm_solver_init / fastpm_solver_setup_lpt / fastpm_solver_evolve of
libfastpm_b200.so exactly as src/fastpm.c drives libfastpm (src/fastpm.c:265-397); event handlers are
ordinary ctypes callbacks with the reference's signature (api/fastpm/events.h:13-14).
"""
import ctypes as C
import numpy as np

from . import _lib
from .device import FORCE_MODES, KERNELS

GROWTH_MODES = {"LCDM": 0, "ODE": 1}
COLUMNS = dict(mask=1 << 0, x=1 << 1, q=1 << 2, v=1 << 3, dx1=1 << 4, dx2=1 << 5, dv1=1 << 6, acc=1 << 7,
               id=1 << 8, potential=1 << 11, pgdc=1 << 13, mass=1 << 20)
_COL_DTYPE = dict(x=(np.float64, 3), v=(np.float32, 3), acc=(np.float32, 3), dx1=(np.float32, 3), dx2=(np.float32, 3),
                  id=(np.uint64, 1), potential=(np.float32, 1), pgdc=(np.float32, 3))

WINDOWS = dict(cic=0, linear=1, quad=2, lanczos=3)                                        # FastPMPainterType, painter.h
SOFTENINGS = dict(none=0, gaussian=1, gadget_long_range=2, two_third=3, gaussian36=4)     # FastPMSofteningType, libfastpm.h:52-54
HANDLER = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p)


class FastPMEvent(C.Structure):
    _fields_ = [("type", C.c_char * 32), ("stage", C.c_int)]


class ForceEvent(C.Structure):                       # api/fastpm/solver.h:38-47
    _fields_ = [("base", FastPMEvent), ("kernel", C.c_int), ("painter", C.c_void_p), ("pm", C.c_void_p),
                ("delta_k", C.c_void_p), ("N", C.c_double), ("a_f", C.c_double), ("a_n", C.c_double)]


class _State(C.Structure):
    _fields_ = [("force", C.c_int), ("x", C.c_int), ("v", C.c_int)]


class Transition(C.Structure):                       # api/fastpm/timemachine.h:16-29
    _fields_ = [("states", C.c_void_p), ("istart", C.c_int), ("iend", C.c_int), ("start", C.POINTER(_State)),
                ("end", C.POINTER(_State)), ("action", C.c_int),
                ("a_i", C.c_double), ("a_f", C.c_double), ("a_r", C.c_double),
                ("i_i", C.c_int), ("i_f", C.c_int), ("i_r", C.c_int)]


class TransitionEvent(C.Structure):
    _fields_ = [("base", FastPMEvent), ("transition", C.POINTER(Transition))]


def _bind(lib):
    vp, dbl, i64, i32 = C.c_void_p, C.c_double, C.c_int64, C.c_int
    lib.fastpm_b200_solver_new.restype = vp
    lib.fastpm_b200_solver_new.argtypes = [i64, dbl, vp, i32, dbl, dbl, i32, i32, i32, i32, dbl, dbl, dbl, dbl, dbl, i32]
    lib.fastpm_b200_solver_new_ex.restype = vp
    lib.fastpm_b200_solver_new_ex.argtypes = [i64, dbl, vp, i32, dbl, dbl, i32, i32, i32, i32, dbl, dbl, dbl, dbl, dbl, i32, vp, i32, i32, i32]
    lib.fastpm_b200_solver_free.argtypes = [vp]
    lib.fastpm_b200_solver_cdm.restype = vp
    lib.fastpm_b200_solver_cdm.argtypes = [vp]
    lib.fastpm_b200_solver_lptpm.restype = vp
    lib.fastpm_b200_solver_lptpm.argtypes = [vp]
    lib.fastpm_b200_store_np.restype = i64
    lib.fastpm_b200_store_np.argtypes = [vp]
    lib.fastpm_b200_store_set_np.argtypes = [vp, i64]
    lib.fastpm_b200_store_meta.argtypes = [vp, vp]
    lib.fastpm_b200_store_set_meta.argtypes = [vp, vp]
    lib.fastpm_b200_store_column_ptr.restype = vp
    lib.fastpm_b200_store_column_ptr.argtypes = [vp, i32]
    lib.fastpm_b200_store_get_column.argtypes = [vp, i32, vp, C.c_size_t, C.c_size_t]
    lib.fastpm_b200_store_set_column.argtypes = [vp, i32, vp, C.c_size_t, C.c_size_t]
    lib.fastpm_b200_mesh_host_size.restype = C.c_size_t
    lib.fastpm_b200_mesh_host_size.argtypes = [vp]
    lib.fastpm_b200_mesh_set_complex.argtypes = [vp, vp, vp]
    lib.fastpm_b200_mesh_get_complex.argtypes = [vp, vp, vp]
    lib.fastpm_b200_mesh_set_real.argtypes = [vp, vp, vp]
    lib.fastpm_b200_mesh_get_real.argtypes = [vp, vp, vp]
    lib.fastpm_b200_add_handler.argtypes = [vp, C.c_char_p, i32, HANDLER, vp]
    lib.fastpm_b200_kick_factor.argtypes = [vp, dbl, dbl, dbl, vp]
    lib.fastpm_b200_drift_factor.argtypes = [vp, dbl, dbl, dbl, vp]
    lib.fastpm_b200_growth.argtypes = [vp, dbl, vp]
    lib.fastpm_b200_schedule.argtypes = [vp, i32, vp, i32]
    lib.fastpm_solver_setup_lpt.argtypes = [vp, i32, vp, vp, dbl]
    lib.fastpm_solver_evolve.argtypes = [vp, vp, i32]
    lib.fastpm_find_pm.restype = vp
    lib.fastpm_find_pm.argtypes = [vp, dbl]
    lib.pm_alloc_details.restype = vp
    lib.pm_alloc_details.argtypes = [vp, C.c_char_p, i32]
    lib.pm_free.argtypes = [vp, vp]
    lib.pm_allocsize.restype = C.c_size_t
    lib.pm_allocsize.argtypes = [vp]
    lib.pm_r2c.argtypes = [vp, vp, vp]
    lib.pm_c2r.argtypes = [vp, vp]
    lib.fastpm_powerspectrum_init_from_delta.argtypes = [vp, vp, vp, vp]
    lib.fastpm_b200_clock_get.argtypes = [C.c_char_p, C.POINTER(dbl)]
    lib.fastpm_b200_memory_trim.argtypes = []
    lib.fastpm_b200_setup_synthetic_ic.argtypes = [vp, C.c_uint64, vp, vp, i32, dbl]
    return lib


class Solver:
    """FastPMSolver on one GPU (one slab)."""

    def __init__(self, nc, boxsize, pm_nc_factor=2, force_mode="fastpm", kernel_type="1_4", growth_mode="ODE",
                 np_alloc_factor=1.0, lpt_nc_factor=1, compute_potential=False, Omega_m=0.307494, h=0.6774,
                 T_cmb=0.0, N_eff=3.046, N_nu=0, nLPT=-2.5, pgdc=None, softening="none", painter="cic", painter_support=2, use_shift=False, use_dx1_only=False):
        """pgdc: None, or (alpha0, A, B, kl, ks) to switch the PGD correction on (pgdcorrection.c, src/fastpm.c:204-217);
        softening: "none", "gaussian", "gadget_long_range", "two_third", "gaussian36" (gravity.c:244-270);
        painter: "cic", "linear", "quad", "lanczos" (+ painter_support for lanczos; painter.c:128-174; any number of GPUs);
        use_shift: ICs at cell centres (FastPMConfig.USE_SHIFT, solver.c:142-150)."""
        self.lib = _bind(_lib.require_device())
        par = None if pgdc is None else np.array([float(v) for v in pgdc], dtype=np.float64)
        if par is not None and par.shape != (5,):
            raise ValueError("pgdc = (alpha0, A, B, kl, ks)")
        pairs = pm_nc_factor if isinstance(pm_nc_factor, (list, tuple)) else [(0.0, pm_nc_factor)]
        flat = np.array([v for pr in pairs for v in pr], dtype=np.float64)
        self.nc, self.boxsize, self.force_mode = int(nc), float(boxsize), force_mode
        self._handlers = []
        self.lib.fastpm_b200_solver_next_options(int(use_shift), int(use_dx1_only))
        self.h = self.lib.fastpm_b200_solver_new_ex(int(nc), float(boxsize), flat.ctypes.data, len(pairs), float(np_alloc_factor),
                                                    float(lpt_nc_factor), FORCE_MODES[force_mode], KERNELS[kernel_type],
                                                    GROWTH_MODES[growth_mode], int(compute_potential), float(nLPT),
                                                    float(Omega_m), float(h), float(T_cmb), float(N_eff), int(N_nu),
                                                    None if par is None else par.ctypes.data, SOFTENINGS[softening],
                                                    WINDOWS[painter], int(painter_support))
        self.cdm = self.lib.fastpm_b200_solver_cdm(self.h)
        self.lptpm = self.lib.fastpm_b200_solver_lptpm(self.h)

    def close(self):
        if self.h:
            self.lib.fastpm_b200_solver_free(self.h)
            self.lib.fastpm_b200_memory_trim()
            self.h = None

    # ---- snapshot files (bigfile directories, libfastpmio/io.c): csrc/host/io.c
    def write_snapshot(self, filebase, sort_by_id=False):
        self.lib.fastpm_b200_write_snapshot.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        self.lib.fastpm_b200_write_snapshot(self.h, str(filebase).encode(), int(sort_by_id))

    def add_snapshots(self, base, aout, sort_by_id=False):
        """Take snapshots "<base>_%0.04f" at the scale factors aout during the next evolve (the CLI's check_snapshots)."""
        ao = np.ascontiguousarray(aout, dtype=np.float64)
        self.lib.fastpm_b200_add_snapshot_handler.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int]
        self.lib.fastpm_b200_add_snapshot_handler(self.h, str(base).encode(), ao.ctypes.data, len(ao), int(sort_by_id))

    def read_snapshot(self, filebase):
        """Restart: reads the catalog into the particle store, returns the scale factor of the snapshot."""
        self.lib.fastpm_b200_read_snapshot.argtypes = [C.c_void_p, C.c_char_p]
        self.lib.fastpm_b200_read_snapshot.restype = C.c_double
        return float(self.lib.fastpm_b200_read_snapshot(self.h, str(filebase).encode()))

    # ---- particles (device-resident; these are host mirrors)
    @property
    def np(self):
        return int(self.lib.fastpm_b200_store_np(self.cdm))

    def set_np(self, n):
        if self.lib.fastpm_b200_store_set_np(self.cdm, int(n)) != 0:
            raise _lib.FastPMB200Error("particle count %d exceeds the allocated %s" % (n, "np_upper"))

    def column_ptr(self, name):
        return self.lib.fastpm_b200_store_column_ptr(self.cdm, COLUMNS[name])

    def get_column(self, name, out=None):
        dt, nm = _COL_DTYPE[name]
        n = self.np
        if out is None:
            out = np.empty((n, nm) if nm > 1 else (n,), dtype=dt)
        _lib.check(self.lib.fastpm_b200_store_get_column(self.cdm, COLUMNS[name], out.ctypes.data, 0, n), "get_column " + name)
        return out

    def set_column(self, name, arr):
        dt, nm = _COL_DTYPE[name]
        arr = np.ascontiguousarray(arr, dtype=dt)
        _lib.check(self.lib.fastpm_b200_store_set_column(self.cdm, COLUMNS[name], arr.ctypes.data, 0, len(arr)), "set_column " + name)

    @property
    def meta(self):
        m = np.zeros(3)
        self.lib.fastpm_b200_store_meta(self.cdm, m.ctypes.data)
        return dict(a_x=m[0], a_v=m[1], M0=m[2])

    def set_meta(self, a_x, a_v, M0):
        m = np.array([a_x, a_v, M0], dtype=np.float64)
        self.lib.fastpm_b200_store_set_meta(self.cdm, m.ctypes.data)

    # ---- IC
    def lpt_host_size(self):
        return int(self.lib.fastpm_b200_mesh_host_size(self.lptpm))

    def setup_lpt(self, delta_k_host, a0):
        """delta_k_host: float32 array in the reference's untransposed layout [x][y][N/2+1] complex (pm_alloc(lptpm))."""
        dk = self.lib.pm_alloc_details(self.lptpm, b"solver.py", 0)
        arr = np.ascontiguousarray(delta_k_host, dtype=np.float32)
        assert arr.size == self.lpt_host_size(), (arr.size, self.lpt_host_size())
        _lib.check(self.lib.fastpm_b200_mesh_set_complex(self.lptpm, dk, arr.ctypes.data), "mesh_set_complex")
        self.lib.fastpm_solver_setup_lpt(self.h, 1, dk, None, float(a0))
        self.lib.pm_free(self.lptpm, dk)

    def setup_synthetic_ic(self, seed, k, p, a0):
        """Counter-based white noise coloured by the table P(k), then 2LPT at a0, all on the device."""
        k = np.ascontiguousarray(k, dtype=np.float64)
        p = np.ascontiguousarray(p, dtype=np.float64)
        self.lib.fastpm_b200_setup_synthetic_ic(self.h, int(seed), k.ctypes.data, p.ctypes.data, len(k), float(a0))

    def setup_ic(self, seed, k, p, a0, remove_variance=False):
        """The reference's initial conditions for this seed (Gadget-scheme RANLUX noise, table P(k), 2LPT at a0), on the device."""
        k = np.ascontiguousarray(k, dtype=np.float64)
        p = np.ascontiguousarray(p, dtype=np.float64)
        self.lib.fastpm_b200_setup_gadget_ic.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double]
        self.lib.fastpm_b200_setup_gadget_ic(self.h, int(seed), int(remove_variance), k.ctypes.data, p.ctypes.data, len(k), float(a0))

    def setup_lpt_device(self, delta_k_dev_ptr, a0):
        self.lib.fastpm_solver_setup_lpt(self.h, 1, delta_k_dev_ptr, None, float(a0))

    # ---- events
    def add_handler(self, etype, stage, pyfunc):
        """pyfunc(solver_ptr, event_ptr, userdata) -> int; stage 0 = before, 1 = after."""
        cb = HANDLER(pyfunc)
        self._handlers.append(cb)
        self.lib.fastpm_b200_add_handler(self.h, etype.encode(), int(stage), cb, None)

    def evolve(self, time_step):
        ts = np.ascontiguousarray(time_step, dtype=np.float64)
        self.lib.fastpm_solver_evolve(self.h, ts.ctypes.data, len(ts))
        _lib.check(self.lib.fpm_sync(), "sync")

    # ---- scalars
    def kick_factor(self, ai, ac, af):
        o = np.zeros(101)
        self.lib.fastpm_b200_kick_factor(self.h, ai, ac, af, o.ctypes.data)
        return dict(ai=o[0], ac=o[1], af=o[2], q1=o[3], q2=o[4], dda=o[5:37].copy(), Dv1=o[37:69].copy(), Dv2=o[69:101].copy())

    def drift_factor(self, ai, ac, af):
        o = np.zeros(101)
        self.lib.fastpm_b200_drift_factor(self.h, ai, ac, af, o.ctypes.data)
        return dict(ai=o[0], ac=o[1], af=o[2], Dv1=o[3], Dv2=o[4], dyyy=o[5:37].copy(), da1=o[37:69].copy(), da2=o[69:101].copy())

    def growth(self, a):
        o = np.zeros(12)
        self.lib.fastpm_b200_growth(self.h, float(a), o.ctypes.data)
        return dict(zip(["D1", "D2", "f1", "f2", "E", "dEda", "d2Eda2", "dD1da", "d2D1da2", "Omega_a", "Omega_Lambda", "Omega_cdm"], o))

    def powerspectrum_of(self, pm_ptr, delta_k_ptr):
        """fastpm_powerspectrum_init_from_delta on a device buffer -> (k, p, nmodes)."""
        class FuncK(C.Structure):
            _fields_ = [("size", C.c_size_t), ("k", C.POINTER(C.c_double)), ("f", C.POINTER(C.c_double))]

        class PS(C.Structure):
            _fields_ = [("base", FuncK), ("edges", C.POINTER(C.c_double)), ("pm", C.c_void_p), ("k0", C.c_double),
                        ("Volume", C.c_double), ("Nmodes", C.POINTER(C.c_double))]
        ps = PS()
        self.lib.fastpm_powerspectrum_init_from_delta(C.byref(ps), pm_ptr, delta_k_ptr, delta_k_ptr)
        n = ps.base.size
        k = np.ctypeslib.as_array(ps.base.k, (n,)).copy()
        p = np.ctypeslib.as_array(ps.base.f, (n,)).copy()
        nm = np.ctypeslib.as_array(ps.Nmodes, (n,)).copy()
        self.lib.fastpm_powerspectrum_destroy(C.byref(ps))
        return k, p, nm


def schedule(time_step):
    lib = _bind(_lib.load())
    ts = np.ascontiguousarray(time_step, dtype=np.float64)
    rows = np.zeros((5 * len(ts) + 8, 7))
    n = lib.fastpm_b200_schedule(ts.ctypes.data, len(ts), rows.ctypes.data, len(rows))
    return rows[:n]
