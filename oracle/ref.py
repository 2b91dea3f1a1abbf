"""ORACLE / TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/_ref/libfastpm_ref.so.

The library is the reference's own libfastpm sources compiled in place from
/root/reference against the shims in oracle/shims (see oracle/Makefile); this
module only marshals numpy arrays into oracle/ref_driver.c.  Importable only
from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference arm.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libfastpm_ref.so")

FORCE_MODES = {"fastpm": 0, "pm": 1, "cola": 2, "2lpt": 3, "za": 4}          # api/fastpm/libfastpm.h:38-43
KERNELS = {"3_4": 0, "3_2": 1, "5_4": 2, "1_4": 3, "1_4_diff0": 4, "gadget": 5, "eastwood": 6, "naive": 7}
GROWTH_MODES = {"LCDM": 0, "ODE": 1}                                           # api/fastpm/cosmology.h:6-9


WINDOWS = dict(cic=0, linear=1, quad=2, lanczos=3)                                        # FastPMPainterType, painter.h
SOFTENINGS = dict(none=0, gaussian=1, gadget_long_range=2, two_third=3, gaussian36=4)     # FastPMSofteningType, libfastpm.h:52-54


class RefConfig(C.Structure):
    _fields_ = [
        ("nc", C.c_int64), ("boxsize", C.c_double), ("pm_nc_factor", C.c_double * 8),
        ("np_alloc_factor", C.c_double), ("lpt_nc_factor", C.c_double),
        ("force_mode", C.c_int), ("kernel_type", C.c_int), ("growth_mode", C.c_int),
        ("compute_potential", C.c_int), ("use_dx1_only", C.c_int), ("verbose", C.c_int),
        ("nLPT", C.c_double),
        ("Omega_m", C.c_double), ("h", C.c_double), ("T_cmb", C.c_double), ("Omega_k", C.c_double),
        ("w0", C.c_double), ("wa", C.c_double), ("N_eff", C.c_double),
        ("N_nu", C.c_int), ("enforce_broadband_kmax", C.c_int),
        ("pgdc", C.c_double * 6), ("softening_type", C.c_int), ("painter_type", C.c_int), ("painter_support", C.c_int),
        ("use_shift", C.c_int),
    ]


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        _lib.ref_session_new.restype = C.c_void_p
        _lib.ref_ic_deltak.restype = C.c_double
        _lib.ref_evolve.restype = C.c_double
        _lib.ref_np.restype = C.c_int64
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Session:
    """One FastPMSolver of the reference (1 rank)."""

    def __init__(self, nc, boxsize, pm_nc_factor=2, force_mode="fastpm", kernel_type="1_4",
                 growth_mode="ODE", np_alloc_factor=4.0, lpt_nc_factor=1, compute_potential=False,
                 Omega_m=0.307494, h=0.6774, T_cmb=0.0, N_eff=3.046, N_nu=0, nLPT=-2.5,
                 use_dx1_only=False, verbose=False, enforce_broadband_kmax=4, pgdc=None, softening="none", painter="cic", painter_support=2, use_shift=False):
        """pgdc: None, or (alpha0, A, B, kl, ks) to switch the PGD correction on (src/fastpm.c:204-217)."""
        cfg = RefConfig()
        cfg.use_shift = int(use_shift)
        cfg.nc = nc
        cfg.boxsize = boxsize
        pairs = pm_nc_factor if isinstance(pm_nc_factor, (list, tuple)) else [(0.0, pm_nc_factor)]
        flat = [0.0] * 8
        for i, (a, b) in enumerate(pairs):
            flat[2 * i], flat[2 * i + 1] = float(a), float(b)
        cfg.pm_nc_factor = (C.c_double * 8)(*flat)
        cfg.np_alloc_factor = np_alloc_factor
        cfg.lpt_nc_factor = lpt_nc_factor
        cfg.force_mode = FORCE_MODES[force_mode]
        cfg.kernel_type = KERNELS[kernel_type]
        cfg.growth_mode = GROWTH_MODES[growth_mode]
        cfg.compute_potential = int(compute_potential)
        cfg.use_dx1_only = int(use_dx1_only)
        cfg.verbose = int(verbose)
        cfg.nLPT = nLPT
        cfg.Omega_m, cfg.h, cfg.T_cmb, cfg.Omega_k = Omega_m, h, T_cmb, 0.0
        cfg.w0, cfg.wa, cfg.N_eff, cfg.N_nu = -1.0, 0.0, N_eff, N_nu
        cfg.enforce_broadband_kmax = enforce_broadband_kmax
        cfg.softening_type = SOFTENINGS[softening]
        cfg.painter_type, cfg.painter_support = WINDOWS[painter], int(painter_support)
        cfg.pgdc = (C.c_double * 6)(*([0.0] * 6 if pgdc is None else [1.0] + [float(v) for v in pgdc]))
        self.cfg = cfg
        self.nc, self.boxsize, self.force_mode = nc, boxsize, force_mode
        self._h = C.c_void_p(lib().ref_session_new(C.byref(cfg)))

    def close(self):
        if self._h:
            lib().ref_session_free(self._h)
            self._h = None

    # ---- meshes: which = 0 force mesh at time a, 1 lpt mesh, 2 base mesh
    def pm_info(self, which=0, a=1.0):
        info = np.zeros(11, dtype=np.int64)
        lib().ref_pm_info(self._h, C.c_int(which), C.c_double(a), _p(info))
        return dict(Nmesh=int(info[0]), allocsize=int(info[1]), istrides=tuple(info[2:5]),
                    ostrides=tuple(info[5:8]), osize=tuple(info[8:11]))

    def _buf(self, which, a):
        return np.zeros(self.pm_info(which, a)["allocsize"], dtype=np.float32)

    def complex_view(self, buf, which=0, a=1.0):
        """k-space buffer -> complex array indexed [kx, ky, kz] (a copy)."""
        info = self.pm_info(which, a)
        n = info["Nmesh"]
        hc = n // 2 + 1
        c = buf.view(np.complex64)
        s = info["ostrides"]
        idx = (np.arange(n)[:, None, None] * s[0] + np.arange(n)[None, :, None] * s[1]
               + np.arange(hc)[None, None, :] * s[2])
        return c[idx]

    def complex_pack(self, arr, which=0, a=1.0):
        info = self.pm_info(which, a)
        n = info["Nmesh"]
        hc = n // 2 + 1
        out = np.zeros(info["allocsize"] // 2, dtype=np.complex64)
        s = info["ostrides"]
        idx = (np.arange(n)[:, None, None] * s[0] + np.arange(n)[None, :, None] * s[1]
               + np.arange(hc)[None, None, :] * s[2])
        out[idx] = arr
        return out.view(np.float32)

    def real_view(self, buf, which=0, a=1.0):
        n = self.pm_info(which, a)["Nmesh"]
        return buf.reshape(n, n, n + 2)[:, :, :n]

    def real_pack(self, arr, which=0, a=1.0):
        n = self.pm_info(which, a)["Nmesh"]
        out = np.zeros((n, n, n + 2), dtype=np.float32)
        out[:, :, :n] = arr
        return out.reshape(-1)

    # ---- IC
    def ic_deltak(self, seed, pk_text, remove_variance=False, linear_density_redshift=0.0):
        out = self._buf(1, 1.0)
        sigma8 = C.c_double()
        var = lib().ref_ic_deltak(self._h, C.c_int(seed), C.c_int(int(remove_variance)),
                                  C.c_char_p(pk_text.encode()), C.c_double(linear_density_redshift),
                                  _p(out), C.byref(sigma8))
        return out, float(var), sigma8.value

    def fill_gaussian(self, seed):
        """Raw Gadget-scheme white noise in k-space on the LPT mesh (pm_alloc layout of lptpm)."""
        out = self._buf(1, 1.0)
        lib().ref_fill_gaussian(self._h, C.c_int(seed), _p(out))
        return out

    def setup_lpt(self, delta_k, a0):
        lib().ref_setup_lpt(self._h, _p(np.ascontiguousarray(delta_k, dtype=np.float32)), C.c_double(a0))
        d1, d2 = np.zeros(3), np.zeros(3)
        lib().ref_lpt_std(self._h, _p(d1), _p(d2))
        return d1, d2

    def lpt_solve(self, delta_k):
        n = self.np
        dx1 = np.zeros((n, 3), dtype=np.float32)
        dx2 = np.zeros((n, 3), dtype=np.float32)
        lib().ref_2lpt_solve(self._h, _p(np.ascontiguousarray(delta_k, dtype=np.float32)), _p(dx1), _p(dx2))
        return dx1, dx2

    # ---- particles
    @property
    def np(self):
        return int(lib().ref_np(self._h))

    def get_particles(self):
        n = self.np
        out = dict(x=np.zeros((n, 3)), v=np.zeros((n, 3), dtype=np.float32), acc=np.zeros((n, 3), dtype=np.float32),
                   id=np.zeros(n, dtype=np.uint64), meta=np.zeros(3))
        dx1 = dx2 = None
        if self.force_mode == "cola":
            dx1 = out["dx1"] = np.zeros((n, 3), dtype=np.float32)
            dx2 = out["dx2"] = np.zeros((n, 3), dtype=np.float32)
        lib().ref_get_particles(self._h, _p(out["x"]), _p(out["v"]), _p(out["acc"]), _p(out["id"]), _p(dx1), _p(dx2), _p(out["meta"]))
        return out

    def pgdc_calculate(self, delta_k, x, par, which=0, a=1.0):
        """fastpm_pgdc_calculate for positions x: par = (alpha0, A, B, kl, ks); returns float32 [np][3]."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        par = np.ascontiguousarray(par, dtype=np.float64)
        dk = np.ascontiguousarray(delta_k, dtype=np.float32)
        out = np.zeros((len(x), 3), dtype=np.float32)
        lib().ref_pgdc(self._h, C.c_int(which), C.c_double(a), _p(dk), _p(x), C.c_int64(len(x)), _p(par), _p(out))
        return out

    def get_pgdc(self):
        out = np.zeros((self.np, 3), dtype=np.float32)
        lib().ref_get_pgdc(self._h, _p(out))
        return out

    def set_pgdc(self, pgdc):
        lib().ref_set_pgdc(self._h, _p(np.ascontiguousarray(pgdc, dtype=np.float32)))

    def remove_variance(self, dk):
        dk = np.ascontiguousarray(dk, dtype=np.float32)
        out = np.zeros_like(dk)
        lib().ref_remove_variance(self._h, _p(dk), _p(out))
        return out

    def mode_op(self, op, dk, mode=(0, 0, 0, 0), value=0.0, method=0):
        """op: "set_mode", "normalize", "c2r_weight" on the LPT mesh; returns (result buffer, get_mode(result, mode))."""
        dk = np.ascontiguousarray(dk, dtype=np.float32)
        out = np.zeros_like(dk)
        m = np.array(mode, dtype=np.int64)
        lib().ref_mode_op.restype = C.c_double
        r = lib().ref_mode_op(self._h, C.c_int(dict(set_mode=0, normalize=1, c2r_weight=2)[op]), _p(dk), _p(out), _p(m), C.c_double(value), C.c_int(method))
        return out, float(r)

    def write_snapshot(self, filebase):
        """write_snapshot_header + fastpm_store_write of the unit-converted CDM store (bigfile directory `filebase`)."""
        lib().ref_write_snapshot(self._h, C.c_char_p(str(filebase).encode()))

    def append_snapshot(self, filebase):
        """fastpm_store_write(..., "a"): the unit-converted CDM store appended to the catalog in `filebase`."""
        lib().ref_append_snapshot(self._h, C.c_char_p(str(filebase).encode()))

    def evolve_snapshots(self, time_step, base, aout):
        """evolve + the CLI's snapshot taking: directories "<base>_%0.04f" for every aout inside the run."""
        ts = np.ascontiguousarray(time_step, dtype=np.float64)
        ao = np.ascontiguousarray(sorted(aout), dtype=np.float64)
        lib().ref_evolve_snapshots(self._h, _p(ts), C.c_int(len(ts)), C.c_char_p(str(base).encode()), _p(ao), C.c_int(len(ao)))

    def snapshot_particles(self):
        """(x, v) after fastpm_set_species_snapshot at the particles' own time: v in km/s, x wrapped into the box."""
        x, v = np.zeros((self.np, 3)), np.zeros((self.np, 3), dtype=np.float32)
        lib().ref_snapshot_particles(self._h, _p(x), _p(v))
        return x, v

    def read_snapshot(self, filebase, restart=False):
        """fastpm_store_read of the CDM catalog; restart=True converts back to the integrator's units (src/fastpm.c:618-635)."""
        lib().ref_read_snapshot.restype = C.c_double
        return float(lib().ref_read_snapshot(self._h, C.c_char_p(str(filebase).encode()), C.c_int(int(restart))))

    def write_complex(self, dk, filename, blockname, which=0, a=1.0):
        dk = np.ascontiguousarray(dk, dtype=np.float32)
        lib().ref_write_complex(self._h, C.c_int(which), C.c_double(a), _p(dk), C.c_char_p(str(filename).encode()), C.c_char_p(blockname.encode()))

    def set_particles(self, x, v=None, id=None, dx1=None, dx2=None, meta=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        c = lambda a, t: None if a is None else np.ascontiguousarray(a, dtype=t)
        v, id, dx1, dx2, meta = c(v, np.float32), c(id, np.uint64), c(dx1, np.float32), c(dx2, np.float32), c(meta, np.float64)
        lib().ref_set_particles(self._h, C.c_int64(len(x)), _p(x), _p(v), _p(id), _p(dx1), _p(dx2), _p(meta))

    def wrap(self):
        """fastpm_store_wrap of the session's particles into [0, boxsize]"""
        lib().ref_wrap(self._h)

    def summary(self, column):
        """fastpm_store_summary of the 3-component column 'x' or 'v': dict of min, max, mean, std"""
        out = np.zeros(12)
        lib().ref_summary(self._h, C.c_int({"x": 1 << 1, "v": 1 << 3}[column]), _p(out))
        return dict(min=out[0:3], max=out[3:6], mean=out[6:9], std=out[9:12])

    # ---- evolve
    def evolve(self, time_step):
        ts = np.ascontiguousarray(time_step, dtype=np.float64)
        return float(lib().ref_evolve(self._h, _p(ts), C.c_int(len(ts))))

    def records(self):
        out = []
        for i in range(lib().ref_nrecords(self._h)):
            nb = lib().ref_record_nbins(self._h, C.c_int(i))
            sc, k, p, nm = np.zeros(16), np.zeros(nb), np.zeros(nb), np.zeros(nb)
            lib().ref_record(self._h, C.c_int(i), _p(sc), _p(k), _p(p), _p(nm))
            out.append(dict(a_f=sc[0], vel_std=sc[1:4].copy(), pos_min=sc[4:7].copy(), pos_max=sc[7:10].copy(),
                            acc_std=sc[10:13].copy(), Plin=sc[13], a_x=sc[14], a_v=sc[15], k=k, p=p, nmodes=nm))
        return out

    def subsample_probe(self, np_upper, fraction, in_place=False, sort_back=False):
        """fill a scratch store, fastpm_store_fill_subsample_mask + fastpm_store_subsample [+ permute (reverse) + sort by id]
        -> (ids, x, mask_sum)"""
        ids, x, ms = np.zeros(np_upper, dtype=np.uint64), np.zeros((np_upper, 3)), C.c_int64(0)
        lib().ref_subsample_probe.restype = C.c_int64
        n = lib().ref_subsample_probe(self._h, C.c_int64(np_upper), C.c_double(fraction), C.c_int(int(in_place)), C.c_int(int(sort_back)),
                                      _p(ids), _p(x), C.byref(ms))
        return ids[:n].copy(), x[:n].copy(), int(ms.value)

    def fill_probe(self, np_upper, as_rank=0):
        """fastpm_store_fill of a scratch store with q and rand columns -> (q [np][3], rand [np_upper]); as_rank: what MPI_Comm_rank
        answers meanwhile (selects the seed of the rand stream, store.c:704-708)"""
        q = np.zeros((int(np_upper), 3), dtype=np.float32)
        r = np.zeros(int(np_upper), dtype=np.float32)
        lib().ref_fill_probe_as_rank.restype = C.c_int64
        n = lib().ref_fill_probe_as_rank(self._h, C.c_int(int(as_rank)), C.c_int64(int(np_upper)), _p(q), _p(r))
        return q[:n].copy(), r

    # ---- per-kernel
    def paint(self, x, which=0, a=1.0):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = self._buf(which, a)
        lib().ref_paint(self._h, C.c_int(which), C.c_double(a), _p(x), C.c_int64(len(x)), _p(out))
        return out

    def readout(self, canvas, x, which=0, a=1.0):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros(len(x), dtype=np.float32)
        lib().ref_readout(self._h, C.c_int(which), C.c_double(a), _p(np.ascontiguousarray(canvas, dtype=np.float32)),
                          _p(x), C.c_int64(len(x)), _p(out))
        return out

    def paint_window(self, x, window, support=0, which=0, a=1.0, diffdir=-1):
        """window: "linear", "quad", "lanczos" (painter.c:128-174); support only matters for lanczos.  diffdir >= 0: through a
        derivative painter (fastpm_painter_init_diff), "cic" allowed then."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = self._buf(which, a)
        if diffdir >= 0:
            lib().ref_paint_window_diff(self._h, C.c_int(which), C.c_double(a), C.c_int(WINDOWS[window]), C.c_int(support), C.c_int(diffdir),
                                        _p(x), C.c_int64(len(x)), _p(out))
            return out
        lib().ref_paint_window(self._h, C.c_int(which), C.c_double(a), C.c_int(WINDOWS[window]), C.c_int(support), _p(x), C.c_int64(len(x)), _p(out))
        return out

    def readout_window(self, canvas, x, window, support=0, which=0, a=1.0, diffdir=-1):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros(len(x), dtype=np.float32)
        if diffdir >= 0:
            lib().ref_readout_window_diff(self._h, C.c_int(which), C.c_double(a), C.c_int(WINDOWS[window]), C.c_int(support), C.c_int(diffdir),
                                          _p(np.ascontiguousarray(canvas, dtype=np.float32)), _p(x), C.c_int64(len(x)), _p(out))
            return out
        lib().ref_readout_window(self._h, C.c_int(which), C.c_double(a), C.c_int(WINDOWS[window]), C.c_int(support),
                                 _p(np.ascontiguousarray(canvas, dtype=np.float32)), _p(x), C.c_int64(len(x)), _p(out))
        return out

    def r2c(self, real_buf, which=0, a=1.0):
        out = self._buf(which, a)
        lib().ref_r2c(self._h, C.c_int(which), C.c_double(a), _p(np.ascontiguousarray(real_buf, dtype=np.float32)), _p(out))
        return out

    def c2r(self, cbuf, which=0, a=1.0):
        out = self._buf(which, a)
        lib().ref_c2r(self._h, C.c_int(which), C.c_double(a), _p(np.ascontiguousarray(cbuf, dtype=np.float32)), _p(out))
        return out

    def kernel_transfer(self, delta_k, memb, attr=0, which=0, a=1.0):
        out = self._buf(which, a)
        lib().ref_kernel_transfer(self._h, C.c_int(which), C.c_double(a), _p(np.ascontiguousarray(delta_k, dtype=np.float32)),
                                  C.c_int(attr), C.c_int(memb), _p(out))
        return out

    def decic(self, delta_k, which=0, a=1.0):
        out = self._buf(which, a)
        lib().ref_decic(self._h, C.c_int(which), C.c_double(a), _p(np.ascontiguousarray(delta_k, dtype=np.float32)), _p(out))
        return out

    def laplace_diff(self, delta_k, d1=-1, d2=-1, which=1, a=1.0):
        out = self._buf(which, a)
        lib().ref_laplace_diff(self._h, C.c_int(which), C.c_double(a), _p(np.ascontiguousarray(delta_k, dtype=np.float32)),
                               C.c_int(d1), C.c_int(d2), _p(out))
        return out

    def powerspectrum(self, delta_k, which=0, a=1.0):
        n = self.pm_info(which, a)["Nmesh"] // 2
        k, p, nm = np.zeros(n), np.zeros(n), np.zeros(n)
        lib().ref_powerspectrum(self._h, C.c_int(which), C.c_double(a), _p(np.ascontiguousarray(delta_k, dtype=np.float32)),
                                _p(k), _p(p), _p(nm))
        return k, p, nm

    def cross_powerspectrum(self, delta1_k, delta2_k, which=0, a=1.0):
        n = self.pm_info(which, a)["Nmesh"] // 2
        k, p, nm = np.zeros(n), np.zeros(n), np.zeros(n)
        lib().ref_cross_powerspectrum(self._h, C.c_int(which), C.c_double(a), _p(np.ascontiguousarray(delta1_k, dtype=np.float32)),
                                      _p(np.ascontiguousarray(delta2_k, dtype=np.float32)), _p(k), _p(p), _p(nm))
        return k, p, nm

    def compute_force(self, a, want_delta_k=False):
        dk = self._buf(0, a) if want_delta_k else None
        lib().ref_compute_force(self._h, C.c_double(a), _p(dk))
        return dk

    def kick_factor(self, ai, ac, af):
        o = np.zeros(101)
        lib().ref_kick_factor(self._h, C.c_double(ai), C.c_double(ac), C.c_double(af), _p(o))
        return dict(ai=o[0], ac=o[1], af=o[2], q1=o[3], q2=o[4], dda=o[5:37].copy(), Dv1=o[37:69].copy(), Dv2=o[69:101].copy())

    def drift_factor(self, ai, ac, af):
        o = np.zeros(101)
        lib().ref_drift_factor(self._h, C.c_double(ai), C.c_double(ac), C.c_double(af), _p(o))
        return dict(ai=o[0], ac=o[1], af=o[2], Dv1=o[3], Dv2=o[4], dyyy=o[5:37].copy(), da1=o[37:69].copy(), da2=o[69:101].copy())

    def kick(self, ai, ac, af):
        lib().ref_kick(self._h, C.c_double(ai), C.c_double(ac), C.c_double(af))

    def drift(self, ai, ac, af):
        lib().ref_drift(self._h, C.c_double(ai), C.c_double(ac), C.c_double(af))

    def growth(self, a):
        o = np.zeros(12)
        lib().ref_growth(self._h, C.c_double(a), _p(o))
        return dict(zip(["D1", "D2", "f1", "f2", "E", "dEda", "d2Eda2", "dD1da", "d2D1da2", "Omega_a", "Omega_Lambda", "Omega_cdm"], o))


def clock_table():
    """Cumulative seconds of the reference's named clocks (CLOCK / ENTER / LEAVE, api/fastpm/prof.h:24-28) as printed by
    fastpm_clock_stat (prof.c:144): {"func:name": seconds}, one rank, so min = max = mean."""
    L = lib()
    buf = C.create_string_buffer(1 << 16)
    L.ref_clock_table.argtypes = [C.c_char_p, C.c_int]
    L.ref_clock_table(buf, len(buf))
    import re
    out = {}
    for m in re.finditer(r"(-?\d+\.\d+)\s+(-?\d+\.\d+)\s+(-?\d+\.\d+)\s+(\w+) : (\w+) :", buf.value.decode(errors="replace")):
        out["%s:%s" % (m.group(5), m.group(4))] = float(m.group(3))
    return out


def schedule(time_step):
    ts = np.ascontiguousarray(time_step, dtype=np.float64)
    rows = np.zeros((5 * len(ts) + 8, 7))
    n = lib().ref_schedule(_p(ts), C.c_int(len(ts)), _p(rows), C.c_int(len(rows)))
    return rows[:n]
