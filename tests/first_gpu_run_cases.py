"""GPU parity cases of the rows either side of the force step.  They are ordinary pytest tests; the file
name keeps them out of the default collection because tests/test_zz_first_gpu_run.py runs each one in its own pytest process (an
abort inside the library then fails that case only).  All of them were green on a B200 at the end of round 1 and run as plain
(non-xfail) tests since round 2; tests/test_cpu_full_emulation.py runs the same cases on the CPU against the emulated library.

  * device initial-condition generator (row N1): fpm_fill_gaussian_gadget against the reference's fastpm_ic_fill_gaussiank; the
    per-column arithmetic is the same source the CPU test runs bit for bit against the oracle, on the device only the last bit of
    double sin / cos / log may differ
  * the plain-C libfastpm user programs of tests/abi/
  * the opt-in one-pass readout of the three force components
  * PGD correction (N3), snapshot files + restart from the device (N2), force softening and the other windows (N4)
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
# tests/test_cpu_full_emulation.py runs these cases on the CPU against the emulated library with a smaller particle grid
NC_SCALE = float(os.environ.get("FASTPM_B200_TEST_NC_SCALE", "1"))


def _nc(n):
    return max(8, int(round(n * NC_SCALE / 4)) * 4)


@pytest.mark.parametrize("n,seed", [(32, 2004), (48, 100)])
def test_device_gadget_ic_matches_reference(ref_mod, n, seed):
    from fastpm_b200 import device
    device._lib.require_device()
    s = ref_mod.Session(nc=n, boxsize=100.0, pm_nc_factor=1)
    want = s.complex_view(s.fill_gaussian(seed), which=1)
    s.close()
    m = device.Mesh(n, 100.0)
    buf = m.alloc()
    m.fill_gaussian_gadget(buf, seed)
    got = m.download_complex(buf)
    # float32(double expression): device libm is within 1-2 ulp in double, so a float result differs only when the double
    # lands on a float rounding boundary -- allow one float ulp on at most 0.1 % of the values, exact equality elsewhere
    a, b = got.view(np.float32), want.view(np.float32)
    diff = a != b
    assert diff.mean() < 1e-3, diff.mean()
    assert np.abs(a - b).max() <= 2e-7 * max(1.0, np.abs(b).max())
    assert abs((np.abs(got) ** 2).mean() - (np.abs(want) ** 2).mean()) < 1e-6


def test_libfastpm_user_program_runs_and_matches_fixture(tmp_path):
    """tests/abi/dropin_example.c (plain C against the libfastpm API names, linked with libfastpm_b200.so) reproduces the
    committed reference fixture: 2LPT + 3 KDK cycles, positions within 1e-4 Mpc/h."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    from test_abi_layout import build_dropin_example
    fx = np.load(os.path.join(here, "golden", "small_run.npz"))
    exe = build_dropin_example(str(tmp_path))
    dk, ts = str(tmp_path / "dk.f32"), str(tmp_path / "ts.f64")
    np.ascontiguousarray(fx["delta_k"], dtype=np.float32).tofile(dk)
    np.ascontiguousarray(fx["steps"], dtype=np.float64).tofile(ts)
    out_x, out_pk = str(tmp_path / "x.f64"), str(tmp_path / "pk.f64")
    r = subprocess.run([exe, dk, ts, out_x, out_pk], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    L = 32.0
    x = np.fromfile(out_x, dtype=np.float64).reshape(-1, 3)
    d = np.abs(np.mod(x, L) - np.mod(fx["x1"], L))
    assert np.minimum(d, L - d).max() < 1e-4
    pk = np.fromfile(out_pk, dtype=np.float64)
    assert len(pk) == 2 * 16 and np.isfinite(pk).all()
    want_p, want_k = fx["pk_p"][-1], fx["pk_k"][-1]
    np.testing.assert_allclose(pk[:16], want_k, rtol=1e-9)
    np.testing.assert_allclose(pk[16:], want_p, rtol=1e-5)        # the tolerance BASELINE.json states for P(k)


def test_fused_readout_option_gives_the_same_run(tmp_path):
    """FASTPM_B200_FUSED_READOUT=1 (three inverse transforms resident, one pass over the particles) is the same run as the
    default one-canvas loop of gravity.c:359-396."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    from test_abi_layout import build_dropin_example
    fx = np.load(os.path.join(here, "golden", "small_run.npz"))
    exe = build_dropin_example(str(tmp_path))
    dk, ts = str(tmp_path / "dk.f32"), str(tmp_path / "ts.f64")
    np.ascontiguousarray(fx["delta_k"], dtype=np.float32).tofile(dk)
    np.ascontiguousarray(fx["steps"], dtype=np.float64).tofile(ts)
    res = []
    for fused in ("0", "1"):
        out_x, out_pk = str(tmp_path / ("x%s.f64" % fused)), str(tmp_path / ("pk%s.f64" % fused))
        env = dict(os.environ, FASTPM_B200_FUSED_READOUT=fused)
        r = subprocess.run([exe, dk, ts, out_x, out_pk], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300, env=env)
        assert r.returncode == 0, r.stdout
        res.append((np.fromfile(out_x, dtype=np.float64), np.fromfile(out_pk, dtype=np.float64)))
    # the readout itself is bit-identical (test_three_component_readout_equals_three_readouts); two RUNS differ by the order of
    # the float reductions of the deposit, far below the 1e-4 Mpc/h / 1e-5 tolerances of the path
    d = np.abs(res[0][0] - res[1][0])
    assert np.minimum(d, 32.0 - d).max() < 1e-5
    np.testing.assert_allclose(res[0][1], res[1][1], rtol=1e-6)


def test_three_component_readout_equals_three_readouts():
    """fpm_readout3 (the FASTPM_B200_FUSED_READOUT path of fastpm_solver_compute_force) == three fpm_readout calls, bit for bit."""
    from fastpm_b200 import device as dev
    dev._lib.require_device()
    nmesh, L, npart = 64, 64.0, 50000
    rng = np.random.default_rng(77)
    x = rng.uniform(0, L, size=(npart, 3))
    x[: npart // 4] = np.round(x[: npart // 4] / (L / 8)) * (L / 8)      # on cell faces, including x == L
    x[0], x[1] = [L, L, L], [0, 0, 0]
    m = dev.Mesh(nmesh, L)
    canv = []
    for d in range(3):
        c = m.alloc()
        m.upload_real(c, rng.normal(size=(nmesh, nmesh, nmesh)).astype(np.float32))
        canv.append(c)
    xd = dev.DeviceBuffer.from_host(x)
    sep = dev.DeviceBuffer(12 * npart)
    one = dev.DeviceBuffer(12 * npart)
    for d in range(3):
        m.readout(canv[d], xd, npart, sep, out_stride=3, out_offset_bytes=4 * d)
    m.readout3(canv, xd, npart, one)
    assert np.array_equal(sep.download(np.float32), one.download(np.float32))


def test_pgd_correction_matches_reference(ref_mod, pk_text):
    """Row N3 of SURVEY.md section 8f: the PGD correction (pgdcorrection.c) switched on -- the pgdc column after every force and
    the extra displacement in every drift (factors.c:108-113) -- against the oracle on identical initial conditions."""
    from fastpm_b200.solver import Solver
    nc, L, B = _nc(32), 2.0 * _nc(32), 2
    par = (0.2, 0.5, 1.0, 1.0, 5.0)                      # alpha0, A, B, kl, ks
    steps = np.linspace(0.1, 1.0, 5)
    s = ref_mod.Session(nc=nc, boxsize=L, pm_nc_factor=B, force_mode="fastpm", growth_mode="LCDM", np_alloc_factor=2.0, pgdc=par)
    dk, _, _ = s.ic_deltak(100, pk_text)
    s.setup_lpt(dk, steps[0])
    s.evolve(steps)
    want, want_pgdc = s.get_particles(), s.get_pgdc()
    s.close()
    # the same run without the correction, to show that the comparison is sensitive to it
    s = ref_mod.Session(nc=nc, boxsize=L, pm_nc_factor=B, force_mode="fastpm", growth_mode="LCDM", np_alloc_factor=2.0)
    s.setup_lpt(dk, steps[0])
    s.evolve(steps)
    plain = s.get_particles()
    s.close()
    g = Solver(nc=nc, boxsize=L, pm_nc_factor=B, force_mode="fastpm", growth_mode="LCDM", np_alloc_factor=2.0, pgdc=par)
    g.setup_lpt(dk, steps[0])
    g.evolve(steps)
    x, v, pg = g.get_column("x"), g.get_column("v"), g.get_column("pgdc")
    g.close()

    def pdist(a, b):
        d = np.abs(np.mod(a, L) - np.mod(b, L))
        return np.minimum(d, L - d).max()
    assert pdist(want["x"], plain["x"]) > 1e-2           # the correction moves particles by much more than the tolerance
    assert pdist(x, want["x"]) < 1e-4                    # Mpc/h, BASELINE.json north_star
    assert np.abs(v - want["v"]).max() < 1e-4 * np.abs(want["v"]).max()
    assert np.abs(pg - want_pgdc).max() < 1e-4 * np.abs(want_pgdc).max()


def _read_block(top, name, dtype, nmemb):
    import os
    a = np.fromfile(os.path.join(top, "1", name, "000000"), dtype=dtype)
    return a.reshape(-1, nmemb) if nmemb > 1 else a


def test_snapshot_files_and_restart_match_reference(ref_mod, pk_text, tmp_path):
    """Row N2 of SURVEY.md section 8f from the device: fastpm_b200_write_snapshot (unit conversion, header, catalog) against the
    directory the reference writes for its own run, the reference reading ours, and a restart from the reference's snapshot on
    both sides (src/fastpm.c:618-635)."""
    import filecmp
    import os
    from fastpm_b200.solver import Solver
    nc, L, B = _nc(16), 2.0 * _nc(16), 2
    kw = dict(nc=nc, boxsize=L, pm_nc_factor=B, force_mode="fastpm", growth_mode="LCDM", np_alloc_factor=2.0)
    first, second = np.linspace(0.1, 0.5, 3), np.linspace(0.5, 1.0, 3)
    s = ref_mod.Session(**kw)
    dk, _, _ = s.ic_deltak(7, pk_text)
    s.setup_lpt(dk, first[0])
    s.evolve(first)
    ref_dir, mine = str(tmp_path / "ref_0.5000"), str(tmp_path / "mine_0.5000")
    s.write_snapshot(ref_dir)
    s.close()
    g = Solver(**kw)
    g.setup_lpt(dk, first[0])
    g.evolve(first)
    v_before = g.get_column("v")
    g.write_snapshot(mine)
    assert np.array_equal(g.get_column("v"), v_before) or np.abs(g.get_column("v") - v_before).max() < 1e-6 * np.abs(v_before).max()
    g.close()
    # same files; headers and attributes identical (the growth numbers of "Header" to 2e-7, see tests/test_cpu_snapshot_io.py)
    files = sorted(os.path.relpath(os.path.join(d, f), mine) for d, _, fs in os.walk(mine) for f in fs)
    assert files == sorted(os.path.relpath(os.path.join(d, f), ref_dir) for d, _, fs in os.walk(ref_dir) for f in fs)
    assert filecmp.cmp(os.path.join(mine, "1", "attr-v2"), os.path.join(ref_dir, "1", "attr-v2"), shallow=False)
    assert filecmp.cmp(os.path.join(mine, "1", "ID", "000000"), os.path.join(ref_dir, "1", "ID", "000000"), shallow=False)
    xa, xb = _read_block(mine, "Position", np.float32, 3), _read_block(ref_dir, "Position", np.float32, 3)
    d = np.abs(xa.astype(np.float64) - xb)
    assert np.minimum(d, L - d).max() < 1e-4
    va, vb = _read_block(mine, "Velocity", np.float32, 3), _read_block(ref_dir, "Velocity", np.float32, 3)
    assert np.abs(va - vb).max() < 1e-4 * np.abs(vb).max()
    # the reference reads our directory
    s = ref_mod.Session(**kw)
    assert s.read_snapshot(mine) == 0.5
    q = s.get_particles()
    assert np.array_equal(q["x"].astype(np.float32), xa) and np.array_equal(q["v"], va)
    s.close()
    # restart from the reference's snapshot on both sides
    s = ref_mod.Session(**kw)
    assert s.read_snapshot(ref_dir, restart=True) == 0.5
    s.evolve(second)
    want = s.get_particles()
    s.close()
    g = Solver(**kw)
    assert g.read_snapshot(ref_dir) == 0.5
    assert g.meta["a_x"] == 0.5 and g.meta["a_v"] == 0.5
    g.evolve(second)
    x, v = g.get_column("x"), g.get_column("v")
    g.close()
    d = np.abs(np.mod(x, L) - np.mod(want["x"], L))
    assert np.minimum(d, L - d).max() < 1e-4
    assert np.abs(v - want["v"]).max() < 1e-4 * np.abs(want["v"]).max()


def test_wrap_and_summary_match_reference(ref_mod):
    """Rows a3 and a16 of SURVEY.md section 8 against the reference itself: fpm_wrap == fastpm_store_wrap (store.c:447-475) bit for bit
    (remainder() is exact), positions on the box faces, at exact multiples of the box and up to 9999 boxes away included; fpm_summary
    == fastpm_store_summary (store.c:808-909): extrema exact, mean and deviation to the order of summation."""
    from fastpm_b200 import device as dev
    lib = dev._lib.require_device()
    nc, L = 16, 50.0
    s = ref_mod.Session(nc=nc, boxsize=L, pm_nc_factor=2, force_mode="fastpm", np_alloc_factor=2.0)
    rng = np.random.default_rng(3)
    n = nc ** 3
    x = rng.uniform(-3 * L, 4 * L, size=(n, 3))
    x[0], x[1], x[2], x[3] = [L, -L, 0.0], [2 * L, -0.0, 9999.5 * L], [-9999.25 * L, 1e-300, -1e-300], [np.nextafter(L, 0), np.nextafter(L, 2 * L), L / 2]
    x[4:1000] = np.round(x[4:1000] / (L / 8)) * (L / 8)                    # many exact multiples of the cell size and of the box
    v = rng.normal(size=(n, 3)).astype(np.float32) * 3
    s.set_particles(x, v=v, id=np.arange(n, dtype=np.uint64))
    s.wrap()
    want = s.get_particles()["x"]
    xd = dev.DeviceBuffer.from_host(x)
    dev.check(lib.fpm_wrap(xd.ptr, n, L), "fpm_wrap")
    assert lib.fpm_wrap_check() == 0
    got = xd.download(np.float64).reshape(n, 3)
    assert np.array_equal(got, want) and np.array_equal(np.signbit(got), np.signbit(want))
    assert got.min() >= 0 and got.max() <= L
    vd = dev.DeviceBuffer.from_host(v)
    for col, buf, dtype in (("x", xd, np.float64), ("v", vd, np.float32)):
        a, b = dev.summary(buf, dtype, 3, n), s.summary(col)
        assert np.array_equal(a["min"], b["min"]) and np.array_equal(a["max"], b["max"]), col
        np.testing.assert_allclose(a["mean"], b["mean"], rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(a["std"], b["std"], rtol=1e-11)
    s.close()


@pytest.mark.parametrize("fraction", [0.3, 1.0, 0.0])
def test_store_subsample_matches_reference(ref_mod, fraction):
    """The command line's particle_fraction (src/fastpm.c:1449-1461) with the reference's own store calls: fastpm_store_fill (rand
    column) -> fastpm_store_fill_subsample_mask -> fastpm_store_get_mask_sum -> fastpm_store_subsample (count only, into a second
    store, in place) [-> fastpm_store_permute (reversed) -> fastpm_store_sort(FastPMLocalSortByID)]: the same particles, in the same
    order, as the reference keeps (store.c:289-299, 380-446, 967-1034)."""
    import ctypes as C
    from fastpm_b200.solver import Solver
    nc, L = 12, 60.0
    n_up = 2 * nc ** 3 + 17
    s = ref_mod.Session(nc=nc, boxsize=L, pm_nc_factor=2, np_alloc_factor=2.0)
    want = {(ip, sb): s.subsample_probe(n_up, fraction, in_place=ip, sort_back=sb) for ip in (False, True) for sb in (False, True)}
    s.close()
    g = Solver(nc=nc, boxsize=L, pm_nc_factor=2, np_alloc_factor=2.0)
    g.lib.fastpm_b200_subsample_probe.restype = C.c_int64
    g.lib.fastpm_b200_subsample_probe.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    for (ip, sb), (ids0, x0, ms0) in want.items():
        ids, x, ms = np.zeros(n_up, dtype=np.uint64), np.zeros((n_up, 3)), C.c_int64(0)
        n = g.lib.fastpm_b200_subsample_probe(g.lptpm, n_up, fraction, int(ip), int(sb), ids.ctypes.data, x.ctypes.data, C.byref(ms))
        assert n == len(ids0) == ms.value == ms0, (ip, sb)
        assert np.array_equal(ids[:n], ids0) and np.array_equal(x[:n], x0), (ip, sb)
        assert np.all(np.diff(ids0.astype(np.int64)) > 0)
    g.close()
    n = len(want[(False, False)][0])
    assert n == nc ** 3 if fraction >= 1 else (n <= 2 if fraction == 0 else 0.2 * nc ** 3 < n < 0.4 * nc ** 3)


def test_store_copy_take_extend(pk_text):
    """fastpm_store_copy / _take / _extend / _get_position (store.c:106-111, 925-966) between two stores with device columns."""
    import ctypes as C
    from fastpm_b200.solver import Solver, COLUMNS
    nc, L = 8, 32.0
    g = Solver(nc=nc, boxsize=L, pm_nc_factor=2, force_mode="fastpm", growth_mode="LCDM", np_alloc_factor=2.0)
    kk, pp = np.loadtxt(pk_text.splitlines(), unpack=True)
    g.setup_ic(5, kk, pp, 0.2)
    x, v, ids = g.get_column("x"), g.get_column("v"), g.get_column("id")
    n = g.np
    lib, vp = g.lib, C.c_void_p
    store = C.create_string_buffer(1 << 16)                   # room for one FastPMStore (its layout is the library's business)
    other = C.cast(store, vp)
    lib.fastpm_store_init_details.argtypes = [vp, C.c_char_p, C.c_size_t, C.c_uint64, C.c_int, C.c_char_p, C.c_int]
    lib.fastpm_store_init_details(other, b"other", 3 * n, COLUMNS["x"] | COLUMNS["v"] | COLUMNS["id"], 0, b"test", 0)
    for f in (lib.fastpm_store_copy, lib.fastpm_store_extend):
        f.argtypes = [vp, vp]
    lib.fastpm_store_take.argtypes = [vp, C.c_ssize_t, vp, C.c_ssize_t]
    lib.fastpm_store_get_position.argtypes = [vp, C.c_ssize_t, vp]
    lib.fastpm_store_destroy.argtypes = [vp]
    lib.fastpm_b200_store_np.argtypes = [vp]
    lib.fastpm_b200_store_np.restype = C.c_int64

    def column(name, count):
        dt, nm = {"x": (np.float64, 3), "v": (np.float32, 3), "id": (np.uint64, 1)}[name]
        out = np.empty((count, nm) if nm > 1 else (count,), dtype=dt)
        assert lib.fastpm_b200_store_get_column(other, COLUMNS[name], out.ctypes.data, 0, count) == 0
        return out

    lib.fastpm_store_copy(g.cdm, other)
    assert lib.fastpm_b200_store_np(other) == n
    assert np.array_equal(column("x", n), x) and np.array_equal(column("v", n), v) and np.array_equal(column("id", n), ids)
    lib.fastpm_store_extend(other, g.cdm)                     # other = cdm + cdm
    assert lib.fastpm_b200_store_np(other) == 2 * n
    assert np.array_equal(column("id", 2 * n), np.concatenate([ids, ids])) and np.array_equal(column("x", 2 * n), np.concatenate([x, x]))
    lib.fastpm_store_take(g.cdm, 7, other, 2 * n)             # one more particle at the end
    assert lib.fastpm_b200_store_np(other) == 2 * n + 1
    assert np.array_equal(column("x", 2 * n + 1)[-1], x[7]) and column("id", 2 * n + 1)[-1] == ids[7]
    pos = np.zeros(3)
    lib.fastpm_store_get_position(other, 2 * n, pos.ctypes.data)
    assert np.array_equal(pos, x[7])
    lib.fastpm_store_destroy(other)
    g.close()


def test_permute_by_dense_id_kernels():
    """fpm_id_order_counts / fpm_permute_by_id (the sort by a dense particle id of fastpm_sort_snapshot, libfastpmio/io.c:860-960, as
    one scatter per column): rows of 24, 12, 8 and 1 bytes, a non-zero first id, ids outside the range, duplicates."""
    import ctypes as C
    from fastpm_b200 import device as dev
    lib = dev._lib.require_device()
    rng = np.random.default_rng(5)
    n, id0 = 70001, 12345

    def counts(ids, first):
        d, out = dev.DeviceBuffer.from_host(ids), np.zeros(2, dtype=np.uint64)
        dev._lib.check(lib.fpm_id_order_counts(d.ptr, len(ids), first, out.ctypes.data), "fpm_id_order_counts")
        return int(out[0]), int(out[1])

    perm = rng.permutation(n).astype(np.uint64)
    ids = perm + np.uint64(id0)
    assert counts(np.arange(n, dtype=np.uint64) + np.uint64(id0), id0) == (0, 0)
    assert counts(ids, id0) == (0, int((perm != np.arange(n)).sum()))
    assert counts(ids, 0)[0] == int((ids >= n).sum())                       # ids beyond the range
    assert counts(ids, id0 + 7)[0] == 7                                     # ids below the first one wrap around
    assert counts(np.zeros(0, dtype=np.uint64), 0) == (0, 0)
    idd = dev.DeviceBuffer.from_host(ids)
    for dtype, nm in ((np.float64, 3), (np.float32, 3), (np.uint64, 1), (np.uint8, 1), (np.uint8, 3)):
        col = rng.integers(0, 250, size=(n, nm)).astype(dtype) if dtype == np.uint8 else rng.normal(size=(n, nm)).astype(dtype)
        if dtype == np.uint64:
            col = ids.reshape(n, 1).copy()
        src, dst = dev.DeviceBuffer.from_host(col), dev.DeviceBuffer(col.nbytes)
        dev._lib.check(lib.fpm_permute_by_id(dst.ptr, src.ptr, idd.ptr, n, id0, col.itemsize * nm), "fpm_permute_by_id")
        want = np.empty_like(col)
        want[perm.astype(np.int64)] = col
        assert np.array_equal(dst.download(dtype).reshape(n, nm), want)
        if dtype == np.uint64:
            assert counts(want.ravel(), id0) == (0, 0)
    # a duplicate id leaves one slot of the scattered id column unwritten: it shows up as displaced
    dup = ids.copy()
    dup[17] = dup[4711]
    scat = dev.DeviceBuffer(8 * n)
    dev._lib.check(lib.fpm_memset(scat.ptr, 0xff, 8 * n), "fpm_memset")
    dupd = dev.DeviceBuffer.from_host(dup)
    dev._lib.check(lib.fpm_permute_by_id(scat.ptr, dupd.ptr, dupd.ptr, n, id0, 8), "fpm_permute_by_id")
    c = counts(scat.download(np.uint64), id0)
    assert c[0] == 1 and c[1] == 1
    assert lib.fpm_permute_by_id(scat.ptr, scat.ptr, dupd.ptr, n, id0, 8) != 0          # in place is refused


def test_sorted_snapshot_of_a_shuffled_store(pk_text, tmp_path):
    """fastpm_sort_snapshot(FastPMSnapshotSortByID) on one rank (libfastpmio/io.c:860-960; the command line's default sort_snapshot):
    a store in arbitrary order is written as the same catalog as the store in id order -- by the device scatter (dense ids) and,
    with FASTPM_B200_HOST_SORT=1 or ids that are not dense, by the radix sort on the host; the store itself ends up in id order."""
    import filecmp
    import os
    from fastpm_b200.solver import Solver
    nc, L = _nc(16), 2.0 * _nc(16)
    g = Solver(nc=nc, boxsize=L, pm_nc_factor=2, force_mode="fastpm", growth_mode="LCDM", np_alloc_factor=2.0)
    kk, pp = np.loadtxt(pk_text.splitlines(), unpack=True)
    g.setup_ic(11, kk, pp, 0.1)
    g.evolve(np.linspace(0.1, 0.3, 2))
    names = [c for c in ("x", "v", "id", "dx1", "dx2", "acc") if g.column_ptr(c)]
    assert set(names) >= {"x", "v", "id", "acc"}
    cols = {c: g.get_column(c) for c in names}
    n = g.np
    assert np.array_equal(cols["id"], np.arange(n, dtype=np.uint64))
    plain = str(tmp_path / "plain")
    g.write_snapshot(plain)
    perm = np.random.default_rng(3).permutation(n)
    blocks = [("Position", np.float32, 3), ("Velocity", np.float32, 3), ("ID", np.uint64, 1)]

    launches = {}

    def shuffled_run(tag, env_host, break_density=False):
        for c in names:
            g.set_column(c, cols[c][perm])
        if break_density:                     # ids 0, 2, 4, ...: sorted order is unchanged, but the scatter does not apply
            g.set_column("id", (cols["id"] * np.uint64(2))[perm])
        out = str(tmp_path / tag)
        old = os.environ.pop("FASTPM_B200_HOST_SORT", None)
        if env_host:
            os.environ["FASTPM_B200_HOST_SORT"] = "1"
        before = g.lib.fpm_kernel_launch_count()
        try:
            g.write_snapshot(out, sort_by_id=True)
        finally:
            launches[tag] = g.lib.fpm_kernel_launch_count() - before
            os.environ.pop("FASTPM_B200_HOST_SORT", None)
            if old is not None:
                os.environ["FASTPM_B200_HOST_SORT"] = old
        want_id = cols["id"] * np.uint64(2) if break_density else cols["id"]
        assert np.array_equal(g.get_column("id"), want_id)
        for c in names[3:]:
            assert np.array_equal(g.get_column(c), cols[c]), c
        assert np.array_equal(np.mod(g.get_column("x"), L), np.mod(cols["x"], L))      # x comes back wrapped into the box
        return out

    dev_dir, host_dir, sparse_dir = shuffled_run("dev", False), shuffled_run("host", True), shuffled_run("sparse", False, True)
    # the dense ids went through the scatter kernels (two order checks, the trial scatter of the ids, one scatter per column); the
    # other two runs were sorted on the host
    assert launches["dev"] - launches["host"] >= 3 + len(names) and launches["sparse"] - launches["host"] == 1, launches
    for name, dtype, nm in blocks:
        a = _read_block(dev_dir, name, dtype, nm)
        assert np.array_equal(a, _read_block(host_dir, name, dtype, nm)), name
        if name != "ID":
            assert np.array_equal(a, _read_block(sparse_dir, name, dtype, nm)), name
        if name != "Velocity":                # the plain snapshot's unit conversion and its inverse may move the last bit of v
            assert np.array_equal(a, _read_block(plain, name, dtype, nm)), name
    va, vb = _read_block(dev_dir, "Velocity", np.float32, 3), _read_block(plain, "Velocity", np.float32, 3)
    assert np.abs(va - vb).max() <= 1e-6 * np.abs(vb).max()
    assert filecmp.cmp(os.path.join(dev_dir, "1", "attr-v2"), os.path.join(host_dir, "1", "attr-v2"), shallow=False)
    # a catalog in id order read back on one rank is in Lagrangian order again: the deposit / gather get their brick walk back
    # (a speed hint only); a catalog in any other order leaves the hint alone
    assert g.lib.fpm_particle_grid_hint_get() == nc
    for c in names:
        g.set_column(c, cols[c][perm])
    unsorted_dir = str(tmp_path / "unsorted")
    g.write_snapshot(unsorted_dir)
    g.close()
    for catalog, want in ((unsorted_dir, 0), (dev_dir, nc)):
        g = Solver(nc=nc, boxsize=L, pm_nc_factor=2, force_mode="fastpm", growth_mode="LCDM", np_alloc_factor=2.0)
        g.lib.fpm_particle_grid_hint(0)
        g.read_snapshot(catalog)
        assert g.lib.fpm_particle_grid_hint_get() == want, catalog
        assert np.array_equal(g.get_column("id"), cols["id"][perm] if want == 0 else cols["id"])
        g.close()


@pytest.mark.parametrize("softening", ["gaussian", "two_third", "gaussian36"])
def test_force_softening_matches_reference(ref_mod, pk_text, softening):
    """Row N4 (softening kernels, gravity.c:244-270): a short run with the dealiasing sweep on delta_k switched on."""
    from fastpm_b200.solver import Solver
    nc, L, B = _nc(16), 2.0 * _nc(16), 2
    kw = dict(nc=nc, boxsize=L, pm_nc_factor=B, force_mode="fastpm", growth_mode="LCDM", np_alloc_factor=2.0, softening=softening)
    steps = np.linspace(0.1, 1.0, 4)
    s = ref_mod.Session(**kw)
    dk, _, _ = s.ic_deltak(100, pk_text)
    s.setup_lpt(dk, steps[0])
    s.evolve(steps)
    want = s.get_particles()
    s.close()
    kw0 = dict(kw, softening="none")
    s = ref_mod.Session(**kw0)
    s.setup_lpt(dk, steps[0])
    s.evolve(steps)
    plain = s.get_particles()
    s.close()
    g = Solver(**kw)
    g.setup_lpt(dk, steps[0])
    g.evolve(steps)
    x, v = g.get_column("x"), g.get_column("v")
    g.close()

    def pdist(a, b):
        d = np.abs(np.mod(a, L) - np.mod(b, L))
        return np.minimum(d, L - d).max()
    assert pdist(want["x"], plain["x"]) > 1e-3           # the sweep matters
    assert pdist(x, want["x"]) < 1e-4
    assert np.abs(v - want["v"]).max() < 1e-4 * np.abs(want["v"]).max()


def test_store_fill_q_and_rand_columns_match_reference(ref_mod):
    """fastpm_store_fill with the q and rand columns allocated (store.c:723-806): q is the grid position rounded to float, rand this
    rank's serial RANLUX stream over all np_upper entries (_fastpm_store_fill_rand, store.c:694-720) -- bit for bit."""
    import ctypes as C
    from fastpm_b200.solver import Solver
    nc, L = 12, 60.0
    s = ref_mod.Session(nc=nc, boxsize=L, pm_nc_factor=2, np_alloc_factor=2.0)
    n_up = 2 * nc ** 3 + 17
    q0, r0 = s.fill_probe(n_up)
    s.close()
    g = Solver(nc=nc, boxsize=L, pm_nc_factor=2, np_alloc_factor=2.0)
    q1 = np.zeros((n_up, 3), dtype=np.float32)
    r1 = np.zeros(n_up, dtype=np.float32)
    g.lib.fastpm_b200_fill_probe.restype = C.c_int64
    g.lib.fastpm_b200_fill_probe.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    n = g.lib.fastpm_b200_fill_probe(g.lptpm, n_up, q1.ctypes.data, r1.ctypes.data)
    g.close()
    assert n == nc ** 3 == len(q0)
    assert np.array_equal(q1[:n], q0)
    assert np.array_equal(r1, r0) and r0.min() >= 0 and r0.max() < 1 and len(np.unique(r0)) > 0.99 * n_up
    # the streams of the other ranks (seed = 0x7fffffff times the (8 r)-th deviate of the fixed seed's stream): the device-layer call
    # against the reference filling the same store while its MPI stub answers as rank r
    from fastpm_b200 import device
    s = ref_mod.Session(nc=nc, boxsize=L, pm_nc_factor=2, np_alloc_factor=2.0)
    for r in (1, 5):
        _, want = s.fill_probe(nc ** 3 + 300, as_rank=r)
        buf = device.DeviceBuffer(4 * len(want))
        device.check(buf.lib.fpm_fill_rand(buf.ptr, len(want), r), "fpm_fill_rand")
        got = buf.download(np.float32)
        assert np.array_equal(got, want), r
        assert not np.array_equal(got, r0[:len(want)])
    s.close()


def test_passive_handler_keeps_the_fused_update(pk_text):
    """Every event handler sees a fully updated store: the queued kicks and drifts are applied before it runs, which splits the
    fused K-K-D-D pass.  A handler marked passive (fastpm_b200_mark_handler_passive: it never reads the particles, like the
    command line's transition log) does not cost that; the run is the same in all three cases."""
    import ctypes as C
    from fastpm_b200 import _lib
    from fastpm_b200.solver import Solver, HANDLER
    nc, L = _nc(16), 2.0 * _nc(16)
    tab = np.array([[float(v) for v in l.split()] for l in pk_text.splitlines() if l.strip() and not l.startswith("#")])
    steps = np.linspace(0.1, 1.0, 5)
    names = ["paint", "readout", "fft_tile", "fft_z", "kick", "drift"]
    seen, keep = [], []

    def on_transition(solver_ptr, event_ptr, userdata):
        seen.append(1)
        return 0

    def run(mode):
        g = Solver(nc=nc, boxsize=L, pm_nc_factor=2, force_mode="cola", growth_mode="LCDM", np_alloc_factor=2.0)
        g.setup_ic(7, tab[:, 0], tab[:, 1], steps[0])
        lib = g.lib
        if mode != "none":
            cb = HANDLER(on_transition)
            keep.append(cb)               # the registry is keyed by the function's address: no thunk may be freed and reused here
            g._handlers.append(cb)
            if mode == "passive":
                lib.fastpm_b200_mark_handler_passive(cb)
            lib.fastpm_b200_add_handler(g.h, b"TRANSITION", 0, cb, None)
        lib.fpm_prof_reset()
        lib.fpm_prof_enable(1)
        g.evolve(steps)
        lib.fpm_prof_enable(0)
        counts, totals = (C.c_int64 * 16)(), (C.c_double * 16)()
        _lib.check(lib.fpm_prof_get(counts, totals, 16))
        x, v = g.get_column("x"), g.get_column("v")
        g.close()
        return {n: int(counts[i]) for i, n in enumerate(names)}, x, v

    c0, x0, v0 = run("none")
    n_before = len(seen)
    c1, x1, v1 = run("passive")
    assert len(seen) > n_before                                   # the passive handler was called
    c2, x2, v2 = run("active")
    assert c1["kick"] + c1["drift"] == c0["kick"] + c0["drift"], (c0, c1)
    assert c2["kick"] + c2["drift"] > c0["kick"] + c0["drift"], (c0, c2)
    # the same run in all three cases: the update kernels are bit-exact either way; on a GPU the deposit order of the float atomics
    # differs from run to run, so the comparison carries the usual tolerances (on the emulated library the runs are identical)
    for xa, va in ((x1, v1), (x2, v2)):
        d = np.abs(x0 - xa)
        assert np.minimum(d, L - d).max() < 1e-4
        assert np.abs(v0 - va).max() < 1e-4 * np.abs(v0).max()


def test_shifted_ics_match_reference(ref_mod, pk_text):
    """FastPMConfig.USE_SHIFT (solver.c:142-150,201-209; the Lua option `shift`): the particle grid sits at the cell centres, the 2LPT
    displacements are read out at the de-shifted positions (pm2lpt.c:30-34,141-145); ICs and a short run against the reference."""
    from fastpm_b200.solver import Solver
    nc, L = _nc(16), 2.0 * _nc(16)
    kw = dict(nc=nc, boxsize=L, pm_nc_factor=2, force_mode="fastpm", growth_mode="LCDM", np_alloc_factor=2.0, use_shift=True)
    steps = np.linspace(0.1, 1.0, 4)
    s = ref_mod.Session(**kw)
    dk, _, _ = s.ic_deltak(100, pk_text)
    s.setup_lpt(dk, steps[0])
    ic = s.get_particles()
    s.evolve(steps)
    want = s.get_particles()
    s.close()
    s = ref_mod.Session(**dict(kw, use_shift=False))
    s.setup_lpt(dk, steps[0])
    plain_ic = s.get_particles()
    s.close()
    g = Solver(**kw)
    g.setup_lpt(dk, steps[0])
    x0, v0 = g.get_column("x"), g.get_column("v")
    g.evolve(steps)
    x, v = g.get_column("x"), g.get_column("v")
    g.close()

    def pdist(a, b):
        d = np.abs(np.mod(a, L) - np.mod(b, L))
        return np.minimum(d, L - d).max()
    assert abs(pdist(ic["x"], plain_ic["x"]) - 0.5 * L / nc) < 0.05 * L / nc      # the grid really moved by half a particle spacing
    assert pdist(x0, ic["x"]) < 1e-5
    assert np.abs(v0 - ic["v"]).max() < 1e-5 * np.abs(ic["v"]).max()
    assert pdist(x, want["x"]) < 1e-4
    assert np.abs(v - want["v"]).max() < 1e-4 * np.abs(want["v"]).max()


@pytest.mark.parametrize("remove_variance", [False, True])
def test_device_ic_chain_matches_reference(ref_mod, pk_text, remove_variance):
    """Row N1 end to end: fastpm_b200_setup_gadget_ic (white noise, [remove_variance], colouring, 2LPT: all on the device) gives the
    particles the reference makes from the same seed and P(k) table (src/fastpm.c:415-545 + fastpm_solver_setup_lpt)."""
    import os
    from fastpm_b200.solver import Solver
    nc, L = _nc(32), 4.0 * _nc(32)
    kw = dict(nc=nc, boxsize=L, pm_nc_factor=2, force_mode="cola", growth_mode="LCDM", np_alloc_factor=2.0)
    s = ref_mod.Session(**kw)
    dk, _, _ = s.ic_deltak(2024, pk_text, remove_variance=remove_variance)
    s.setup_lpt(dk, 0.1)
    want = s.get_particles()
    s.close()
    tab = np.loadtxt(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "powerspec.txt"))
    g = Solver(**kw)
    g.setup_ic(2024, tab[:, 0], tab[:, 1], 0.1, remove_variance=remove_variance)
    x, v, dx1, dx2 = g.get_column("x"), g.get_column("v"), g.get_column("dx1"), g.get_column("dx2")
    assert np.array_equal(g.get_column("id"), want["id"])
    g.close()
    d = np.abs(x - want["x"])
    assert np.minimum(d, L - d).max() < 1e-5
    assert np.abs(v - want["v"]).max() < 1e-5 * max(1.0, np.abs(want["v"]).max())
    assert np.abs(dx1 - want["dx1"]).max() < 1e-5 * np.abs(want["dx1"]).max()
    assert np.abs(dx2 - want["dx2"]).max() < 1e-4 * np.abs(want["dx2"]).max()


@pytest.mark.parametrize("painter,support", [("quad", 3), ("lanczos", 4)])
def test_non_cic_painter_matches_reference(ref_mod, pk_text, painter, support):
    """Row N4 (windows, painter.c:17-125,217-317): the force step with a quadratic / Lanczos window instead of CIC, one GPU."""
    from fastpm_b200.solver import Solver
    nc, L, B = _nc(16), 2.0 * _nc(16), 2
    kw = dict(nc=nc, boxsize=L, pm_nc_factor=B, force_mode="fastpm", growth_mode="LCDM", np_alloc_factor=2.0, painter=painter, painter_support=support)
    steps = np.linspace(0.1, 1.0, 4)
    s = ref_mod.Session(**kw)
    dk, _, _ = s.ic_deltak(100, pk_text)
    s.setup_lpt(dk, steps[0])
    s.evolve(steps)
    want = s.get_particles()
    s.close()
    s = ref_mod.Session(**dict(kw, painter="cic", painter_support=2))
    s.setup_lpt(dk, steps[0])
    s.evolve(steps)
    plain = s.get_particles()
    s.close()
    g = Solver(**kw)
    g.setup_lpt(dk, steps[0])
    g.evolve(steps)
    x, v = g.get_column("x"), g.get_column("v")
    g.close()

    def pdist(a, b):
        d = np.abs(np.mod(a, L) - np.mod(b, L))
        return np.minimum(d, L - d).max()
    # The Lanczos window has negative lobes: the mesh is a sum of terms of both signs, so the order of the float additions
    # matters far more than for CIC.  The reference itself scatters by 2.8e-4 Mpc/h between two 8-thread runs of this very
    # four-step configuration (OpenMP atomics in arbitrary order; 1e-6 for CIC): the four-step comparison is held to twice that,
    # the window must change the run by at least ten times the tolerance (a run that ignored the window cannot pass), and a
    # two-entry time table (two force evaluations, one kick-drift-kick: no time for round-off to grow) is held to 2e-5.
    tol = 6e-4 if painter == "lanczos" else 1e-4
    assert pdist(want["x"], plain["x"]) > 10 * tol, pdist(want["x"], plain["x"])
    err = pdist(x, want["x"])
    print("%s window: four steps %.3g Mpc/h from the reference (window effect %.3g)" % (painter, err, pdist(want["x"], plain["x"])))
    assert err < tol, err
    assert np.abs(v - want["v"]).max() < tol * np.abs(want["v"]).max()
    short = steps[:2]
    s = ref_mod.Session(**kw)
    s.setup_lpt(dk, short[0])
    s.evolve(short)
    want2 = s.get_particles()
    s.close()
    g = Solver(**kw)
    g.setup_lpt(dk, short[0])
    g.evolve(short)
    err2 = pdist(g.get_column("x"), want2["x"])
    g.close()
    print("%s window: one kick-drift-kick %.3g Mpc/h from the reference" % (painter, err2))
    assert err2 < 2e-5, err2


def test_single_mode_transfers_match_reference(ref_mod):
    """fastpm_apply_set_mode / get_mode / normalize / c2r_weight transfers (transfer.c:223-366) on a device mesh, bit for bit."""
    import ctypes as C
    from fastpm_b200.solver import Solver
    n, L = 16, 50.0
    s = ref_mod.Session(nc=n, boxsize=L, pm_nc_factor=1)
    dk = s.fill_gaussian(5)
    c = s.complex_view(dk, which=1).copy()
    c[0, 0, 0] = 2.5                                    # a mean to normalise by
    dk = s.complex_pack(c, which=1)
    g = Solver(nc=n, boxsize=L, pm_nc_factor=1)
    lib, pm = g.lib, g.lptpm
    lib.fastpm_apply_get_mode_transfer.restype = C.c_double
    lib.fastpm_apply_get_mode_transfer.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.fastpm_apply_set_mode_transfer.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int]
    lib.fastpm_apply_normalize_transfer.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.fastpm_apply_c2r_weight_transfer.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    a, b = lib.pm_alloc_details(pm, b"test", 0), lib.pm_alloc_details(pm, b"test", 0)
    host = np.ascontiguousarray(dk, dtype=np.float32)
    out = np.zeros_like(host)

    def run(op, mode, value=0.0, method=0):
        assert lib.fastpm_b200_mesh_set_complex(pm, a, host.ctypes.data) == 0
        m = (C.c_ssize_t * 4)(*mode)
        if op == "set_mode":
            lib.fastpm_apply_set_mode_transfer(pm, a, b, m, value, method)
        elif op == "normalize":
            lib.fastpm_apply_normalize_transfer(pm, a, b)
        else:
            lib.fastpm_apply_c2r_weight_transfer(pm, a, b)
        r = lib.fastpm_apply_get_mode_transfer(pm, b, m)
        assert lib.fastpm_b200_mesh_get_complex(pm, b, out.ctypes.data) == 0
        return out.copy(), r
    cases = [("set_mode", (3, 5, 2, 0), 1.25, 0), ("set_mode", (3, 5, 2, 1), -0.5, 1), ("set_mode", (0, 8, 0, 1), 7.0, 0),
             ("set_mode", (2, 0, 0, 1), 0.75, 1), ("set_mode", (1, 2, 9, 0), 1.0, 0), ("normalize", (0, 0, 0, 0), 0, 0),
             ("c2r_weight", (8, 8, 8, 0), 0, 0), ("c2r_weight", (1, 1, 1, 1), 0, 0)]
    for op, mode, value, method in cases:
        want, wr = s.mode_op(op, dk, mode, value, method)
        got, gr = run(op, mode, value, method)
        assert gr == wr, (op, mode)
        assert np.array_equal(s.complex_view(got, which=1).view(np.float32), s.complex_view(want, which=1).view(np.float32)), (op, mode)
    lib.pm_free(pm, b)
    lib.pm_free(pm, a)
    g.close()
    s.close()


def test_snapshots_during_evolve_match_reference(ref_mod, pk_text, tmp_path):
    """The CLI's way of taking snapshots (check_snapshots, src/fastpm.c:1130-1208): at every requested aout inside a step the
    particles are drifted / kicked there with the INTERPOLATION event's factors, written, and put back.  Same directories as the
    reference's (data within the path's tolerances), same final state (the round trip perturbs it on both sides)."""
    import os
    from fastpm_b200.solver import Solver
    nc, L = _nc(16), 2.0 * _nc(16)
    kw = dict(nc=nc, boxsize=L, pm_nc_factor=2, force_mode="cola", growth_mode="LCDM", np_alloc_factor=2.0)
    steps, aout = np.linspace(0.1, 1.0, 5), [0.1, 0.37, 0.6, 1.0]
    s = ref_mod.Session(**kw)
    dk, _, _ = s.ic_deltak(7, pk_text)
    s.setup_lpt(dk, steps[0])
    s.evolve_snapshots(steps, str(tmp_path / "ref"), aout)
    want = s.get_particles()
    s.close()
    g = Solver(**kw)
    g.setup_lpt(dk, steps[0])
    g.add_snapshots(str(tmp_path / "mine"), aout)
    g.evolve(steps)
    x, v = g.get_column("x"), g.get_column("v")
    g.close()
    d = np.abs(np.mod(x, L) - np.mod(want["x"], L))
    assert np.minimum(d, L - d).max() < 1e-4
    assert np.abs(v - want["v"]).max() < 1e-4 * np.abs(want["v"]).max()
    names = sorted(n for n in os.listdir(tmp_path) if n.startswith("ref_"))
    assert names == ["ref_%0.04f" % a for a in aout]
    for n in names:
        a, b = str(tmp_path / n.replace("ref_", "mine_")), str(tmp_path / n)
        assert sorted(os.listdir(os.path.join(a, "1"))) == sorted(os.listdir(os.path.join(b, "1")))      # Position Velocity ID DX1 DX2 ...
        assert open(os.path.join(a, "1", "attr-v2")).read() == open(os.path.join(b, "1", "attr-v2")).read()
        xa, xb = _read_block(a, "Position", np.float32, 3), _read_block(b, "Position", np.float32, 3)
        dd = np.abs(xa.astype(np.float64) - xb)
        assert np.minimum(dd, L - dd).max() < 1e-4, n
        va, vb = _read_block(a, "Velocity", np.float32, 3), _read_block(b, "Velocity", np.float32, 3)
        assert np.abs(va - vb).max() < 1e-4 * np.abs(vb).max(), n
        assert np.array_equal(_read_block(a, "ID", np.uint64, 1), _read_block(b, "ID", np.uint64, 1))


def test_cli_run_loop_program_matches_reference(ref_mod, pk_text, tmp_path):
    """tests/abi/cli_like_example.c (reference API names only) run on the device: same snapshots and power-spectrum files as the
    reference driven the same way."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    from test_abi_layout import build_dropin_example
    exe = build_dropin_example(str(tmp_path), "cli_like_example")
    nc, L, seed, aout = _nc(16), 3.0 * _nc(16), 42, [0.1, 0.5, 1.0]
    r = subprocess.run([exe, os.path.join(here, "golden", "powerspec.txt"), str(tmp_path / "mine"), str(nc), str(L), str(seed),
                        ",".join("%g" % a for a in aout)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "3 snapshots" in r.stdout, r.stdout
    s = ref_mod.Session(nc=nc, boxsize=L, pm_nc_factor=2, force_mode="cola", growth_mode="LCDM", np_alloc_factor=2.0)
    dk, _, _ = s.ic_deltak(seed, pk_text)
    s.setup_lpt(dk, 0.1)
    s.evolve_snapshots(np.array([0.1, 0.325, 0.55, 0.775, 1.0]), str(tmp_path / "ref"), aout)
    recs = s.records()
    s.close()
    for a in aout:
        mine, ref_dir = str(tmp_path / ("mine_%0.04f" % a)), str(tmp_path / ("ref_%0.04f" % a))
        assert sorted(os.listdir(os.path.join(mine, "1"))) == sorted(os.listdir(os.path.join(ref_dir, "1")))
        xa, xb = _read_block(mine, "Position", np.float32, 3), _read_block(ref_dir, "Position", np.float32, 3)
        d = np.abs(xa.astype(np.float64) - xb)
        assert np.minimum(d, L - d).max() < 1e-4, a
        va, vb = _read_block(mine, "Velocity", np.float32, 3), _read_block(ref_dir, "Velocity", np.float32, 3)
        assert np.abs(va - vb).max() < 1e-4 * np.abs(vb).max(), a
    # the power spectrum files: "# k p N" rows against the reference's FORCE/after records
    for rec in recs:
        rows = np.loadtxt(str(tmp_path / ("mine_powerspec_%0.04f.txt" % rec["a_f"])))
        sel = rec["nmodes"] > 0
        assert np.array_equal(rows[:, 2], rec["nmodes"])
        np.testing.assert_allclose(rows[sel, 0], rec["k"][sel], rtol=1e-5)          # "%g": 6 significant digits
        np.testing.assert_allclose(rows[sel, 1], rec["p"][sel], rtol=2e-5)


def test_command_line_particle_fraction_snapshot(ref_mod, tmp_path):
    """particle_fraction = 0.25 through fastpm_b200_run (src/fastpm.c:1449-1461): the snapshot of the command line holds exactly the
    particles the reference's fastpm_store_fill_subsample_mask / fastpm_store_subsample keep on the same particle grid, sorted by
    id.  (On the CPU the same run goes through the emulated library: tests/test_cpu_full_emulation.py.)"""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    run_bin = os.path.join(root, "fastpm_b200", "lua_front", "_build", "fastpm_b200_run")
    if os.environ.get("FASTPM_B200_TEST_EMUL") or not os.path.exists(run_bin):
        pytest.skip("needs the command line built against the CUDA library")
    nc = 16
    shutil.copy(os.path.join(root, "tests", "golden", "powerspec.txt"), str(tmp_path / "powerspec.txt"))
    r = subprocess.run([run_bin, os.path.join(root, "tests", "lua", "small_nc16.lua"), str(nc), "3", "0.25"], cwd=str(tmp_path),
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    top = tmp_path / "out" / "fastpm_1.0000"
    ids = np.fromfile(str(top / "1" / "ID" / "000000"), dtype=np.uint64)
    s = ref_mod.Session(nc=nc, boxsize=2.0 * nc, pm_nc_factor=2, np_alloc_factor=3.0)
    want, _, _ = s.subsample_probe(3 * nc ** 3, 0.25)
    s.close()
    assert 0.15 * nc ** 3 < len(ids) < 0.35 * nc ** 3 and np.array_equal(ids, want)
    x = np.fromfile(str(top / "1" / "Position" / "000000"), dtype=np.float32).reshape(-1, 3)
    assert len(x) == len(ids) and np.isfinite(x).all() and x.min() >= 0 and x.max() <= 2.0 * nc
    assert "[ 0.25 ]" in open(str(top / "Header" / "attr-v2")).read()
