"""CPU parity of the particle KERNEL SOURCES against the oracle, bit for bit (tests/emul/particles_emul.cpp: csrc/paint.cu and
csrc/particles.cu compiled for the host).  The deposit runs in particle order there, which is the reference's own order with one
OpenMP thread, so even the float32 mesh is identical; readout, kick, drift, the fused K-K-D-D pass and wrap are exact anyway."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    if not os.path.exists(os.path.join(INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    exe = str(tmp_path_factory.mktemp("emul") / "particles_emul")
    env = dict(os.environ)
    env.pop("CXX", None)
    r = subprocess.run(["g++", "-O1", "-ffp-contract=off", "-std=c++17", "-pthread", "-w", "-I" + INC, "-o", exe,
                        os.path.join(ROOT, "tests", "emul", "particles_emul.cpp")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    assert r.returncode == 0, r.stdout

    def run(op, payload, tmp):
        fin, fout = os.path.join(tmp, op + ".in"), os.path.join(tmp, op + ".out")
        with open(fin, "wb") as f:
            f.write(payload)
        subprocess.run([exe, op, fin, fout], check=True)
        return open(fout, "rb").read()
    return run


@pytest.fixture(scope="module")
def one_thread_ref(ref_mod):
    old = os.environ.get("OMP_NUM_THREADS")
    os.environ["OMP_NUM_THREADS"] = "1"          # serial deposit order on the reference side
    # the variable is read once, when the OpenMP runtime starts: if an earlier test of this process has already run the reference,
    # the runtime itself has to be told
    import ctypes
    gomp, old_n = None, 0
    try:
        gomp = ctypes.CDLL("libgomp.so.1")
        old_n = gomp.omp_get_max_threads()
        gomp.omp_set_num_threads(1)
    except OSError:
        pass
    yield ref_mod
    if gomp is not None and old_n > 0:
        gomp.omp_set_num_threads(old_n)
    if old is None:
        os.environ.pop("OMP_NUM_THREADS", None)
    else:
        os.environ["OMP_NUM_THREADS"] = old


def _positions(rng, n, L):
    x = rng.uniform(0, L, size=(n, 3))
    x[: n // 8] = np.round(x[: n // 8] / (L / 16)) * (L / 16)        # on cell faces, including x == L
    x[0] = [L, L, L]
    x[1] = [0, 0, 0]
    x[2] = [np.nextafter(L, 0)] * 3
    return x


@pytest.mark.parametrize("vec", [0, 2, 4])
def test_paint_and_readout_kernels_bit_exact(emul, one_thread_ref, tmp_path, vec):
    nmesh, L, npart = 32, 100.0, 6000
    rng = np.random.default_rng(40 + vec)
    x = _positions(rng, npart, L)
    s = one_thread_ref.Session(nc=nmesh, boxsize=L, pm_nc_factor=1)
    want = s.real_view(s.paint(x)).copy()
    head = struct.pack("<iiiiddq", nmesh, 0, 0, vec, L, 1.0, npart)
    out = emul("paint", head + x.tobytes(), str(tmp_path))
    got = np.frombuffer(out, dtype=np.float32, count=nmesh ** 3).reshape(nmesh, nmesh, nmesh)
    # same cells, same double-precision weights, same order.  The reference adds the DOUBLE weight to the float cell
    # (painter-cic.c:25, one rounding); a float reduction rounds the weight first (two roundings): cells differ by <= 1 ulp per add
    assert np.array_equal(got != 0, want != 0)
    assert np.abs(got - want).max() <= 4e-7 * want.max()
    assert abs(got.sum(dtype=np.float64) - npart) < 1e-4
    # readout of a random mesh at the same positions
    field = rng.standard_normal((nmesh, nmesh, nmesh)).astype(np.float32)
    want_r = s.readout(s.real_pack(field), x)
    out = emul("readout", head + x.tobytes() + field.tobytes(), str(tmp_path))
    assert np.array_equal(np.frombuffer(out, dtype=np.float32), want_r)
    s.close()


def _lagrangian_positions(rng, nc, L, amp, nstray, wrapped=True):
    """nc^3 particles in fastpm_store_fill order on the grid, displaced by a smooth field of amplitude `amp` plus small noise
    (a brick of 8^3 stays compact), `nstray` of them thrown somewhere else in the box (they must take the global path)."""
    q = (np.indices((nc, nc, nc)).reshape(3, -1).T + 0.5) * (L / nc)
    ph = rng.uniform(0, 2 * np.pi, size=(3, 3))
    disp = np.stack([amp * np.sin(2 * np.pi * q[:, (d + 1) % 3] / L + ph[d, 0]) + 0.5 * amp * np.cos(4 * np.pi * q[:, (d + 2) % 3] / L + ph[d, 1])
                     for d in range(3)], axis=1)
    x = q + disp + rng.normal(scale=0.02 * L / nc, size=q.shape)
    stray = rng.choice(len(x), size=nstray, replace=False)
    x[stray] = rng.uniform(0, L, size=(nstray, 3))
    x[7] = [L, L, L]
    return np.mod(x, L) if wrapped else x


@pytest.mark.parametrize("factor,amp_cells", [(2, 1.5), (1, 1.5)])
def test_tile_kernels_against_reference(emul, one_thread_ref, tmp_path, factor, amp_cells):
    """cic_paint_tile_kernel / cic_readout_tile_kernel (csrc/paint.cu: 8^3 Lagrangian bricks, mesh box in shared memory, flushed /
    filled in aligned 16-byte groups) against the reference's paint and readout: the gather bit for bit, the deposit up to the order
    of the float32 additions; most particles must really go through the tiles, the strays through the global path."""
    nc, L = 16, 64.0
    nmesh = nc * factor
    rng = np.random.default_rng(70 + factor)
    x = _lagrangian_positions(rng, nc, L, amp_cells * L / nmesh, nstray=40)
    npart = len(x)
    s = one_thread_ref.Session(nc=nmesh, boxsize=L, pm_nc_factor=1)
    want = s.real_view(s.paint(x)).copy()
    head = struct.pack("<iiiiddqii", nmesh, nc, 0, 4, L, 1.0, npart, 1, 0)
    out = emul("tpaint", head + x.tobytes(), str(tmp_path))
    got = np.frombuffer(out, dtype=np.float32, count=nmesh ** 3).reshape(nmesh, nmesh, nmesh)
    stats = np.frombuffer(out[-16:], dtype=np.uint64)
    assert np.array_equal(got != 0, want != 0)
    assert np.abs(got - want).max() <= 2e-6 * want.max()              # a cell receives its float32 increments in another order
    assert abs(got.sum(dtype=np.float64) - npart) < 1e-3
    # the strays beyond the tile's reach (and little else) took the global path; on the 16-cell mesh every cell is within reach
    assert (5 if factor == 2 else 0) <= stats[0] <= 0.05 * npart and stats[1] == 0, stats
    field = rng.standard_normal((nmesh, nmesh, nmesh)).astype(np.float32)
    want_r = s.readout(s.real_pack(field), x)
    out = emul("treadout", head + x.tobytes() + field.tobytes(), str(tmp_path))
    assert np.array_equal(np.frombuffer(out[:-16], dtype=np.float32), want_r)
    s.close()


def test_tile_kernels_fused_wrap_and_unordered_store(emul, one_thread_ref, tmp_path):
    """The deposit with the periodic wrap folded in (positions outside [0, L) come back wrapped, store.c:447-475), and a store in
    random order: no brick is compact, every CTA gives up its tile and the result is still the reference's."""
    nc, L, nmesh = 16, 64.0, 32
    rng = np.random.default_rng(81)
    x = _lagrangian_positions(rng, nc, L, 1.5 * L / nmesh, nstray=10, wrapped=False)
    x[100:200] += L
    x[300:350] -= 2 * L
    npart = len(x)
    s = one_thread_ref.Session(nc=nmesh, boxsize=L, pm_nc_factor=1)
    xw = np.mod(x, L)
    want = s.real_view(s.paint(xw)).copy()
    head = struct.pack("<iiiiddqii", nmesh, nc, 1, 4, L, 1.0, npart, 1, 0)
    out = emul("tpaint", head + x.tobytes(), str(tmp_path))
    got = np.frombuffer(out, dtype=np.float32, count=nmesh ** 3).reshape(nmesh, nmesh, nmesh)
    xo = np.frombuffer(out, dtype=np.float64, count=3 * npart, offset=4 * nmesh ** 3).reshape(npart, 3)
    assert np.abs(got - want).max() <= 2e-6 * want.max()
    assert np.all((xo >= 0) & (xo <= L)) and np.abs(np.minimum(np.abs(xo - xw), L - np.abs(xo - xw))).max() < 1e-9
    xs = rng.permutation(xw)
    want = s.real_view(s.paint(xs)).copy()
    head = struct.pack("<iiiiddqii", nmesh, nc, 0, 4, L, 1.0, npart, 1, 0)
    out = emul("tpaint", head + xs.tobytes(), str(tmp_path))
    got = np.frombuffer(out, dtype=np.float32, count=nmesh ** 3).reshape(nmesh, nmesh, nmesh)
    stats = np.frombuffer(out[-16:], dtype=np.uint64)
    assert np.abs(got - want).max() <= 2e-6 * want.max()
    assert stats[0] > 0.5 * npart, stats
    field = rng.standard_normal((nmesh, nmesh, nmesh)).astype(np.float32)
    out = emul("treadout", head + xs.tobytes() + field.tobytes(), str(tmp_path))
    assert np.array_equal(np.frombuffer(out[:-16], dtype=np.float32), s.readout(s.real_pack(field), xs))
    s.close()


def test_tile_kernels_on_a_slab(emul, one_thread_ref, tmp_path):
    """x-slab geometry of rank 1 of 2 (local planes 0 .. nxl, the last one the halo plane): the particles of that slab, deposited
    through the tile kernel, give the reference's planes [x0, x0 + nxl] (the halo plane being the next rank's plane 0)."""
    nc, L, nmesh = 16, 64.0, 32
    rng = np.random.default_rng(93)
    x = _lagrangian_positions(rng, nc, L, 1.5 * L / nmesh, nstray=0)
    cell = np.floor(x[:, 0] / (L / nmesh)).astype(int) % nmesh
    nxl, x0 = nmesh // 2, nmesh // 2
    mine = x[(cell >= x0) & (cell < x0 + nxl)]
    mine = mine[: (len(mine) // (8 * nc * nc)) * 8 * nc * nc]            # complete groups of 8 i-planes (the launcher's rule)
    assert len(mine) >= 8 * nc * nc
    s = one_thread_ref.Session(nc=nmesh, boxsize=L, pm_nc_factor=1)
    full = s.real_view(s.paint(mine)).copy()
    want = np.concatenate([full[x0:], full[:1]])                          # planes x0 .. N-1, then the halo plane = global plane 0
    head = struct.pack("<iiiiddqii", nmesh, nc, 0, 4, L, 1.0, len(mine), 2, 1)
    out = emul("tpaint", head + mine.tobytes(), str(tmp_path))
    got = np.frombuffer(out, dtype=np.float32, count=(nxl + 1) * nmesh ** 2).reshape(nxl + 1, nmesh, nmesh)
    assert np.abs(got - want).max() <= 2e-6 * full.max()
    assert abs(got.sum(dtype=np.float64) - len(mine)) < 1e-3
    field = rng.standard_normal((nmesh, nmesh, nmesh)).astype(np.float32)
    local = np.concatenate([field[x0:], field[:1]])
    out = emul("treadout", head + mine.tobytes() + local.tobytes(), str(tmp_path))
    assert np.array_equal(np.frombuffer(out[:-16], dtype=np.float32), s.readout(s.real_pack(field), mine))
    s.close()


@pytest.mark.parametrize("lag_nc", [0, 16])
def test_three_component_readout_equals_three_readouts(emul, one_thread_ref, tmp_path, lag_nc):
    """cic_readout3_kernel (FASTPM_B200_FUSED_READOUT=1: positions read once, ACC written as whole elements) against three
    passes of cic_readout_kernel and against the reference's readout, bit for bit, in particle order and in the brick walk."""
    nmesh, L = 32, 77.0
    npart = 16 ** 3 + 333
    rng = np.random.default_rng(5 + lag_nc)
    x = _positions(rng, npart, L)
    fields = rng.standard_normal((3, nmesh, nmesh, nmesh)).astype(np.float32)
    out = emul("readout3", struct.pack("<iidq", nmesh, lag_nc, L, npart) + x.tobytes() + fields.tobytes(), str(tmp_path))
    both = np.frombuffer(out, dtype=np.float32).reshape(2, npart, 3)
    assert np.array_equal(both[0], both[1])
    s = one_thread_ref.Session(nc=nmesh, boxsize=L, pm_nc_factor=1)
    for d in range(3):
        assert np.array_equal(both[1][:, d], s.readout(s.real_pack(fields[d]), x))
    s.close()


WINDOW_IDS = {"cic": 0, "linear": 1, "quad": 2, "lanczos": 3}


def _window_case(emul, ref, tmp_path, window, support, diffdir, nslab, seed, exact_readout=True):
    nmesh, L, npart = 16, 50.0, 3000
    rng = np.random.default_rng(seed)
    x = _positions(rng, npart, L)
    s = ref.Session(nc=nmesh, boxsize=L, pm_nc_factor=1)
    want = s.real_view(s.paint_window(x, window, support, diffdir=diffdir)).copy()
    head = struct.pack("<iiiiiddq", nmesh, WINDOW_IDS[window], support, diffdir, nslab, L, 1.0, npart)
    got = np.frombuffer(emul("wpaint", head + x.tobytes(), str(tmp_path)), dtype=np.float32).reshape(nmesh, nmesh, nmesh)
    assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()
    if diffdir < 0:
        assert abs(got.sum(dtype=np.float64) - npart) < 1e-3       # the windows are normalised: mass is conserved
    else:
        assert np.abs(want).max() > 0.1                            # the derivative painter really did something
    field = rng.standard_normal((nmesh, nmesh, nmesh)).astype(np.float32)
    want_r = s.readout_window(s.real_pack(field), x, window, support, diffdir=diffdir)
    got_r = np.frombuffer(emul("wreadout", head + x.tobytes() + field.tobytes(), str(tmp_path)), dtype=np.float32)
    if exact_readout:
        assert np.array_equal(got_r, want_r)
    else:
        assert np.abs(got_r - want_r).max() <= 1e-6 * np.abs(want_r).max()
    s.close()


@pytest.mark.parametrize("window,support", [("linear", 2), ("quad", 3), ("lanczos", 4), ("lanczos", 6)])
def test_generic_window_kernels_against_reference(emul, one_thread_ref, tmp_path, window, support):
    """_generic_paint / _generic_readout (painter.c:176-317) with the linear, quadratic and Lanczos windows (the last with the
    reference's 1e-3 table quantisation): readout bit for bit, deposit up to the float rounding of each add (as for CIC)."""
    _window_case(emul, one_thread_ref, tmp_path, window, support, -1, 1, 13 + support)


@pytest.mark.parametrize("window,support,diffdir", [("cic", 2, 0), ("cic", 2, 2), ("linear", 2, 1), ("quad", 3, 0), ("lanczos", 4, 2), ("lanczos", 6, 1)])
def test_derivative_painters_against_reference(emul, one_thread_ref, tmp_path, window, support, diffdir):
    """fastpm_painter_init_diff (painter.c:178-205, painter-cic.c:57-60): the window along one axis replaced by its derivative /
    cellsize, the CIC window included; the Lanczos derivative with the reference's shared-table behaviour (csrc/window.h)."""
    _window_case(emul, one_thread_ref, tmp_path, window, support, diffdir, 1, 31 + support + diffdir)


@pytest.mark.parametrize("window,support,diffdir,nslab", [("quad", 3, -1, 2), ("lanczos", 4, -1, 4), ("lanczos", 6, -1, 2), ("linear", 2, -1, 4),
                                                          ("lanczos", 6, 0, 4), ("cic", 2, 1, 2)])
def test_window_kernels_on_slabs(emul, one_thread_ref, tmp_path, window, support, diffdir, nslab):
    """The same kernels in their several-GPU form: every slab deposits into its planes and its block of halo planes, the blocks
    are exchanged the way csrc/comm.cu does (fpm_halo_add_wide_from / fpm_halo_fetch_wide_from), and the result is the
    reference's full mesh / readout (the reference moves ghost particles instead, pmghosts.c:45-78)."""
    _window_case(emul, one_thread_ref, tmp_path, window, support, diffdir, nslab, 57 + support + nslab)


def test_fused_wrap_and_brick_walk(emul, one_thread_ref, tmp_path):
    """Positions outside the box: wrap folded into the deposit == the reference's wrap followed by its paint; the Lagrangian
    brick walk (any permutation of the particles) gives the same mesh up to the order of float additions."""
    nc, nmesh, L = 16, 32, 64.0
    npart = nc ** 3 + 4 * nc * nc + 77
    rng = np.random.default_rng(9)
    x = rng.uniform(-1.5 * L, 2.5 * L, size=(npart, 3))
    x[0] = [L, -L, 2 * L]
    xw = np.frombuffer(emul("wrap", struct.pack("<qd", 3 * npart, L) + x.tobytes(), str(tmp_path)), dtype=np.float64, count=3 * npart).reshape(npart, 3)
    dist = np.abs(xw - np.mod(x, L))                                 # store.c:447-475: remainder() + fold, result in [0, L]
    assert np.minimum(dist, L - dist).max() < 1e-9 and xw.min() >= 0 and xw.max() <= L
    s = one_thread_ref.Session(nc=nmesh, boxsize=L, pm_nc_factor=1)
    want = s.real_view(s.paint(xw)).copy()
    s.close()
    out = emul("paint", struct.pack("<iiiiddq", nmesh, 0, 1, 4, L, 1.0, npart) + x.tobytes(), str(tmp_path))
    got = np.frombuffer(out, dtype=np.float32, count=nmesh ** 3).reshape(nmesh, nmesh, nmesh)
    x_after = np.frombuffer(out, dtype=np.float64, count=3 * npart, offset=4 * nmesh ** 3).reshape(npart, 3)
    assert np.array_equal(x_after, xw)                               # same operations as the stand-alone wrap kernel
    assert np.array_equal(got != 0, want != 0) and np.abs(got - want).max() <= 4e-7 * want.max()
    out = emul("paint", struct.pack("<iiiiddq", nmesh, nc, 0, 4, L, 1.0, npart) + xw.tobytes(), str(tmp_path))
    brick = np.frombuffer(out, dtype=np.float32, count=nmesh ** 3).reshape(nmesh, nmesh, nmesh)
    assert abs(brick.sum(dtype=np.float64) - npart) < 1e-2
    assert np.abs(brick - want).max() <= 4e-6 * want.max()


@pytest.mark.parametrize("mode", ["fastpm", "pm", "cola"])
def test_kick_drift_and_fused_update_kernels_bit_exact(emul, one_thread_ref, pk_text, tmp_path, mode):
    nc, L = 16, 64.0
    s = one_thread_ref.Session(nc=nc, boxsize=L, pm_nc_factor=2, force_mode=mode, growth_mode="LCDM")
    dk, _, _ = s.ic_deltak(11, pk_text)
    s.setup_lpt(dk, 0.1)
    s.compute_force(0.1)
    p0 = s.get_particles()
    n = s.np
    ai, ac, af = 0.1, 0.1, 0.2
    kf, df = s.kick_factor(ai, ac, af), s.drift_factor(ai, ac, af)
    s.kick(ai, ac, af)
    s.drift(ai, ac, af)
    p1 = s.get_particles()
    s.close()
    cola = 1 if mode == "cola" else 0
    fm = {"fastpm": 0, "pm": 1, "cola": 2}[mode]
    k5 = [kf["dda"][-1] - kf["dda"][0], kf["q1"], kf["q2"], kf["Dv1"][-1] - kf["Dv1"][0], kf["Dv2"][-1] - kf["Dv2"][0]]
    d5 = [df["dyyy"][-1] - df["dyyy"][0], df["da1"][-1] - df["da1"][0], df["da2"][-1] - df["da2"][0], df["Dv1"], df["Dv2"]]
    zeros = np.zeros((n, 3), dtype=np.float32)
    payload = struct.pack("<qii", n, cola, fm) + np.array(k5).tobytes() + np.array(d5).tobytes() + p0["x"].tobytes() + p0["v"].tobytes() + \
        p0["acc"].tobytes() + (p0["dx1"] if cola else zeros).tobytes() + (p0["dx2"] if cola else zeros).tobytes()
    out = emul("update", payload, str(tmp_path))
    off = 0

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(out, dtype=dtype, count=count, offset=off)
        off += a.nbytes
        return a.reshape(n, 3)
    x_kd, v_kd = take(np.float64, 3 * n), take(np.float32, 3 * n)
    x_seq, v_seq = take(np.float64, 3 * n), take(np.float32, 3 * n)
    x_fused, v_fused = take(np.float64, 3 * n), take(np.float32, 3 * n)
    assert np.array_equal(v_kd, p1["v"]) and np.array_equal(x_kd, p1["x"])           # kick + drift == the reference's
    assert np.array_equal(v_fused, v_seq) and np.array_equal(x_fused, x_seq)         # one fused pass == five separate kernels


@pytest.mark.parametrize("mode", ["fastpm", "cola"])
def test_drift_with_pgd_column_bit_exact(emul, one_thread_ref, pk_text, tmp_path, mode):
    """fastpm_drift_one with a pgdc column (factors.c:108-113): drift_kernel + pgd_shift_kernel == the reference's positions."""
    nc, L = 16, 64.0
    s = one_thread_ref.Session(nc=nc, boxsize=L, pm_nc_factor=2, force_mode=mode, growth_mode="LCDM", pgdc=(0.8, 4.0, 2.0, 1.0, 10.0))
    dk, _, _ = s.ic_deltak(12, pk_text)
    s.setup_lpt(dk, 0.1)
    n = s.np
    rng = np.random.default_rng(3)
    pg = (rng.standard_normal((n, 3)) * 0.3).astype(np.float32)
    s.set_pgdc(pg)
    p0 = s.get_particles()
    ai, ac, af = 0.1, 0.15, 0.2
    df = s.drift_factor(ai, ac, af)
    s.drift(ai, ac, af)
    p1 = s.get_particles()
    s.close()
    fm = {"fastpm": 0, "pm": 1, "cola": 2}[mode]
    d5 = [df["dyyy"][-1] - df["dyyy"][0], df["da1"][-1] - df["da1"][0], df["da2"][-1] - df["da2"][0], df["Dv1"], df["Dv2"]]
    zeros = np.zeros((n, 3), dtype=np.float32)
    payload = struct.pack("<qi", n, fm) + np.array(d5).tobytes() + struct.pack("<d", df["dyyy"][-1]) + p0["x"].tobytes() + p0["v"].tobytes() + \
        (p0["dx1"] if fm == 2 else zeros).tobytes() + (p0["dx2"] if fm == 2 else zeros).tobytes() + pg.tobytes()
    x = np.frombuffer(emul("pgddrift", payload, str(tmp_path)), dtype=np.float64).reshape(n, 3)
    assert np.abs(x - p0["x"]).max() > 1e-3                 # something moved
    assert np.array_equal(x, p1["x"])


# ---------------------------------------------------------------- k-space kernels (tests/emul/kspace_emul.cpp)
@pytest.fixture(scope="module")
def kspace_emul(tmp_path_factory):
    if not os.path.exists(os.path.join(INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    exe = str(tmp_path_factory.mktemp("kemul") / "kspace_emul")
    env = dict(os.environ)
    env.pop("CXX", None)
    r = subprocess.run(["g++", "-O1", "-ffp-contract=off", "-std=c++17", "-pthread", "-w", "-I" + INC, "-o", exe,
                        os.path.join(ROOT, "tests", "emul", "kspace_emul.cpp")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    assert r.returncode == 0, r.stdout

    def run(op, payload, tmp):
        fin, fout = os.path.join(tmp, op + ".in"), os.path.join(tmp, op + ".out")
        with open(fin, "wb") as f:
            f.write(payload)
        subprocess.run([exe, op, fin, fout], check=True)
        return open(fout, "rb").read()
    return run


def _sinc(x):
    x = np.asarray(x, dtype=np.float64)
    small = np.abs(x) < 1e-5
    xs = np.where(small, 1.0, x)
    return np.where(small, 1.0 - x * x / 6.0 + x ** 4 / 120.0, np.sin(xs) / xs)


def _tables(n, L):
    """pm_create_k_factors (pmapi.c:235-275) and the decic table (transfer.c:88-96), every intermediate in float where the
    reference's is -- the same restatement as host_ktables in csrc/capi.cu."""
    cell = L / n
    ii = np.arange(n)
    ii = np.where(ii >= n // 2, ii - n, ii)
    k = (ii * 2 * np.pi / L).astype(np.float32)
    w = (k.astype(np.float64) * cell).astype(np.float32)
    ff1 = _sinc(0.5 * w.astype(np.float64)).astype(np.float32)
    ff2 = _sinc(w.astype(np.float64)).astype(np.float32)
    kk = (k * k).astype(np.float32)
    kf = (1 / cell * (1 / 6.0 * (8 * np.sin(w.astype(np.float64)) - np.sin(2 * w.astype(np.float64))))).astype(np.float32)
    kkf = (kk * (ff1 * ff1).astype(np.float32)).astype(np.float32)
    kkf2 = (kk.astype(np.float64) * (4 / 3.0 * ff1.astype(np.float64) * ff1.astype(np.float64) - 1 / 3.0 * ff2.astype(np.float64) * ff2.astype(np.float64))).astype(np.float32)
    dec = 1.0 / _sinc(0.5 * (k.astype(np.float64) * L / n)) ** 2
    return np.stack([k, kk, kf, kkf, kkf2]).astype(np.float32), dec


def _to_device_layout(c, n):
    pitch_c = ((n // 2 + 1 + 15) // 16) * 16
    full = np.zeros((n, n, pitch_c), dtype=np.complex64)
    full[:, :, :n // 2 + 1] = np.transpose(c, (1, 0, 2))          # [kx, ky, kz] -> [ky][kx][kz]
    return full


def _from_device_layout(buf, n):
    pitch_c = ((n // 2 + 1 + 15) // 16) * 16
    full = np.frombuffer(buf, dtype=np.complex64).reshape(n, n, pitch_c)
    return np.transpose(full[:, :, :n // 2 + 1], (1, 0, 2))


def test_kspace_kernels_against_reference(kspace_emul, ref_mod, tmp_path):
    nmesh, L = 32, 123.0
    rng = np.random.default_rng(5)
    field = rng.normal(size=(nmesh, nmesh, nmesh)).astype(np.float32)
    s = ref_mod.Session(nc=nmesh, boxsize=L, pm_nc_factor=1, kernel_type="1_4")
    dk = s.r2c(s.real_pack(field))
    c = s.complex_view(dk)
    tab, dec = _tables(nmesh, L)
    head = struct.pack("<id", nmesh, L) + tab.tobytes() + dec.tobytes() + _to_device_layout(c, nmesh).tobytes()
    # force components and potential of the default kernel (potorder 0, gradorder 1): gravity.c:174-242
    for attr, memb in [(0, 0), (0, 1), (0, 2), (1, 0)]:
        want = s.complex_view(s.kernel_transfer(dk, memb, attr=attr))
        spec = struct.pack("<iiiiiii", 0, 1, 1 if attr == 0 else 0, memb, 0, 1, 1)
        got = _from_device_layout(kspace_emul("transfer", head + spec, str(tmp_path)), nmesh)
        assert np.array_equal(got.view(np.float32), want.view(np.float32)), (attr, memb)
    # CIC deconvolution, transfer.c:78-113
    want = s.complex_view(s.decic(dk))
    got = _from_device_layout(kspace_emul("decic", head, str(tmp_path)), nmesh)
    assert np.array_equal(got.view(np.float32), want.view(np.float32))
    # P(k): powerspectrum.c:35-124, plain and with the deconvolution folded into the read
    nb = nmesh // 2
    for decic, src in ((0, dk), (1, s.decic(dk))):
        k0, p0, n0 = s.powerspectrum(src)
        sums = np.frombuffer(kspace_emul("pk", head + struct.pack("<i", decic), str(tmp_path)), dtype=np.float64)
        nm, sp, sk = sums[:nb], sums[nb:2 * nb], sums[2 * nb:3 * nb]
        assert np.array_equal(nm, n0)
        sel = nm > 0
        np.testing.assert_allclose(sk[sel] / nm[sel], k0[sel], rtol=1e-13)
        np.testing.assert_allclose(sp[sel] / nm[sel] * L ** 3, p0[sel], rtol=1e-12)
    s.close()


def test_row_streaming_powerspectrum_kernel_against_reference(kspace_emul, ref_mod, tmp_path):
    """powerspectrum.c:35-124 through powerspectrum_rows_kernel (csrc/kspace.cu; meshes with (N/2) % 64 == 0, i.e. every bench size):
    mode counts exact, P(k) to 1e-12 of the compiled reference, plain and with the deconvolution folded into the read, and equal
    to the shuffle-reduction kernel it replaces on those sizes."""
    nmesh, L = 128, 200.0
    rng = np.random.default_rng(11)
    field = rng.normal(size=(nmesh, nmesh, nmesh)).astype(np.float32)
    s = ref_mod.Session(nc=nmesh, boxsize=L, pm_nc_factor=1)
    dk = s.r2c(s.real_pack(field))
    c = s.complex_view(dk)
    tab, dec = _tables(nmesh, L)
    head = struct.pack("<id", nmesh, L) + tab.tobytes() + dec.tobytes() + _to_device_layout(c, nmesh).tobytes()
    nb = nmesh // 2
    # the four emulated runs at once (each spends most of its time in the barriers that stand in for warp shuffles)
    from concurrent.futures import ThreadPoolExecutor
    jobs = {}
    with ThreadPoolExecutor(max_workers=4) as ex:
        for decic in (0, 1):
            for op in ("pk_rows", "pk"):
                d = tmp_path / ("%s_%d" % (op, decic))
                d.mkdir()
                jobs[op, decic] = ex.submit(kspace_emul, op, head + struct.pack("<i", decic), str(d))
    for decic, src in ((0, dk), (1, s.decic(dk))):
        k0, p0, n0 = s.powerspectrum(src)
        sums = np.frombuffer(jobs["pk_rows", decic].result(), dtype=np.float64)
        old = np.frombuffer(jobs["pk", decic].result(), dtype=np.float64)
        nm, sp = sums[:nb], sums[nb:2 * nb]
        assert np.array_equal(nm, n0)
        sel = nm > 0
        np.testing.assert_allclose(sp[sel] / nm[sel] * L ** 3, p0[sel], rtol=1e-12)
        np.testing.assert_allclose(sums, old, rtol=1e-12)                  # includes the all-mode variance slot
    s.close()


def test_pgd_transfer_kernel_against_reference(kspace_emul, ref_mod, tmp_path):
    """pgdcorrection.c:28-137: the PGD potential sweep + gradient of the kernel sources, pushed through the reference's own
    c2r and readout, equals fastpm_pgdc_calculate bit for bit (same libm exp on the CPU)."""
    nmesh, L = 32, 50.0
    rng = np.random.default_rng(6)
    field = rng.normal(size=(nmesh, nmesh, nmesh)).astype(np.float32)
    x = rng.uniform(0, L, size=(3000, 3))
    par = (0.8, 4.0, 2.0, 1.3, 9.5)                          # alpha0, A, B, kl, ks
    a = 0.7
    alpha = par[0] * 10 ** (par[1] * a * a - par[2] * a)    # fastpm_pgdc_get_alpha, pgdcorrection.c:10-14
    s = ref_mod.Session(nc=nmesh, boxsize=L, pm_nc_factor=1)
    dk = s.r2c(s.real_pack(field))
    want = s.pgdc_calculate(dk, x, par, a=a)
    assert np.abs(want).max() > 0
    tab, dec = _tables(nmesh, L)
    head = struct.pack("<id", nmesh, L) + tab.tobytes() + dec.tobytes() + _to_device_layout(s.complex_view(dk), nmesh).tobytes()
    for d in range(3):
        got_k = _from_device_layout(kspace_emul("pgd", head + struct.pack("<dddi", alpha, par[3], par[4], d), str(tmp_path)), nmesh)
        got = s.readout(s.c2r(s.complex_pack(got_k)), x)
        assert np.array_equal(got, want[:, d]), d
    s.close()


@pytest.mark.parametrize("softening", ["gaussian", "gadget_long_range", "two_third", "gaussian36"])
def test_softening_kernels_against_reference(kspace_emul, one_thread_ref, pk_text, tmp_path, softening):
    """apply_softening_transfer (gravity.c:244-270) on delta_k inside fastpm_solver_compute_force: the reference's softened
    delta_k against our sweep of its unsoftened one, bit for bit.  The two Gaussian kinds are the per-axis factor kernel (the
    same one as the CIC deconvolution) with the table csrc/host/gravity.c builds; the other two the radial kernel."""
    nc, L, B = 8, 40.0, 2
    nmesh = nc * B
    kw = dict(nc=nc, boxsize=L, pm_nc_factor=B, force_mode="fastpm", growth_mode="LCDM", np_alloc_factor=2.0)
    out = {}
    for kind in ("none", softening):
        s = one_thread_ref.Session(softening=kind, **kw)                  # one thread: the same deposit order in both runs
        dk, _, _ = s.ic_deltak(3, pk_text)
        s.setup_lpt(dk, 0.5)
        out[kind] = s.complex_view(s.compute_force(0.5, want_delta_k=True), a=0.5).copy()
        s.close()
    assert not np.array_equal(out["none"], out[softening])
    tab, dec = _tables(nmesh, L)
    k_nq = np.pi / L * nmesh
    if softening in ("gaussian", "gadget_long_range"):
        N = 1.0 if softening == "gaussian" else 2 ** 0.5 * 1.25
        r0 = N * L / nmesh
        import ctypes
        libm = ctypes.CDLL("libm.so.6")                                   # the C library's exp, as host/gravity.c and the reference use
        libm.exp.restype, libm.exp.argtypes = ctypes.c_double, [ctypes.c_double]
        f = np.array([libm.exp(-0.5 * (float(k) * r0) ** 2) for k in tab[0]])          # exp(-0.5 * pow(k * r0, 2)), gravity.c:82
        head = struct.pack("<id", nmesh, L) + tab.tobytes() + f.tobytes() + _to_device_layout(out["none"], nmesh).tobytes()
        got = _from_device_layout(kspace_emul("decic", head, str(tmp_path)), nmesh)
    else:
        mode, param = (0, (2.0 / 3 * k_nq) ** 2) if softening == "two_third" else (1, k_nq)
        head = struct.pack("<id", nmesh, L) + tab.tobytes() + dec.tobytes() + _to_device_layout(out["none"], nmesh).tobytes()
        got = _from_device_layout(kspace_emul("radial", head + struct.pack("<id", mode, param), str(tmp_path)), nmesh)
    assert np.array_equal(got.view(np.float32), out[softening].view(np.float32))


def test_remove_variance_kernel_against_reference(kspace_emul, ref_mod, tmp_path):
    """fastpm_ic_remove_variance (initialcondition.c:66-99): unit amplitude, phase kept; zero modes stay zero."""
    n, L = 16, 100.0
    s = ref_mod.Session(nc=n, boxsize=L, pm_nc_factor=1)
    dk = s.fill_gaussian(31)
    c = s.complex_view(dk, which=1).copy()
    c[3, 5, 2] = 0
    dk = s.complex_pack(c, which=1)
    want = s.complex_view(s.remove_variance(dk), which=1)
    s.close()
    tab, dec = _tables(n, L)
    head = struct.pack("<id", n, L) + tab.tobytes() + dec.tobytes() + _to_device_layout(c, n).tobytes()
    got = _from_device_layout(kspace_emul("unitamp", head, str(tmp_path)), n)
    assert got[3, 5, 2] == 0
    assert np.array_equal(got.view(np.float32), want.view(np.float32))
    assert (c != 0).sum() > 2000
    np.testing.assert_allclose(np.abs(got[c != 0]), 1.0, rtol=2e-7)
    assert np.all(got[c == 0] == 0)


@pytest.mark.parametrize("remove_variance", [False, True])
def test_initial_condition_chain_against_reference(kspace_emul, ref_mod, pk_text, tmp_path, remove_variance):
    """prepare_deltak (src/fastpm.c:415-545) = Gadget white noise -> [remove_variance] -> induce_correlation with the P(k) table:
    the last two as kernel sources on the reference's white noise, against the reference's delta_k (the noise itself:
    tests/test_cpu_oracle_and_host.py).  fastpm_b200_setup_gadget_ic chains exactly these on the device."""
    n, L = 16, 200.0
    s = ref_mod.Session(nc=n, boxsize=L, pm_nc_factor=1)
    white = s.complex_view(s.fill_gaussian(77), which=1).copy()
    want = s.complex_view(s.ic_deltak(77, pk_text, remove_variance=remove_variance)[0], which=1).copy()
    s.close()
    tab, dec = _tables(n, L)
    base = struct.pack("<id", n, L) + tab.tobytes() + dec.tobytes()
    cur = white
    if remove_variance:
        cur = _from_device_layout(kspace_emul("unitamp", base + _to_device_layout(cur, n).tobytes(), str(tmp_path)), n)
    pk = np.loadtxt(os.path.join(ROOT, "tests", "golden", "powerspec.txt"))
    payload = struct.pack("<i", len(pk)) + np.ascontiguousarray(pk[:, 0]).tobytes() + np.ascontiguousarray(pk[:, 1]).tobytes()
    got = _from_device_layout(kspace_emul("induce", base + _to_device_layout(cur, n).tobytes() + payload, str(tmp_path)), n).copy()
    got[0, 0, 0] = 1.0                                   # fastpm_apply_modify_mode_transfer, src/fastpm.c:541-544
    assert np.array_equal(got.view(np.float32), want.view(np.float32))
