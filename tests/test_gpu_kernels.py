"""GPU parity tests: every CUDA kernel against the oracle (the compiled reference) on the same seeded inputs.

Integer/byte-exact where the arithmetic is reproducible (readout, transfer, decic, kick, drift, wrap);
a stated float tolerance where float32 summation order differs by construction (paint, FFT).
All calls go through the C ABI (include/fastpm_b200.h) via fastpm_b200.device.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from fastpm_b200 import device
    device._lib.require_device()
    return device


def _positions(rng, n, L, kind):
    if kind == "uniform":
        return rng.uniform(0, L, size=(n, 3))
    if kind == "clustered":
        c = rng.uniform(0, L, size=(8, 3))
        x = c[rng.integers(0, 8, n)] + rng.normal(0, 0.01 * L, size=(n, 3))
        return np.mod(x, L)
    if kind == "edges":
        x = rng.uniform(0, L, size=(n, 3))
        x[: n // 4] = np.round(x[: n // 4] / (L / 8)) * (L / 8)      # exactly on cell faces, including x == L
        x[0] = [L, L, L]
        x[1] = [0, 0, 0]
        x[2] = [np.nextafter(L, 0)] * 3
        return x
    raise ValueError(kind)


@pytest.mark.parametrize("nmesh,kind", [(32, "uniform"), (48, "clustered"), (64, "edges")])
def test_paint_matches_reference(dev, ref_mod, nmesh, kind):
    L = 100.0
    rng = np.random.default_rng(nmesh)
    x = _positions(rng, 20000, L, kind)
    s = ref_mod.Session(nc=nmesh, boxsize=L, pm_nc_factor=1)
    want = s.real_view(s.paint(x))
    m = dev.Mesh(nmesh, L)
    canvas = m.alloc()
    xd = dev.DeviceBuffer.from_host(x)
    m.paint(canvas, xd, len(x), M0=1.0)
    got = m.download_real(canvas)
    # same double-precision weights; only the order of float32 additions into a cell differs
    assert abs(got.sum(dtype=np.float64) - len(x)) < 1e-3 * len(x) ** 0.5
    np.testing.assert_allclose(got, want, rtol=0, atol=4e-6 * max(1.0, want.max()))
    s.close()


@pytest.mark.parametrize("nmesh,kind", [(32, "uniform"), (64, "edges")])
def test_readout_bit_exact(dev, ref_mod, nmesh, kind):
    L = 64.0
    rng = np.random.default_rng(nmesh + 1)
    x = _positions(rng, 30000, L, kind)
    field = rng.normal(size=(nmesh, nmesh, nmesh)).astype(np.float32)
    s = ref_mod.Session(nc=nmesh, boxsize=L, pm_nc_factor=1)
    want = s.readout(s.real_pack(field), x)
    m = dev.Mesh(nmesh, L)
    canvas = m.alloc()
    m.upload_real(canvas, field)
    xd = dev.DeviceBuffer.from_host(x)
    out = dev.DeviceBuffer(4 * len(x))
    m.readout(canvas, xd, len(x), out)
    got = out.download(np.float32)
    assert np.array_equal(got, want)
    s.close()


@pytest.mark.parametrize("window,support,diffdir", [("cic", 2, 0), ("linear", 2, 2), ("quad", 3, 1), ("lanczos", 6, 0), ("lanczos", 4, -1)])
def test_window_and_derivative_painters_match_reference(dev, ref_mod, window, support, diffdir):
    """fpm_paint_window_ex / fpm_readout_window_ex: the generic windows (painter.c:217-317) and the derivative painters of
    fastpm_painter_init_diff (painter.c:178-205; CIC: painter-cic.c:57-60): readout bit for bit, deposit to the float rounding of
    the adds."""
    nmesh, L = 32, 50.0
    rng = np.random.default_rng(77 + support + diffdir)
    x = _positions(rng, 20000, L, "uniform")
    wid = {"cic": 0, "linear": 1, "quad": 2, "lanczos": 3}[window]
    s = ref_mod.Session(nc=nmesh, boxsize=L, pm_nc_factor=1)
    want = s.real_view(s.paint_window(x, window, support, diffdir=diffdir)).copy()
    m = dev.Mesh(nmesh, L)
    lib = m.lib
    canvas = m.alloc()
    dev.check(lib.fpm_memset(canvas.ptr, 0, canvas.nbytes))
    xd = dev.DeviceBuffer.from_host(x)
    dev.check(lib.fpm_paint_window_ex(m.h, wid, support, diffdir, canvas.ptr, None, xd.ptr, len(x), 1.0, None, None, 1))
    got = m.download_real(canvas)
    assert np.abs(got - want).max() <= 4e-6 * np.abs(want).max()
    field = rng.normal(size=(nmesh, nmesh, nmesh)).astype(np.float32)
    want_r = s.readout_window(s.real_pack(field), x, window, support, diffdir=diffdir)
    m.upload_real(canvas, field)
    out = dev.DeviceBuffer(4 * len(x))
    dev.check(lib.fpm_readout_window_ex(m.h, wid, support, diffdir, canvas.ptr, None, xd.ptr, len(x), out.ptr, 1))
    assert np.array_equal(out.download(np.float32), want_r)
    s.close()


def test_empty_and_single_particle(dev, ref_mod):
    """Edge cases of the store loops (painter.c:320-374, factors.c:176-197): np = 0 leaves everything untouched, np = 1 on a
    cell corner, on the box edge and outside the box deposits unit mass into the periodic images the reference uses."""
    nmesh, L = 16, 32.0
    m = dev.Mesh(nmesh, L)
    lib = m.lib
    canvas = m.alloc()
    xd = dev.DeviceBuffer(24)
    m.paint(canvas, xd, 0)
    out = dev.DeviceBuffer(4)
    out.upload(np.array([7.0], dtype=np.float32))
    m.readout(canvas, xd, 0, out)
    dev.check(lib.fpm_kick(out.ptr, out.ptr, out.ptr, None, None, 0, 0, 1.0, 0.0, 0.0, 0.0, 0.0))
    dev.check(lib.fpm_wrap(xd.ptr, 0, L))
    assert m.download_real(canvas).sum() == 0 and out.download(np.float32)[0] == 7.0
    s = ref_mod.Session(nc=nmesh, boxsize=L, pm_nc_factor=1)
    for pos in ([0.0, 0.0, 0.0], [L, L, L], [-0.25, L + 0.5, 3.0 * L + 1.0], [L - 1e-12, 1.0, 2.0]):
        x = np.array([pos], dtype=np.float64)
        canvas.zero()
        xd.upload(x)
        m.paint(canvas, xd, 1, M0=1.0)
        got = m.download_real(canvas)
        want = s.real_view(s.paint(x))
        assert np.array_equal(got, want), pos             # a single particle: no summation-order freedom, bit-exact
        assert abs(got.sum(dtype=np.float64) - 1.0) < 1e-6
    s.close()


def test_brick_traversal_is_a_permutation(dev):
    """The Lagrangian-brick walk of paint / readout (large meshes) visits every particle exactly once: same mesh (up to the
    order of float additions) and bit-identical readout as the linear walk, including a tail that is not a whole group."""
    nc, nmesh, L = 16, 32, 64.0
    rng = np.random.default_rng(21)
    n = nc ** 3 + 4 * nc * nc + 77                       # 5 complete groups of 4 planes + a ragged tail
    x = rng.uniform(-0.2 * L, 1.2 * L, size=(n, 3))
    m = dev.Mesh(nmesh, L)
    lib = m.lib
    xd = dev.DeviceBuffer.from_host(x)
    res = {}
    for tag, hint in (("linear", 0), ("bricks", -nc)):
        dev.check(lib.fpm_particle_grid_hint(hint))
        canvas = m.alloc()
        m.paint(canvas, xd, n, M0=1.5)
        out = dev.DeviceBuffer(4 * n)
        out.zero()
        m.readout(canvas, xd, n, out)
        res[tag] = (m.download_real(canvas), out.download(np.float32))
    dev.check(lib.fpm_particle_grid_hint(0))
    a, b = res["linear"], res["bricks"]
    assert abs(a[0].sum(dtype=np.float64) - 1.5 * n) < 1e-3 * n * 1e-3 + 1e-2
    assert np.abs(a[0] - b[0]).max() <= 4e-6 * np.abs(a[0]).max()
    # each readout gathers from its own canvas: compare each against a gather of the SAME canvas walked the other way
    dev.check(lib.fpm_particle_grid_hint(-nc))
    canvas = m.alloc()
    m.upload_real(canvas, a[0])
    out = dev.DeviceBuffer(4 * n)
    m.readout(canvas, xd, n, out)
    dev.check(lib.fpm_particle_grid_hint(0))
    out2 = dev.DeviceBuffer(4 * n)
    m.readout(canvas, xd, n, out2)
    assert np.array_equal(out.download(np.float32), out2.download(np.float32))


@pytest.mark.parametrize("nmesh", [16, 24, 32, 48, 64, 96, 128, 160])
def test_r2c_c2r_match_reference(dev, ref_mod, nmesh):
    L = 200.0
    rng = np.random.default_rng(nmesh)
    field = rng.normal(size=(nmesh, nmesh, nmesh)).astype(np.float32)
    s = ref_mod.Session(nc=nmesh, boxsize=L, pm_nc_factor=1)
    ck = s.r2c(s.real_pack(field))
    want_k = s.complex_view(ck)
    m = dev.Mesh(nmesh, L)
    real, cplx = m.alloc(), m.alloc()
    m.upload_real(real, field)
    m.r2c(real, cplx)
    got_k = m.download_complex(cplx)
    scale = np.abs(want_k).max()
    err = np.abs(got_k - want_k).max() / scale
    assert err < 2e-6, err                      # two float32 FFTs of the same field
    # back: the round trip is the identity because r2c carried 1/N^3 (pmpfft.c:373-391)
    back = m.alloc()
    m.c2r(cplx, back)
    got_r = m.download_real(back)
    want_r = s.real_view(s.c2r(ck))
    assert np.abs(got_r - want_r).max() < 5e-6 * np.abs(want_r).max()
    assert np.abs(got_r - field).max() < 5e-6 * np.abs(field).max()
    s.close()


@pytest.mark.parametrize("kernel_type", ["1_4", "3_4", "5_4", "3_2", "naive"])
def test_gravity_kernel_bit_exact(dev, ref_mod, kernel_type):
    nmesh, L = 32, 123.0
    rng = np.random.default_rng(5)
    field = rng.normal(size=(nmesh, nmesh, nmesh)).astype(np.float32)
    s = ref_mod.Session(nc=nmesh, boxsize=L, pm_nc_factor=1, kernel_type=kernel_type)
    dk = s.r2c(s.real_pack(field))
    m = dev.Mesh(nmesh, L)
    src, dst = m.alloc(), m.alloc()
    m.upload_complex(src, s.complex_view(dk))
    kt = m.ktables()
    for attr, memb in [(0, 0), (0, 1), (0, 2), (1, 0)]:
        want = s.complex_view(s.kernel_transfer(dk, memb, attr=attr))
        m.apply_transfer(src, dst, m.transfer_for_kernel(kernel_type, attr, memb))
        got = m.download_complex(dst)
        assert np.array_equal(got.view(np.float32), want.view(np.float32)), (kernel_type, attr, memb)
    # the tidal components (gravity.c:194-231): potential, gradient along d1, gradient along d2 -- the same spec with two gradients
    # (what host/gravity.c: gravity_apply_kernel_transfer builds for COLUMN_TIDAL)
    D1, D2 = (0, 1, 2, 0, 1, 2), (0, 1, 2, 1, 2, 0)
    for memb in range(6):
        want = s.complex_view(s.kernel_transfer(dk, memb, attr=3))
        t = m.transfer_for_kernel(kernel_type, 0, D1[memb])
        t.ngrad = 2
        t.graddir[1] = D2[memb]
        m.apply_transfer(src, dst, t)
        got = m.download_complex(dst)
        assert np.array_equal(got.view(np.float32), want.view(np.float32)), (kernel_type, "tidal", memb)
    assert np.all(np.isfinite(kt["k_finite"]))
    s.close()


def test_ic_kernel_and_decic_bit_exact(dev, ref_mod):
    nmesh, L = 32, 77.0
    rng = np.random.default_rng(6)
    field = rng.normal(size=(nmesh, nmesh, nmesh)).astype(np.float32)
    s = ref_mod.Session(nc=nmesh, boxsize=L, pm_nc_factor=1)
    dk = s.r2c(s.real_pack(field))
    m = dev.Mesh(nmesh, L)
    src, dst = m.alloc(), m.alloc()
    m.upload_complex(src, s.complex_view(dk))
    for d1, d2 in [(0, -1), (2, -1), (1, 1), (0, 2)]:
        want = s.complex_view(s.laplace_diff(dk, d1, d2, which=0))
        dirs = tuple(d for d in (d1, d2) if d >= 0)
        m.apply_transfer(src, dst, dev.ic_transfer(0, dirs, 1))
        got = m.download_complex(dst)
        assert np.array_equal(got.view(np.float32), want.view(np.float32)), (d1, d2)
    want = s.complex_view(s.decic(dk))
    m.decic(src, dst)
    assert np.array_equal(m.download_complex(dst).view(np.float32), want.view(np.float32))
    s.close()


def test_fused_transfer_c2r_matches_separate(dev, ref_mod):
    nmesh, L = 48, 150.0
    rng = np.random.default_rng(7)
    field = rng.normal(size=(nmesh, nmesh, nmesh)).astype(np.float32)
    s = ref_mod.Session(nc=nmesh, boxsize=L, pm_nc_factor=1)
    dk = s.r2c(s.real_pack(field))
    m = dev.Mesh(nmesh, L)
    src, out = m.alloc(), m.alloc()
    m.upload_complex(src, s.complex_view(dk))
    for memb in range(3):
        want = s.real_view(s.c2r(s.kernel_transfer(dk, memb)))
        m.c2r(src, out, m.transfer_for_kernel("1_4", 0, memb))
        got = m.download_real(out)
        assert np.abs(got - want).max() < 5e-6 * np.abs(want).max()
    s.close()


def test_powerspectrum_matches_reference(dev, ref_mod):
    nmesh, L = 48, 300.0
    rng = np.random.default_rng(8)
    field = rng.normal(size=(nmesh, nmesh, nmesh)).astype(np.float32)
    s = ref_mod.Session(nc=nmesh, boxsize=L, pm_nc_factor=1)
    dk = s.r2c(s.real_pack(field))
    k0, p0, n0 = s.powerspectrum(dk)
    m = dev.Mesh(nmesh, L)
    src = m.alloc()
    m.upload_complex(src, s.complex_view(dk))
    k1, p1, n1 = m.powerspectrum(src)
    assert np.array_equal(n0, n1)                       # integer shell counts
    np.testing.assert_allclose(k1, k0, rtol=1e-13)
    np.testing.assert_allclose(p1, p0, rtol=1e-12)       # same doubles, different summation order
    k2, p2, n2 = m.powerspectrum(src, decic=True)
    k3, p3, n3 = s.powerspectrum(s.decic(dk))
    np.testing.assert_allclose(p2, p3, rtol=1e-12)
    # deferred deconvolution (solver.c:471 before the FORCE/after event): P(k) folds it into its read ...
    dev.check(m.lib.fpm_decic_defer(m.h, src.ptr))
    k4, p4, n4 = m.powerspectrum(src)
    np.testing.assert_allclose(p4, p2, rtol=1e-12)       # same values, double atomics in a different order
    assert np.array_equal(n4, n2)
    # ... and any other access through the library sees the deconvolved field (bit-exact with the reference's sweep)
    got = m.download_complex(src)
    assert np.array_equal(got, s.complex_view(s.decic(dk)))
    k5, p5, n5 = m.powerspectrum(src)                    # now materialised: a plain measurement gives the same numbers
    np.testing.assert_allclose(p5, p2, rtol=1e-12)
    # cross spectrum of two different fields (delta1_k != delta2_k, powerspectrum.c:87-91); a pending deconvolution is applied first
    field2 = (0.6 * field + 0.8 * rng.normal(size=field.shape)).astype(np.float32)
    dkb = s.r2c(s.real_pack(field2))
    kc, pc, nc_ = s.cross_powerspectrum(dk, dkb)
    src1, src2 = m.alloc(), m.alloc()
    m.upload_complex(src1, s.complex_view(dk))
    m.upload_complex(src2, s.complex_view(dkb))
    k6, p6, n6 = m.cross_powerspectrum(src1, src2)
    assert np.array_equal(n6, nc_)
    np.testing.assert_allclose(k6, kc, rtol=1e-13)
    np.testing.assert_allclose(p6, pc, rtol=1e-11, atol=1e-12 * np.abs(pc).max())
    assert np.abs(pc - p0).max() > 0.05 * np.abs(p0).max()        # really a different quantity from the auto spectrum
    dev.check(m.lib.fpm_decic_defer(m.h, src1.ptr))
    k7, p7, n7 = m.cross_powerspectrum(src1, src2)
    kd, pd, nd = s.cross_powerspectrum(s.decic(dk), dkb)
    np.testing.assert_allclose(p7, pd, rtol=1e-11, atol=1e-12 * np.abs(pd).max())
    s.close()


@pytest.mark.parametrize("mode", ["fastpm", "pm", "cola"])
def test_kick_drift_bit_exact(dev, ref_mod, pk_text, mode):
    import ctypes as C
    nc, L = 16, 64.0
    s = ref_mod.Session(nc=nc, boxsize=L, pm_nc_factor=2, force_mode=mode, growth_mode="LCDM")
    dk, _, _ = s.ic_deltak(11, pk_text)
    s.setup_lpt(dk, 0.1)
    s.compute_force(0.1)
    p0 = s.get_particles()
    n = s.np
    lib = dev._lib.require_device()
    ai, ac, af = 0.1, 0.1, 0.2
    kf = s.kick_factor(ai, ac, af)
    df = s.drift_factor(ai, ac, af)
    s.kick(ai, ac, af)
    s.drift(ai, ac, af)
    p1 = s.get_particles()
    B = dev.DeviceBuffer.from_host
    x, v, acc = B(p0["x"]), B(p0["v"]), B(p0["acc"])
    dx1 = B(p0["dx1"]) if mode == "cola" else None
    dx2 = B(p0["dx2"]) if mode == "cola" else None
    fm = dev.FORCE_MODES[mode]
    # factors at af minus factors at the particle time stamp (= ai): factors.c:148-160
    dev.check(lib.fpm_kick(v.ptr, v.ptr, acc.ptr, dx1.ptr if dx1 else None, dx2.ptr if dx2 else None, n, fm,
                           kf["dda"][-1] - kf["dda"][0], kf["q1"], kf["q2"], kf["Dv1"][-1] - kf["Dv1"][0], kf["Dv2"][-1] - kf["Dv2"][0]))
    # the reference drifted with the kicked velocity: do the same
    dev.check(lib.fpm_drift(x.ptr, x.ptr, v.ptr, dx1.ptr if dx1 else None, dx2.ptr if dx2 else None, n, fm,
                            df["dyyy"][-1] - df["dyyy"][0], df["da1"][-1] - df["da1"][0], df["da2"][-1] - df["da2"][0], df["Dv1"], df["Dv2"]))
    assert np.array_equal(v.download(np.float32).reshape(n, 3), p1["v"])
    assert np.array_equal(x.download(np.float64).reshape(n, 3), p1["x"])
    # the same kick and drift, then a second kick and two half drifts, as ONE fused pass: bit-identical to the sequence
    x2, v2 = B(p0["x"]), B(p0["v"])
    cola = 1 if mode == "cola" else 0
    k_op = [0, cola, kf["dda"][-1] - kf["dda"][0], kf["q1"], kf["q2"], kf["Dv1"][-1] - kf["Dv1"][0], kf["Dv2"][-1] - kf["Dv2"][0]]
    d_op = [1, fm, df["dyyy"][-1] - df["dyyy"][0], df["da1"][-1] - df["da1"][0], df["da2"][-1] - df["da2"][0], df["Dv1"], df["Dv2"]]
    d_half = [1, fm, 0.5 * d_op[2], 0.5 * d_op[3], 0.5 * d_op[4], df["Dv1"], df["Dv2"]]
    ops = np.array([k_op, d_op, k_op, d_half, d_half], dtype=np.float64)
    dev.check(lib.fpm_update_fused(x2.ptr, v2.ptr, acc.ptr, dx1.ptr if dx1 else None, dx2.ptr if dx2 else None, n, len(ops), ops.ctypes.data))
    dev.check(lib.fpm_kick(v.ptr, v.ptr, acc.ptr, dx1.ptr if dx1 else None, dx2.ptr if dx2 else None, n, fm, *k_op[2:]))
    for _ in range(2):
        dev.check(lib.fpm_drift(x.ptr, x.ptr, v.ptr, dx1.ptr if dx1 else None, dx2.ptr if dx2 else None, n, fm, *d_half[2:]))
    assert np.array_equal(v2.download(np.float32), v.download(np.float32))
    assert np.array_equal(x2.download(np.float64), x.download(np.float64))
    s.close()


def test_wrap_and_summary(dev):
    rng = np.random.default_rng(3)
    L = 50.0
    x = rng.uniform(-3 * L, 4 * L, size=(5000, 3))
    x[0] = [L, -L, 0.0]
    want = np.remainder(x, L)                  # same as remainder() + fold for these inputs except exact multiples
    xd = dev.DeviceBuffer.from_host(x)
    lib = dev._lib.require_device()
    dev.check(lib.fpm_wrap(xd.ptr, len(x), L))
    got = xd.download(np.float64).reshape(-1, 3)
    assert got.min() >= 0 and got.max() <= L
    d = np.abs(got - want)
    assert np.all(np.minimum(d, L - d) < 1e-9)
    st = dev.summary(xd, np.float64, 3, len(x))
    np.testing.assert_allclose(st["min"], got.min(axis=0))
    np.testing.assert_allclose(st["max"], got.max(axis=0))
    np.testing.assert_allclose(st["mean"], got.mean(axis=0), rtol=1e-12)
    np.testing.assert_allclose(st["std"], got.std(axis=0), rtol=1e-10)
    xbad = dev.DeviceBuffer.from_host(np.array([[1e9, 0.0, 0.0]]))
    assert lib.fpm_wrap(xbad.ptr, 1, L) == 0
    assert lib.fpm_wrap_check() != 0


@pytest.mark.parametrize("nmesh", [512, 768, 1024])
def test_tma_fft_matches_generic_and_oracle(dev, nmesh):
    """The TMA/register FFT passes (fft_tma.cu) against the generic shared-memory passes (fft.cu) and the oracle's CPU FFT."""
    import ctypes as C
    from oracle import port
    L = 1000.0
    rng = np.random.default_rng(nmesh)
    field = rng.standard_normal((nmesh, nmesh, nmesh), dtype=np.float32)
    m = dev.Mesh(nmesh, L)
    lib = m.lib
    real, ck_tma, ck_gen = m.alloc(), m.alloc(), m.alloc()
    m.upload_real(real, field)
    lib.fpm_fft_set_generic(0)
    m.r2c(real, ck_tma)
    a = m.download_complex(ck_tma)
    m.upload_real(real, field)
    lib.fpm_fft_set_generic(1)
    m.r2c(real, ck_gen)
    b = m.download_complex(ck_gen)
    lib.fpm_fft_set_generic(0)
    scale = np.abs(b).max()
    assert np.abs(a - b).max() < 2e-6 * scale
    want = port.fft3_r2c(field) / float(nmesh) ** 3              # oracle CPU FFT (oracle/shims/src/cpufft.c)
    assert np.abs(a - want).max() < 3e-6 * scale
    del b, want
    # inverse with the fused gravity kernel
    out_tma, out_gen = real, ck_gen
    kern = m.transfer_for_kernel("1_4", 0, 1)
    m.c2r(ck_tma, out_tma, kern)
    r1 = m.download_real(out_tma)
    lib.fpm_fft_set_generic(1)
    ck2 = m.alloc()
    m.upload_complex(ck2, a)
    m.c2r(ck2, out_gen, kern)
    lib.fpm_fft_set_generic(0)
    r2 = m.download_real(out_gen)
    assert np.abs(r1 - r2).max() < 5e-6 * np.abs(r2).max()
    # plain round trip
    m.upload_real(real, field)
    m.r2c(real, ck_tma)
    m.c2r(ck_tma, real)
    back = m.download_real(real)
    assert np.abs(back - field).max() < 1e-5 * np.abs(field).max()


@pytest.mark.parametrize("n", [2048, 1536])
def test_fft_2048_round_trip_and_generic_planes(dev, n):
    """BASELINE.json's full mesh size (2048^3, 34 GB per buffer) and the 1536^3 mesh of configs[3] (radix-24 first stage):
    size-independent properties instead of the oracle -- r2c -> c2r is the identity, Parseval's sum, and the TMA/register passes
    agree with the generic ones on sampled planes, without and with the fused gravity kernel."""
    L = 1024.0
    m = dev.Mesh(n, L)
    lib = m.lib
    import ctypes as C
    free_b, total_b = C.c_size_t(), C.c_size_t()
    lib.fpm_device_mem_info(C.byref(free_b), C.byref(total_b))
    if free_b.value < 3.2 * m.alloc_floats * 4:
        pytest.skip("needs 3 mesh buffers of %.1f GB" % (m.alloc_floats * 4 / 1e9))
    real, work, ck = dev.DeviceBuffer(m.alloc_floats * 4), dev.DeviceBuffer(m.alloc_floats * 4), dev.DeviceBuffer(m.alloc_floats * 4)
    plane = n * m.pitch_r
    from fastpm_b200 import _lib
    _lib.check(lib.fpm_fill_whitenoise(m.h, real.ptr, 1234), "whitenoise")
    sample = [0, 1, 777, n - 1]
    orig = {p: real.download(np.float32, plane, p * plane * 4).reshape(n, m.pitch_r)[:, :n].copy() for p in sample}
    # forward keeping the input (work buffer), TMA path
    lib.fpm_fft_set_generic(0)
    _lib.check(lib.fpm_r2c_ws(m.h, real.ptr, work.ptr, ck.ptr, 1.0 / float(n) ** 3), "r2c")
    kplanes = {p: ck.download(np.complex64, n * m.pitch_c, p * n * m.pitch_c * 8).reshape(n, m.pitch_c)[:, :n // 2 + 1].copy() for p in sample}
    # Parseval on the device: sum w |delta_k|^2 (last slot of the P(k) sums) == mean of the squared field
    sums = np.zeros(3 * (n // 2) + 1)
    _lib.check(lib.fpm_powerspectrum_sums(m.h, ck.ptr, 0, sums.ctypes.data), "pk sums")
    var_k = sums[-1]
    assert abs(var_k - 1.0) < 2e-3, var_k                # unit-variance white noise, 8.6e9 samples
    # generic path on the same input: sampled k-space planes agree
    lib.fpm_fft_set_generic(1)
    _lib.check(lib.fpm_r2c_ws(m.h, real.ptr, work.ptr, ck.ptr, 1.0 / float(n) ** 3), "r2c generic")
    lib.fpm_fft_set_generic(0)
    for p in sample:
        g = ck.download(np.complex64, n * m.pitch_c, p * n * m.pitch_c * 8).reshape(n, m.pitch_c)[:, :n // 2 + 1]
        scale = np.abs(g).max()
        assert np.abs(g - kplanes[p]).max() < 3e-6 * scale, p
    # inverse (TMA path) of the generic result gives the field back
    _lib.check(lib.fpm_c2r_ws(m.h, ck.ptr, work.ptr, work.ptr, None), "c2r")
    for p in sample:
        back = work.download(np.float32, plane, p * plane * 4).reshape(n, m.pitch_r)[:, :n]
        assert np.abs(back - orig[p]).max() < 2e-5 * np.abs(orig[p]).max(), p
    # the inverse with the fused gravity kernel (Green's function x i k_d, every direction): register passes against generic ones
    _lib.check(lib.fpm_r2c_ws(m.h, real.ptr, work.ptr, ck.ptr, 1.0 / float(n) ** 3), "r2c")
    for d in range(3):
        kern = m.transfer_for_kernel("1_4", 0, d)
        _lib.check(lib.fpm_c2r_ws(m.h, ck.ptr, work.ptr, work.ptr, C.byref(kern)), "c2r kernel")
        fast = {p: work.download(np.float32, plane, p * plane * 4).reshape(n, m.pitch_r)[:, :n].copy() for p in sample}
        lib.fpm_fft_set_generic(1)
        _lib.check(lib.fpm_c2r_ws(m.h, ck.ptr, work.ptr, work.ptr, C.byref(kern)), "c2r kernel generic")
        lib.fpm_fft_set_generic(0)
        for p in sample:
            g = work.download(np.float32, plane, p * plane * 4).reshape(n, m.pitch_r)[:, :n]
            assert np.abs(g - fast[p]).max() < 1e-5 * np.abs(g).max(), (d, p)
    for b in (real, work, ck):
        b.free()
    m.close()
