"""smoke(): one small PM run (2LPT + 3 KDK cycles, nc = 16, 32^3 mesh) on cuda:0 through the libfastpm API
mirror, checked against the committed reference fixture and, when the compiled reference is present,
against the oracle itself."""
import os
import numpy as np


def run_smoke():
    from fastpm_b200.solver import Solver
    here = os.path.dirname(os.path.abspath(__file__))
    fx = np.load(os.path.join(here, "golden", "small_run.npz"))
    L = 32.0
    g = Solver(nc=16, boxsize=L, pm_nc_factor=2, force_mode="fastpm", growth_mode="LCDM", np_alloc_factor=2.0)
    g.setup_lpt(fx["delta_k"], 0.1)
    x0 = g.get_column("x")
    d0 = np.abs(x0 - fx["x0"])
    assert np.minimum(d0, L - d0).max() < 1e-5, "2LPT positions differ from the reference fixture"
    g.evolve(fx["steps"])
    x1, v1 = g.get_column("x"), g.get_column("v")
    d1 = np.abs(np.mod(x1, L) - np.mod(fx["x1"], L))
    err = np.minimum(d1, L - d1).max()
    assert err < 1e-4, "positions differ from the reference fixture by %g Mpc/h" % err
    assert np.abs(v1 - fx["v1"]).max() < 1e-4 * np.abs(fx["v1"]).max()
    launches = int(g.lib.fpm_kernel_launch_count())
    g.close()
    msg = "smoke ok: max position error vs reference fixture %.3g Mpc/h, %d kernel launches" % (err, launches)
    from oracle import ref
    if ref.available():
        s = ref.Session(nc=16, boxsize=L, pm_nc_factor=2, force_mode="fastpm", growth_mode="LCDM", np_alloc_factor=2.0)
        s.setup_lpt(fx["delta_k"], 0.1)
        s.evolve(fx["steps"])
        p = s.get_particles()
        d2 = np.abs(np.mod(x1, L) - np.mod(p["x"], L))
        err2 = np.minimum(d2, L - d2).max()
        s.close()
        assert err2 < 1e-4, "positions differ from the oracle by %g Mpc/h" % err2
        msg += "; vs oracle %.3g Mpc/h" % err2
    print(msg)
    print(large_mesh_smoke())


def large_mesh_smoke():
    """Two force evaluations and one kick-drift-kick cycle on a 512^3 mesh (nc = 64, B = 8): the size from which the TMA tile pass
    (fft_tma_kernel), the bulk-copy row pass (fft_zrow_kernel) and the fused Green's function serve the transforms -- the kernels
    bench.py times -- checked against the oracle on the same delta_k when the compiled reference is present."""
    import ctypes as C
    from fastpm_b200.solver import Solver
    from oracle import ref
    nc, L, B = 64, 64.0, 8
    steps = np.array([0.1, 0.15])
    here = os.path.dirname(os.path.abspath(__file__))
    want = None
    if ref.available():
        s = ref.Session(nc=nc, boxsize=L, pm_nc_factor=B, force_mode="cola", growth_mode="LCDM", np_alloc_factor=2.0)
        dk, _, _ = s.ic_deltak(100, open(os.path.join(here, "golden", "powerspec.txt")).read())
        s.setup_lpt(dk, steps[0])
        s.evolve(steps)
        want = s.get_particles()
        s.close()
    g = Solver(nc=nc, boxsize=L, pm_nc_factor=B, force_mode="cola", growth_mode="LCDM", np_alloc_factor=1.0)
    if want is not None:
        g.setup_lpt(dk, steps[0])
    else:
        tab = np.loadtxt(os.path.join(here, "golden", "powerspec.txt"))
        g.setup_ic(100, tab[:, 0], tab[:, 1], steps[0])
    cnt = (C.c_uint64 * 4)()
    g.lib.fpm_path_counts.argtypes = [C.c_void_p, C.c_int]
    g.lib.fpm_path_counts(cnt, 4)
    before = [int(v) for v in cnt]
    g.evolve(steps)
    g.lib.fpm_path_counts(cnt, 4)
    used = [int(v) - b for v, b in zip(cnt, before)]
    x = g.get_column("x")
    g.close()
    assert used[0] >= 16 and used[2] >= 8 and used[1] == 0 and used[3] == 0, "512^3 mesh did not take the TMA / bulk-copy FFT passes: %s" % used
    assert np.isfinite(x).all()
    msg = "smoke (512^3 mesh) ok: %d TMA tile passes, %d bulk-copy row passes" % (used[0], used[2])
    if want is not None:
        d = np.abs(np.mod(x, L) - np.mod(want["x"], L))
        err = np.minimum(d, L - d).max()
        assert err < 1e-4, "512^3 mesh: positions differ from the oracle by %g Mpc/h" % err
        msg += "; max position error vs oracle %.3g Mpc/h" % err
    return msg
