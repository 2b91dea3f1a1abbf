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
