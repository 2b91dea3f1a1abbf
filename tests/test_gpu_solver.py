"""GPU parity tests of the whole path through the libfastpm API mirror: 2LPT set-up and
fastpm_solver_evolve against the oracle (compiled reference) on identical initial conditions.

Tolerances are BASELINE.json's: positions <= 1e-4 Mpc/h (periodic distance, matched by particle id),
P(k) <= 1e-5 relative per bin (compared in memory as doubles).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _periodic_maxdiff(a, b, L):
    d = np.abs(a - b)
    return np.minimum(d, L - d).max()


def _run_pair(ref_mod, pk_text, nc, L, B, mode, growth, steps, seed=100):
    from fastpm_b200.solver import Solver, ForceEvent
    import ctypes as C
    s = ref_mod.Session(nc=nc, boxsize=L, pm_nc_factor=B, force_mode=mode, growth_mode=growth, np_alloc_factor=2.0)
    dk, _, _ = s.ic_deltak(seed, pk_text)
    s.setup_lpt(dk, steps[0])
    p0 = s.get_particles()

    g = Solver(nc=nc, boxsize=L, pm_nc_factor=B, force_mode=mode, growth_mode=growth, np_alloc_factor=2.0)
    g.setup_lpt(dk, steps[0])
    x0, v0 = g.get_column("x"), g.get_column("v")
    ic = dict(x_err=_periodic_maxdiff(x0, p0["x"], L), v_err=np.abs(v0 - p0["v"]).max(), v_scale=np.abs(p0["v"]).max())
    assert np.array_equal(g.get_column("id"), p0["id"])

    spectra = []

    def on_force_after(solver_ptr, event_ptr, userdata):
        ev = C.cast(event_ptr, C.POINTER(ForceEvent)).contents
        spectra.append((ev.a_f,) + g.powerspectrum_of(ev.pm, ev.delta_k))
        return 0

    g.add_handler("FORCE", 1, on_force_after)
    s.evolve(steps)
    g.evolve(steps)
    p1 = s.get_particles()
    out = dict(ic=ic, ref=p1, records=s.records(), spectra=spectra,
               x=g.get_column("x"), v=g.get_column("v"), acc=g.get_column("acc"), meta=g.meta)
    s.close()
    g.close()
    return out


@pytest.mark.parametrize("mode,growth", [("fastpm", "LCDM"), ("pm", "ODE"), ("cola", "LCDM")])
def test_evolve_matches_reference(ref_mod, pk_text, mode, growth):
    nc, L, B = 32, 64.0, 2
    steps = np.linspace(0.1, 1.0, 5)
    r = _run_pair(ref_mod, pk_text, nc, L, B, mode, growth, steps)
    # 2LPT initial conditions: same delta_k, float32 FFT + readout on both sides
    assert r["ic"]["x_err"] < 1e-5, r["ic"]
    assert r["ic"]["v_err"] < 1e-5 * max(1.0, r["ic"]["v_scale"]), r["ic"]
    # end state
    xerr = _periodic_maxdiff(np.mod(r["x"], L), np.mod(r["ref"]["x"], L), L)
    assert xerr < 1e-4, xerr                                  # Mpc/h, BASELINE.json north_star
    assert np.abs(r["v"] - r["ref"]["v"]).max() < 1e-4 * np.abs(r["ref"]["v"]).max()
    assert r["meta"]["a_x"] == 1.0 and r["meta"]["a_v"] == 1.0
    # P(k) at every force evaluation
    assert len(r["spectra"]) == len(r["records"]) == len(steps)
    for (a_f, k, p, nm), rec in zip(r["spectra"], r["records"]):
        assert a_f == rec["a_f"]
        assert np.array_equal(nm, rec["nmodes"])
        sel = rec["nmodes"] > 0
        np.testing.assert_allclose(k[sel], rec["k"][sel], rtol=1e-12)
        np.testing.assert_allclose(p[sel], rec["p"][sel], rtol=1e-5)      # BASELINE.json north_star


def test_variable_mesh_switch(ref_mod, pk_text):
    """vpm.c: pm_nc_factor = {{0, 1}, {0.5, 3}} switches the force mesh mid-run (BASELINE configs[3] in small)."""
    nc, L = 16, 48.0
    steps = np.linspace(0.1, 1.0, 6)
    r = _run_pair(ref_mod, pk_text, nc, L, [(0.0, 1), (0.5, 3)], "fastpm", "LCDM", steps)
    nbins = [len(s[1]) for s in r["spectra"]]
    assert nbins[0] == nc // 2 and nbins[-1] == 3 * nc // 2         # Nmesh = 16 -> 48
    xerr = _periodic_maxdiff(np.mod(r["x"], L), np.mod(r["ref"]["x"], L), L)
    assert xerr < 1e-4, xerr
    for (a_f, k, p, nm), rec in zip(r["spectra"], r["records"]):
        sel = rec["nmodes"] > 0
        np.testing.assert_allclose(p[sel], rec["p"][sel], rtol=1e-5)


def test_golden_lightcone_config(pk_text):
    """The reference's own golden log values (tests/run-test-lightcone.check:4-5,8,28,...,88) reproduced by the
    GPU path alone: nc=64, box 512, B=1, fastpm mode, LCDM growth, seed 100, remove_cosmic_variance.
    The white-noise field comes from the committed fixture tests/golden/lightcone_deltak.npz (made by
    tests/golden/make_fixtures.py from the reference's Gadget-scheme generator)."""
    import os, ctypes as C
    from fastpm_b200.solver import Solver, ForceEvent
    from fastpm_b200 import device
    fx = os.path.join(os.path.dirname(__file__), "golden", "lightcone_deltak.npz")
    dk = np.load(fx)["delta_k"]
    g = Solver(nc=64, boxsize=512.0, pm_nc_factor=1, force_mode="fastpm", growth_mode="LCDM", np_alloc_factor=2.0, compute_potential=True)
    g.setup_lpt(dk, 0.1)
    plin = []

    def on_force_after(solver_ptr, event_ptr, userdata):
        ev = C.cast(event_ptr, C.POINTER(ForceEvent)).contents
        k, p, nm = g.powerspectrum_of(ev.pm, ev.delta_k)
        kmax = 4 * 2 * np.pi / 512.0
        P = N = 0.0
        i = 0
        while i == 0 or (i < len(k) and k[i] <= kmax):             # fastpm_powerspectrum_large_scale, powerspectrum.c:170-186
            P += p[i] * nm[i]
            N += nm[i]
            i += 1
        P /= N
        plin.append(P / g.growth(ev.a_f)["D1"] ** 2)
        return 0

    g.add_handler("FORCE", 1, on_force_after)
    g.evolve(np.linspace(0.1, 1, 8))
    golden = [17305.5, 17200.9, 17110, 17064.7, 17043.4, 17028.1, 17014.2, 17002.2]
    assert ["%g" % v for v in plin] == ["%g" % v for v in golden], plin
    g.close()
