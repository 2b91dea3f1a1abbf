"""End-to-end parity on the configurations and kernels the bench times.

BASELINE.json configs[1] ("C1"): nc = 256^3 particles, 512^3 mesh (B = 2), box 256 Mpc/h, COLA, 10 steps linspace(0.1, 1, 10),
the reference's Gadget-scheme initial conditions for seed 100 -- the GPU run against the compiled reference (oracle/_ref) on the
same delta_k.  Mesh sizes >= 512 take the TMA tile pass (fft_tma_kernel), the bulk-copy row pass (fft_zrow_kernel) and the
float-float fused Green's function; the tests assert through fpm_path_counts that those kernels -- not the generic
shared-memory passes the small solver tests exercise -- served the run.  A second case puts the 1024^3-mesh instantiation and the
Lagrangian brick walk of paint / readout under the same oracle comparison.

Tolerances: P(k) <= 1e-5 relative per bin (BASELINE.json).  Positions (periodic distance, matched by id): BASELINE.json's
1e-4 Mpc/h is what the 4096-particle runs of test_gpu_solver.py meet with a wide margin (3e-6); at 16.7 million particles and ten
steps the REFERENCE does not meet it against ITSELF -- its float32 deposit is an unordered `omp atomic`, and the compiled reference
run with 8 and with 5 threads differs by up to 1.16e-4 Mpc/h (99.99 % of the particles within 9.5e-6, mean 5.6e-7; P(k) 7e-8;
scripts/ref_scatter.py).  The test therefore asks for 99.99 % of the particles within 1e-4 Mpc/h and for the worst particle
within 5e-4 (a few times the reference's own worst case), and prints what it found.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATHS = ["fft_tma", "fft_tile_generic", "fft_zrow", "fft_z_generic", "fft_tma_multi", "paint_bricks", "readout_bricks", "pk_rows",
         "staged_transpose", "paint_tiles", "readout_tiles", "readout3"]


def path_counts(lib):
    out = (C.c_uint64 * len(PATHS))()
    lib.fpm_path_counts.argtypes = [C.c_void_p, C.c_int]
    lib.fpm_path_counts(out, len(PATHS))
    return dict(zip(PATHS, [int(v) for v in out]))


def _pdist(a, b, L):
    d = np.abs(np.mod(a, L) - np.mod(b, L))
    return np.minimum(d, L - d)


def run_pair(ref_mod, pk_text, nc, L, B, steps, hint=None, mode="cola"):
    from fastpm_b200 import _lib
    from fastpm_b200.solver import Solver, ForceEvent
    s = ref_mod.Session(nc=nc, boxsize=L, pm_nc_factor=B, force_mode=mode, growth_mode="LCDM", np_alloc_factor=2.0)
    dk, _, _ = s.ic_deltak(100, pk_text)
    s.setup_lpt(dk, steps[0])
    s.evolve(steps)
    want, recs = s.get_particles(), s.records()
    s.close()

    g = Solver(nc=nc, boxsize=L, pm_nc_factor=B, force_mode=mode, growth_mode="LCDM", np_alloc_factor=1.0)
    lib = _lib.require_device()
    g.setup_lpt(dk, steps[0])
    if hint is not None:
        _lib.check(lib.fpm_particle_grid_hint(hint), "hint")
    spectra = []

    def on_force_after(solver_ptr, event_ptr, userdata):
        ev = C.cast(event_ptr, C.POINTER(ForceEvent)).contents
        spectra.append((ev.a_f,) + g.powerspectrum_of(ev.pm, ev.delta_k))
        return 0

    g.add_handler("FORCE", 1, on_force_after)
    before = path_counts(lib)
    g.evolve(steps)
    after = path_counts(lib)
    x, v, ids = g.get_column("x"), g.get_column("v"), g.get_column("id")
    g.close()
    if hint is not None:
        _lib.check(lib.fpm_particle_grid_hint(0), "hint")
    used = {k: after[k] - before[k] for k in PATHS}
    assert np.array_equal(ids, want["id"])
    return dict(x=x, v=v, want=want, recs=recs, spectra=spectra, used=used)


def check_pair(r, L, nsteps):
    d = _pdist(r["x"], r["want"]["x"], L).max(axis=1)
    xerr, q = d.max(), np.quantile(d, 0.9999)
    print("position error: max %.3g, 99.99 %% quantile %.3g, mean %.3g Mpc/h" % (xerr, q, d.mean()))
    assert q < 1e-4 and xerr < 5e-4, "position error: max %.3g Mpc/h, 99.99 %% quantile %.3g" % (xerr, q)
    assert np.abs(r["v"] - r["want"]["v"]).max() < 2e-4 * np.abs(r["want"]["v"]).max()
    assert len(r["spectra"]) == len(r["recs"]) == nsteps
    worst = 0.0
    for (a_f, k, p, nm), rec in zip(r["spectra"], r["recs"]):
        assert a_f == rec["a_f"]
        assert np.array_equal(nm, rec["nmodes"])
        sel = rec["nmodes"] > 0
        np.testing.assert_allclose(k[sel], rec["k"][sel], rtol=1e-10)       # mean |k| of up to 5e8 modes, summed in another order
        worst = max(worst, np.abs(p[sel] / rec["p"][sel] - 1).max())
        np.testing.assert_allclose(p[sel], rec["p"][sel], rtol=1e-5)
    return xerr, worst


def test_c1_matches_reference(ref_mod, pk_text):
    """BASELINE configs[1] in full: 10 COLA steps at nc = 256 / N = 512 through the TMA + bulk-copy FFT passes."""
    nc, L, B = 256, 256.0, 2
    steps = np.linspace(0.1, 1.0, 10)
    r = run_pair(ref_mod, pk_text, nc, L, B, steps)
    u = r["used"]
    assert u["fft_tma"] >= 8 * len(steps) and u["fft_zrow"] >= 4 * len(steps), u       # 4 transforms per force evaluation
    assert u["fft_tile_generic"] == 0 and u["fft_z_generic"] == 0, u
    assert u["pk_rows"] >= len(steps), u                 # the row-streaming P(k) kernel (the shared-memory tiles of paint / readout are opt-in)
    xerr, perr = check_pair(r, L, len(steps))
    print("C1: max position error %.3g Mpc/h, max P(k) deviation %.3g, paths %s" % (xerr, perr, u))


def test_large_mesh_and_brick_walk_match_reference(ref_mod, pk_text):
    """The 1024^3-mesh instantiation (fft_tma_kernel<16,16,4,16>, fft_zrow_kernel<8,8,8>) with the Lagrangian brick walk of the
    deposit and the gather forced on: nc = 256, B = 4, the first two entries of the C1 time table (2 force evaluations, one
    kick-drift-kick cycle) -- the oracle's 1024^3 transforms take about a minute each on the host."""
    nc, L, B = 256, 256.0, 4
    steps = np.linspace(0.1, 1.0, 10)[:2]
    r = run_pair(ref_mod, pk_text, nc, L, B, steps, hint=-nc)
    u = r["used"]
    assert u["fft_tma"] >= 8 * len(steps) and u["fft_zrow"] >= 4 * len(steps) and u["fft_tile_generic"] == 0, u
    # B = 4: a brick of 8 particles spans 32 cells, too wide for the shared-memory tiles -> the brick walk with global reductions
    assert u["paint_bricks"] >= len(steps) and u["readout_bricks"] >= 3 * len(steps) and u["paint_tiles"] == 0, u
    xerr, perr = check_pair(r, L, len(steps))
    print("N=1024 + bricks: max position error %.3g Mpc/h, max P(k) deviation %.3g, paths %s" % (xerr, perr, u))


_ONE_GPU_DIR = None


def _one_gpu_run(pk_text):
    """C1 on one GPU (this process), stored for the workers: delta_k.npy (the oracle's when it is built, else the device's own
    Gadget-scheme field read back), one_gpu.npz (particles and every step's P(k))."""
    global _ONE_GPU_DIR
    if _ONE_GPU_DIR:
        return _ONE_GPU_DIR
    import tempfile
    from fastpm_b200 import _lib
    from fastpm_b200.solver import Solver, ForceEvent
    from oracle import ref
    nc, L, B = 256, 256.0, 2
    steps = np.linspace(0.1, 1.0, 10)
    if not ref.available():
        pytest.skip("oracle/_ref not built: no delta_k to start both runs from")
    s = ref.Session(nc=nc, boxsize=L, pm_nc_factor=B, force_mode="cola", growth_mode="LCDM", np_alloc_factor=1.0)
    dk, _, _ = s.ic_deltak(100, pk_text)
    s.close()
    g = Solver(nc=nc, boxsize=L, pm_nc_factor=B, force_mode="cola", growth_mode="LCDM", np_alloc_factor=1.0)
    g.setup_lpt(dk, steps[0])
    spectra = []

    def on_force_after(solver_ptr, event_ptr, userdata):
        ev = C.cast(event_ptr, C.POINTER(ForceEvent)).contents
        spectra.append(g.powerspectrum_of(ev.pm, ev.delta_k))
        return 0

    g.add_handler("FORCE", 1, on_force_after)
    g.evolve(steps)
    d = tempfile.mkdtemp(prefix="fpm_c1_")
    np.save(os.path.join(d, "delta_k.npy"), dk)
    np.savez(os.path.join(d, "one_gpu.npz"), nc=nc, L=L, B=B, steps=steps, id=g.get_column("id"), x=g.get_column("x"), v=g.get_column("v"),
             p=np.array([sp[1] for sp in spectra]), nmodes=np.array([sp[2] for sp in spectra]))
    g.close()
    _ONE_GPU_DIR = d
    return d


@pytest.mark.parametrize("stage", ["staged", "direct"])
def test_c1_two_gpus_match_one_gpu(pk_text, stage):
    """The same C1 run on 2 GPUs (MULTI instantiation of the TMA pass; staged and direct slab transposes) against the one-GPU run:
    particles matched by id, P(k) bin by bin (tests/mp_c1_worker.py; rank 0 also runs the one-GPU reference of the comparison)."""
    from fastpm_b200 import _lib
    if _lib.load().fpm_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, MP_C1_DIR=_one_gpu_run(pk_text))
    if stage == "direct":
        env["FASTPM_B200_NO_STAGE"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29671" if stage == "staged" else "29673", os.path.join(ROOT, "tests", "mp_c1_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500)
    print(r.stdout[-3000:])
    assert "MP_C1_OK" in r.stdout, r.stdout[-4000:]
