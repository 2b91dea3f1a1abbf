"""The whole library on the CPU: tests/emul/emul_lib/build.py compiles the host parts of csrc/*.cu and the C host layer together with
the emulated kernel sources (tests/emul/cuda_emul.h) and a stand-in CUDA runtime into tests/emul/_build/libfastpm_b200_emul.so.
The `gpu` cases of tests/first_gpu_run_cases.py -- written after this round's GPU budget was spent -- then run HERE, unchanged
except for a smaller particle grid, against the compiled reference: PGD correction, force softening, the other windows, the device
initial-condition chain, snapshots written from "device" columns, restart, snapshots during evolve, the two plain-C libfastpm user
programs.  Several ranks are several processes (arenas in POSIX shared memory): the 4-rank slab run.
What this does not cover: the TMA / bulk-copy FFT fast path (Nmesh >= 512; emulated kernel by kernel in test_cpu_oracle_and_host.py).

Test infrastructure only: the product loader never looks for this library (tests/conftest.py swaps the path when
FASTPM_B200_TEST_EMUL is set, in a pytest process of its own)."""
import os
import subprocess
import sys

import numpy as np

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_zz_first_gpu_run import CASES          # noqa: E402


@pytest.fixture(scope="module")
def emul_lib():
    if not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("CUDA headers not found")
    sys.path.insert(0, os.path.join(ROOT, "tests", "emul", "emul_lib"))
    import build
    return build.build()


# cases grouped so that the whole file stays around a minute on a few cores (the groups run as parallel processes)
GROUPS = [
    ["test_three_component_readout_equals_three_readouts", "test_device_gadget_ic_matches_reference", "test_single_mode_transfers_match_reference",
     "test_device_ic_chain_matches_reference", "test_pgd_correction_matches_reference"],
    ["test_force_softening_matches_reference", "test_non_cic_painter_matches_reference", "test_shifted_ics_match_reference"],
    ["test_passive_handler_keeps_the_fused_update", "test_store_fill_q_and_rand_columns_match_reference", "test_permute_by_dense_id_kernels",
     "test_wrap_and_summary_match_reference", "test_sorted_snapshot_of_a_shuffled_store"],
    ["test_snapshot_files_and_restart_match_reference", "test_snapshots_during_evolve_match_reference"],
    ["test_store_subsample_matches_reference", "test_store_copy_take_extend", "test_command_line_particle_fraction_snapshot"],      # the last one skips here
    ["test_cli_run_loop_program_matches_reference"],
]
# the two cases built on the committed nc = 16 fixture take 3 minutes each under emulation (all 17 cases: 8 minutes, all green on
# 2026-10-17): only with FASTPM_B200_TEST_EMUL_ALL=1
SLOW = ["test_libfastpm_user_program_runs_and_matches_fixture", "test_fused_readout_option_gives_the_same_run"]
if os.environ.get("FASTPM_B200_TEST_EMUL_ALL"):
    GROUPS += [[c] for c in SLOW]


def test_groups_cover_every_case():
    assert sorted([c for g in GROUPS for c in g] + ([] if os.environ.get("FASTPM_B200_TEST_EMUL_ALL") else SLOW)) == sorted(CASES)


@pytest.fixture(scope="module")
def runs(emul_lib, tmp_path_factory):
    """Everything below starts at once, as separate processes: the case groups and one bench.py run."""
    env = dict(os.environ, FASTPM_B200_TEST_EMUL="1", FASTPM_B200_TEST_NC_SCALE="0.25", OMP_NUM_THREADS="2")
    procs = {}
    for g in GROUPS:
        cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider", os.path.join(ROOT, "tests", "first_gpu_run_cases.py"),
               "-k", " or ".join(g)]
        procs[tuple(g)] = subprocess.Popen(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    script = tmp_path_factory.mktemp("bench") / "run_bench.py"
    script.write_text(
        "import sys, runpy\n"
        "sys.path.insert(0, %r)\n"
        "from fastpm_b200 import _lib\n"
        "_lib.LIB_PATH = %r\n"
        "sys.argv = ['bench.py', '--nc', '8', '--steps', '3', '--warmup', '1', '--no-cpu-baseline']\n"
        "runpy.run_path(%r, run_name='__main__')\n" % (ROOT, emul_lib, os.path.join(ROOT, "bench.py")))
    bench = subprocess.Popen([sys.executable, str(script)], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    shm_before = set(os.listdir("/dev/shm"))
    ranks4 = subprocess.Popen([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "4", "--master-addr", "127.0.0.1",
                               "--master-port", "29688", os.path.join(ROOT, "tests", "mp_worker.py"), "emul"],
                              cwd=ROOT, env=dict(os.environ, MP_EXTRAS="1"), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    # two ranks whose migration buffers hold 16 particles per destination: every decompose needs several exchange rounds
    ranks2 = subprocess.Popen([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                               "--master-port", "29687", os.path.join(ROOT, "tests", "mp_worker.py"), "emul"],
                              cwd=ROOT, env=dict(os.environ, FASTPM_B200_MIGRATE_CAP="16"), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    # the command line (fastpm_b200/lua_front) linked against the emulated library: one process, and `-n 2` (two forked ranks that find
    # each other through the shared segment their parent made -- no launcher, no callbacks)
    cli = {}
    if os.path.isdir("/root/reference/lua"):
        import shutil
        mk = subprocess.run(["make", "-C", os.path.join(ROOT, "fastpm_b200", "lua_front"), "emul"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert mk.returncode == 0, mk.stdout[-2000:]
        exe = os.path.join(ROOT, "fastpm_b200", "lua_front", "_build", "fastpm_b200_run_emul")
        for tag, extra, more in (("one", [], []), ("two", ["-n", "2"], []), ("sub", [], ["0.3"]), ("sub2", ["-n", "2"], ["0.3", "false"])):
            d = tmp_path_factory.mktemp("cli_" + tag)
            shutil.copy(os.path.join(ROOT, "tests", "golden", "powerspec.txt"), str(d))
            cli[tag] = (str(d), subprocess.Popen([exe] + extra + [os.path.join(ROOT, "tests", "lua", "small_nc16.lua"), "8", "3"] + more, cwd=str(d),
                                                 stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    out = {}
    for g, p in procs.items():
        o, _ = p.communicate(timeout=1500)
        out[g] = (p.returncode, o)
    bo, be = bench.communicate(timeout=1500)
    out["bench"] = (bench.returncode, bo, be)
    ro, _ = ranks4.communicate(timeout=1500)
    r2o, _ = ranks2.communicate(timeout=1500)
    out["ranks2_rounds"] = (ranks2.returncode, r2o)
    out["cli"] = {}
    for tag, (d, p) in cli.items():
        o, _ = p.communicate(timeout=1500)
        out["cli"][tag] = (p.returncode, d, o)
    out["cli_shm_left"] = sorted(set(os.listdir("/dev/shm")) - shm_before)      # evaluated when every process of this fixture has ended
    out["ranks4"] = (ranks4.returncode, ro, sorted(f for f in set(os.listdir("/dev/shm")) - shm_before if f.startswith("fpm_emul_")))
    return out


def test_gpu_cases_on_the_emulated_library(runs):
    failed = [(g, r[1][-3000:]) for g, r in runs.items() if isinstance(g, tuple) and r[0] != 0]       # the case groups
    assert not failed, "\n\n".join("%s\n%s" % (g, o) for g, o in failed)


def test_bench_contract_on_the_emulated_library(runs):
    """bench.py end to end (tiny grid, emulated library): one JSON line on stdout with the keys the driver reads."""
    import json
    rc, stdout, stderr = runs["bench"]
    assert rc == 0, stderr[-3000:]
    lines = [l for l in stdout.splitlines() if l.strip()]
    assert len(lines) == 1, stdout[-2000:]                     # exactly one line on stdout
    line = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert key in line, key
    assert line["metric"] == "pm_step_particles_per_second" and line["n_gpus"] == 1 and line["steps"] == 3 and line["warmup"] == 1
    assert line["value"] > 0 and line["gpu_launches"] > 0 and "workload" in line["config"]
    assert line["e2e"]["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in line["roofline"], key


def test_four_rank_slab_run_on_the_emulated_library(runs):
    """The x-slab run of tests/mp_worker.py ("gpu" mode on 2 and 8 B200s) with FOUR ranks -- a count no hardware run of this round
    covered -- as four processes on the emulated library: symmetric arenas in POSIX shared memory mapped through the stand-in CUDA
    IPC, the cross-GPU barrier kernel spinning on flags in the peers' memory, staged slab transposes, halo planes, particle
    migration; rank 0 gathers the particles and checks them against the reference fixture.  Then (MP_EXTRAS) the PGD correction on
    four slabs against the oracle and a snapshot written by every rank from its own columns."""
    rc, stdout, leftover = runs["ranks4"]
    assert "MP_GPU_OK ranks=4" in stdout, stdout[-3000:]
    assert "MP_EXTRAS_OK ranks=4" in stdout, stdout[-3000:]
    assert "MP_WINDOWS_OK ranks=4" in stdout, stdout[-3000:]          # quadratic / Lanczos windows: halo planes on both sides of a slab
    assert not leftover                                                   # the shared-memory arenas were unlinked


def test_migration_in_rounds_on_the_emulated_library(runs):
    """More leavers for one slab than a pack buffer holds (store.c:486-657 moves any number): the exchange runs in rounds,
    every rank the same number of them.  Two emulated ranks with 16-particle buffers reproduce the reference fixture."""
    import re
    rc, stdout = runs["ranks2_rounds"]
    m = re.search(r"MP_GPU_OK ranks=2 .* migration rounds <= (\d+)", stdout)
    assert m, stdout[-3000:]
    assert int(m.group(1)) > 1, stdout[-3000:]


def test_command_line_with_two_forked_ranks(runs):
    """fastpm_b200_run -n 2 (where the reference says `mpirun -n 2 fastpm`): the run loop of lua_front/run.c on two slabs against the same
    program on one -- same log lines of the power spectrum, same P(k) files, same snapshot (sorted by id) up to the float order of the
    deposits; nothing left in /dev/shm.  Linked against the emulated library: the program logic, not the GPU, is under test."""
    import re
    if not runs["cli"]:
        pytest.skip("needs /root/reference (the Lua runtime is compiled from there)")
    (rc1, d1, o1), (rc2, d2, o2) = runs["cli"]["one"], runs["cli"]["two"]
    assert rc1 == 0, o1[-3000:]
    assert rc2 == 0, o2[-3000:]
    pat = re.compile(r"D\^2\(([0-9.]+), 1.0\) P\(k<[0-9.]+\) = ([0-9.eE+-]+) Sigma8 = ([0-9.eE+-]+)")
    l1, l2 = pat.findall(o1), pat.findall(o2)
    assert len(l1) == 3 and len(l1) == len(l2), (l1, l2)
    for (a1, p1, s1), (a2, p2, s2) in zip(l1, l2):
        assert a1 == a2 and abs(float(p1) / float(p2) - 1) < 1e-5 and abs(float(s1) / float(s2) - 1) < 1e-5, (l1, l2)
    assert "ThisTask = 1" not in o2                                   # only rank 0 logs
    for f in sorted(os.listdir(os.path.join(d1, "out"))):
        if not f.endswith(".txt"):
            continue
        x, y = np.loadtxt(os.path.join(d1, "out", f)), np.loadtxt(os.path.join(d2, "out", f))
        assert x.shape == y.shape and np.array_equal(x[:, 2], y[:, 2]), f
        sel = x[:, 2] > 0
        np.testing.assert_allclose(x[sel, 1], y[sel, 1], rtol=1e-5, err_msg=f)
    rd = lambda top, name, dt, nm: np.fromfile(os.path.join(top, "out", "fastpm_1.0000", "1", name, "000000"), dtype=dt).reshape(-1, nm)
    ia, ib = rd(d1, "ID", np.uint64, 1)[:, 0], rd(d2, "ID", np.uint64, 1)[:, 0]
    assert np.array_equal(ia, np.arange(8 ** 3, dtype=np.uint64)) and np.array_equal(ia, ib)      # sorted by id, on two ranks as on one
    dd = np.abs(rd(d1, "Position", np.float32, 3).astype(np.float64) - rd(d2, "Position", np.float32, 3))
    assert np.minimum(dd, 16.0 - dd).max() < 1e-4
    va, vb = rd(d1, "Velocity", np.float32, 3), rd(d2, "Velocity", np.float32, 3)
    assert np.abs(va - vb).max() < 1e-4 * np.abs(va).max()
    hdr = lambda d: open(os.path.join(d, "out", "fastpm_1.0000", "Header", "attr-v2")).read()
    assert hdr(d1) == hdr(d2)
    assert not [f for f in runs["cli_shm_left"] if f.startswith("fastpm_b200_") or f.startswith("fpm_emul_")], runs["cli_shm_left"]


def test_command_line_particle_fraction(runs):
    """particle_fraction = 0.3 (src/fastpm.c:1449-1461): the snapshot holds exactly the particles the reference's
    fastpm_store_fill_subsample_mask / fastpm_store_subsample keep on the same grid (rand column = the rank's serial RANLUX stream),
    sorted by id, with the rows they have in the full snapshot; the rest of the run is untouched."""
    if not runs["cli"]:
        pytest.skip("needs /root/reference (the Lua runtime is compiled from there)")
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref/libfastpm_ref.so not built")
    (rc1, d1, o1), (rc3, d3, o3) = runs["cli"]["one"], runs["cli"]["sub"]
    assert rc1 == 0 and rc3 == 0, o3[-3000:]
    rd = lambda top, name, dt, nm: np.fromfile(os.path.join(top, "out", "fastpm_1.0000", "1", name, "000000"), dtype=dt).reshape(-1, nm)
    s = ref.Session(nc=8, boxsize=16.0, pm_nc_factor=2, np_alloc_factor=3.0)
    want_ids, _, _ = s.subsample_probe(3 * 8 ** 3, 0.3)
    s.close()
    ids = rd(d3, "ID", np.uint64, 1)[:, 0]
    assert 0.2 * 8 ** 3 < len(ids) < 0.4 * 8 ** 3 and np.array_equal(ids, want_ids)
    sel = ids.astype(np.int64)
    for name, dt, nm in (("Position", np.float32, 3), ("Velocity", np.float32, 3), ("DX1", np.float32, 3), ("DX2", np.float32, 3)):
        a, b = rd(d3, name, dt, nm).astype(np.float64), rd(d1, name, dt, nm)[sel].astype(np.float64)       # two runs: the float order of
        dd = np.abs(a - b)                                                                                   # the deposits differs
        if name == "Position":
            dd = np.minimum(dd, 16.0 - dd)
        assert dd.max() < 1e-4 * max(1.0, np.abs(b).max()), name
    assert sorted(os.listdir(os.path.join(d3, "out", "fastpm_1.0000", "1"))) == sorted(os.listdir(os.path.join(d1, "out", "fastpm_1.0000", "1")))
    hdr = open(os.path.join(d3, "out", "fastpm_1.0000", "Header", "attr-v2")).read()
    assert "ParticleFraction <f8 1" in hdr and "[ 0.3 ]" in hdr
    for f in sorted(os.listdir(os.path.join(d1, "out"))):                      # the power spectra do not know about the sub-sample
        if f.endswith(".txt"):
            x, y = np.loadtxt(os.path.join(d1, "out", f)), np.loadtxt(os.path.join(d3, "out", f))
            assert x.shape == y.shape and np.array_equal(x[:, 2], y[:, 2]), f
            np.testing.assert_allclose(x[x[:, 2] > 0, 1], y[x[:, 2] > 0, 1], rtol=1e-5, err_msg=f)


def test_command_line_particle_fraction_on_two_ranks(runs):
    """The same on two forked ranks (sort_snapshot = false: the distributed sort needs dense ids): every rank keeps the particles whose
    deviate of ITS OWN serial stream (store.c:694-720: the seed depends on the rank) is below the fraction -- the `rand` column having
    travelled with the particles through every migration of the run.  Expected set: the reference's streams for rank 0 and rank 1
    applied to the particles each rank was filled with."""
    if not runs["cli"]:
        pytest.skip("needs /root/reference (the Lua runtime is compiled from there)")
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref/libfastpm_ref.so not built")
    rc, d, o = runs["cli"]["sub2"]
    assert rc == 0, o[-3000:]
    ids = np.fromfile(os.path.join(d, "out", "fastpm_1.0000", "1", "ID", "000000"), dtype=np.uint64)
    nc, per_rank = 8, 8 ** 3 // 2
    s = ref.Session(nc=nc, boxsize=16.0, pm_nc_factor=2, np_alloc_factor=3.0)
    want = []
    for r in range(2):
        _, rnd = s.fill_probe(int(3.0 * nc ** 3 / 2), as_rank=r)
        want.append(np.nonzero(rnd[:per_rank] <= 0.3)[0] + r * per_rank)      # rank r was filled with the ids [r, r + 1) * nc^3 / 2
    s.close()
    want = np.concatenate(want).astype(np.uint64)
    assert len(ids) == len(want) == len(set(ids.tolist())) and np.array_equal(np.sort(ids), want)
    x = np.fromfile(os.path.join(d, "out", "fastpm_1.0000", "1", "Position", "000000"), dtype=np.float32).reshape(-1, 3)
    d1 = runs["cli"]["one"][1]
    full = np.fromfile(os.path.join(d1, "out", "fastpm_1.0000", "1", "Position", "000000"), dtype=np.float32).reshape(-1, 3)
    dd = np.abs(x.astype(np.float64) - full[ids.astype(np.int64)])
    assert np.minimum(dd, 16.0 - dd).max() < 1e-4                                # the rows are the ones of the full one-rank snapshot
