"""Worker for the multi-process tests (launched with torch.distributed.run, one process per rank).

mode "cpu":  world_size-2 gloo test of the host plumbing (no GPU): callback collectives, slab ownership.
mode "emul": the "gpu" mode on the CPU, against the emulated build of the whole library (tests/test_cpu_full_emulation.py).
mode "gpu":  x-slab run on one GPU per rank from the committed fixture; rank 0 gathers the particles by id and checks
             them against the reference fixture (tests/golden/small_run.npz).
"""
import os
import sys
import ctypes as C
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def snapshot_two_writers(lib, allreduce, allgather, rank, world):
    """Row N2: every rank writes its slice of a catalog and its slab of a k-space mesh (csrc/host/io.c); the result is the
    directory the compiled reference writes from one rank, byte for byte (needs oracle/_ref; skipped without it)."""
    import filecmp
    import shutil
    import tempfile
    import torch
    import torch.distributed as dist
    from oracle import ref
    if not ref.available():
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_cpu_snapshot_io import IoColumn, IoMeta, _files
    box = [tempfile.mkdtemp(prefix="fpm_mp_io_") if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    tmp = box[0]
    nc, L = 8, 32.0
    if rank == 0:
        s = ref.Session(nc=nc, boxsize=L, pm_nc_factor=2, force_mode="fastpm", growth_mode="LCDM", np_alloc_factor=2.0)
        dk, _, _ = s.ic_deltak(7, open(os.path.join(ROOT, "tests", "golden", "powerspec.txt")).read())
        s.setup_lpt(dk, 0.1)
        s.write_snapshot(os.path.join(tmp, "ref"))
        s.write_complex(dk, os.path.join(tmp, "ref_dk"), "LinearDensityK", which=1)
        x, v = s.snapshot_particles()
        p = s.get_particles()
        M0 = p["meta"][2]
        pack = [x, v, p["id"], np.ascontiguousarray(np.transpose(s.complex_view(dk, which=1), (1, 0, 2))), float(M0)]
        s.close()
    else:
        pack = None
    box = [pack]
    dist.broadcast_object_list(box, src=0)
    x, v, ids, rows_all, M0 = box[0]
    lib.fastpm_b200_comm_init_host.argtypes = [C.c_int, C.c_int, type(allreduce), type(allgather), C.c_void_p]
    lib.fastpm_b200_comm_init_host(rank, world, allreduce, allgather, None)
    n = len(x)
    lo, hi = rank * n // world + (3 if rank else 0), (rank + 1) * n // world + (3 if rank + 1 < world else 0)      # uneven slices
    xs, vs, js = np.ascontiguousarray(x[lo:hi]), np.ascontiguousarray(v[lo:hi]), np.ascontiguousarray(ids[lo:hi])
    cols = (IoColumn * 3)(IoColumn(b"Position", b"f4", b"f8", 3, xs.ctypes.data, 0), IoColumn(b"Velocity", b"f4", b"f4", 3, vs.ctypes.data, 0),
                          IoColumn(b"ID", b"i8", b"i8", 1, js.ctypes.data, 0))
    m = IoMeta((C.c_int64 * 3)(nc * nc, nc, 1), (C.c_double * 3)(L / nc, L / nc, L / nc), (C.c_double * 3)(0, 0, 0), nc ** 3, 0.1, 0.1, M0)
    mine = os.path.join(tmp, "mine")
    lib.fastpm_b200_io_write_columns.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int]
    assert lib.fastpm_b200_io_write_columns(mine.encode(), b"1", cols, 3, hi - lo, C.byref(m), 1) == 0
    hc = nc // 2 + 1
    pitch_c = ((hc + 15) // 16) * 16
    nyl, y0 = nc // world, rank * (nc // world)
    rows = np.zeros((nyl, nc, pitch_c), dtype=np.complex64)
    rows[:, :, :hc] = rows_all[y0:y0 + nyl]
    lib.fastpm_b200_io_write_complex_rows.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
    assert lib.fastpm_b200_io_write_complex_rows(os.path.join(tmp, "mine_dk").encode(), b"LinearDensityK", nc, L, y0, nyl, rows.ctypes.data, pitch_c, 1, 1) == 0
    dist.barrier()
    if rank == 0:
        for mine_d, ref_d, skip in ((mine, os.path.join(tmp, "ref"), "Header"), (os.path.join(tmp, "mine_dk"), os.path.join(tmp, "ref_dk"), None)):
            want = [f for f in _files(ref_d) if not (skip and f.startswith(skip))]
            assert _files(mine_d) == want, (_files(mine_d), want)
            for f in want:
                assert filecmp.cmp(os.path.join(mine_d, f), os.path.join(ref_d, f), shallow=False), f
        shutil.rmtree(tmp, ignore_errors=True)
    lib.fastpm_b200_comm_init_host(0, 1, allreduce, allgather, None)


def main():
    mode = sys.argv[1]
    import torch
    import torch.distributed as dist
    from fastpm_b200 import _lib, multigpu
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if mode == "cpu":
        dist.init_process_group(backend="gloo")
        lib = _lib.load()
        a, g = multigpu.make_callbacks()
        lib.fastpm_b200_comm_selftest.argtypes = [C.c_int, C.c_int, multigpu.ALLREDUCE, multigpu.ALLGATHER, C.c_void_p]
        rc = lib.fastpm_b200_comm_selftest(rank, world, a, g, None)
        assert rc == 0, "comm selftest failed with code %d" % rc
        # slab ownership: every x belongs to exactly one rank and the slabs tile the box
        lib.fastpm_b200_slab_owner.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int]
        L, n = 100.0, 64
        xs = np.concatenate([np.linspace(-L, 2 * L, 997), [0.0, L, np.nextafter(L, 0), L / 2]])
        owners = np.array([lib.fastpm_b200_slab_owner(float(x), L, n, world) for x in xs])
        want = (np.floor(xs * (1.0 / (L / n))).astype(np.int64) % n) // (n // world)
        assert np.array_equal(owners, want)
        mine = int((owners == rank).sum())
        t = torch.tensor([mine])
        dist.all_reduce(t)
        assert int(t.item()) == len(xs)
        dist.barrier()
        snapshot_two_writers(lib, a, g, rank, world)
        dist.barrier()
        if rank == 0:
            print("MP_CPU_OK")
        return
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if mode == "emul":
        # the same slab run on the CPU: every rank is a process that loads the emulated build of the whole library
        # (tests/emul/emul_lib/), the peers' arenas are POSIX shared memory mapped through the stand-in CUDA IPC
        _lib.LIB_PATH = os.path.join(ROOT, "tests", "emul", "_build", "libfastpm_b200_emul.so")
        os.environ.setdefault("FASTPM_B200_ARENA_GB", "0.25")
        dist.init_process_group(backend="gloo")
    else:
        torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    lib = _lib.require_device(local)
    multigpu.init_comm(lib)
    # host scalars go through the shared segment of host/shmcoll.c on one host, through the callbacks when asked to
    want_shared = 0 if os.environ.get("FASTPM_B200_HOST_COLL") == "callbacks" else 1
    assert lib.fastpm_b200_host_collectives_shared() == want_shared, "host collectives: shared = %d" % lib.fastpm_b200_host_collectives_shared()
    from fastpm_b200.solver import Solver
    fx = np.load(os.path.join(ROOT, "tests", "golden", "small_run.npz"))
    L = 32.0
    os.environ.setdefault("FASTPM_B200_MIGRATE_FRAC", "1.0")
    g = Solver(nc=16, boxsize=L, pm_nc_factor=2, force_mode="fastpm", growth_mode="LCDM", np_alloc_factor=3.0)
    g.setup_lpt(fx["delta_k"], 0.1)
    cnt = (C.c_uint64 * 16)()
    lib.fpm_path_counts.argtypes = [C.c_void_p, C.c_int]
    lib.fpm_path_counts(cnt, 16)
    r3_before = cnt[11]                                      # "readout3" (tests/test_gpu_c1.py: PATHS)
    g.evolve(fx["steps"])
    lib.fpm_path_counts(cnt, 16)
    # staged + pipelined transposes with room for three canvases: one gather for the three force components per force evaluation
    if not (os.environ.get("FASTPM_B200_NO_STAGE") or os.environ.get("FASTPM_B200_NO_PIPELINE") or os.environ.get("FASTPM_B200_FUSED_READOUT") == "0"):
        assert cnt[11] - r3_before == len(fx["steps"]), (cnt[11] - r3_before, len(fx["steps"]))
    ids, x, v = g.get_column("id"), g.get_column("x"), g.get_column("v")
    out = [None] * world
    dist.all_gather_object(out, (ids, x, v))
    if rank == 0:
        ids = np.concatenate([o[0] for o in out]); x = np.concatenate([o[1] for o in out]); v = np.concatenate([o[2] for o in out])
        assert len(ids) == 16 ** 3 and len(np.unique(ids)) == 16 ** 3, "particles lost or duplicated in migration"
        order = np.argsort(ids)
        x, v = x[order], v[order]
        ref_order = np.argsort(fx["id"])
        d = np.abs(np.mod(x, L) - np.mod(fx["x1"][ref_order], L))
        err = np.minimum(d, L - d).max()
        assert err < 1e-4, "positions differ from the reference fixture by %g Mpc/h" % err
        assert np.abs(v - fx["v1"][ref_order]).max() < 1e-4 * np.abs(fx["v1"]).max()
        print("MP_GPU_OK ranks=%d max position error %.3g Mpc/h, np per rank %s, migration rounds <= %d" % (
            world, err, [len(o[0]) for o in out], lib.fastpm_b200_migrate_rounds_max()))
    g.close()
    dist.barrier()
    if os.environ.get("MP_EXTRAS"):
        extras(rank, world)
        dist.barrier()


def extras(rank, world):
    """Rows N3 and N2 on several ranks: the PGD correction (halo planes + distributed inverse transforms) against the oracle, and a
    snapshot written by every rank from its device columns (unsorted, like sort_snapshot = false) against the reference's."""
    import shutil
    import tempfile
    import torch.distributed as dist
    from fastpm_b200.solver import Solver
    from oracle import ref
    if not ref.available():
        return
    nc, L, par, steps = 16, 32.0, (0.2, 0.5, 1.0, 1.0, 5.0), np.linspace(0.1, 1.0, 4)
    kw = dict(nc=nc, boxsize=L, pm_nc_factor=2, force_mode="fastpm", growth_mode="LCDM", pgdc=par)
    box = [tempfile.mkdtemp(prefix="fpm_mp_extras_") if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    tmp = box[0]
    pack = None
    if rank == 0:
        s = ref.Session(np_alloc_factor=2.0, **kw)
        dk, _, _ = s.ic_deltak(9, open(os.path.join(ROOT, "tests", "golden", "powerspec.txt")).read())
        s.setup_lpt(dk, steps[0])
        s.evolve(steps)
        want, want_pgdc = s.get_particles(), s.get_pgdc()
        s.write_snapshot(os.path.join(tmp, "ref"))
        s.close()
        pack = dk
    box = [pack]
    dist.broadcast_object_list(box, src=0)
    dk = box[0]
    g = Solver(np_alloc_factor=3.0, **kw)
    g.setup_lpt(dk, steps[0])
    g.evolve(steps)
    g.write_snapshot(os.path.join(tmp, "mine"))
    g.write_snapshot(os.path.join(tmp, "mine_sorted"), sort_by_id=True)      # every row at the file position its id gives
    out = [None] * world
    dist.all_gather_object(out, (g.get_column("id"), g.get_column("x"), g.get_column("v"), g.get_column("pgdc")))
    # the public mesh calls on the slabs (fastpm_paint with its halo exchange, pm_r2c, the P(k) measurement, pm_c2r, readout): what a
    # user's own density / P(k) code does with libfastpm -- against the reference's paint -> r2c -> P(k) and readout on one rank
    nb = nc * 2 // 2
    kk, pp, nm = np.zeros(nb), np.zeros(nb), np.zeros(nb)
    dens = np.zeros(len(out[rank][0]), dtype=np.float32)
    g.lib.fastpm_b200_public_mesh_probe.restype = C.c_int64
    g.lib.fastpm_b200_public_mesh_probe.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    n_here = g.lib.fastpm_b200_public_mesh_probe(g.h, 1.0, kk.ctypes.data, pp.ctypes.data, nm.ctypes.data, dens.ctypes.data)
    assert n_here == len(dens)
    M0 = float(g.meta["M0"])                      # the store's particle mass: the reference side below paints unit masses
    probe = [None] * world
    dist.all_gather_object(probe, (out[rank][0], dens / M0, kk, pp / M0 ** 2, nm))
    g.close()
    if rank == 0:
        ids = np.concatenate([o[0] for o in out])
        order = np.argsort(ids)
        x, v, pg = (np.concatenate([o[i] for o in out])[order] for i in (1, 2, 3))
        ro = np.argsort(want["id"])
        d = np.abs(np.mod(x, L) - np.mod(want["x"][ro], L))
        err = np.minimum(d, L - d).max()
        assert err < 1e-4, err
        assert np.abs(v - want["v"][ro]).max() < 1e-4 * np.abs(want["v"]).max()
        assert np.abs(pg - want_pgdc[ro]).max() < 1e-4 * np.abs(want_pgdc).max()
        # public mesh chain: the same P(k) on every rank, equal to the reference's of OUR final particles; density at the particles
        s2 = ref.Session(np_alloc_factor=2.0, **kw)
        canvas = s2.paint(x, a=1.0)
        k0, p0, n0 = s2.powerspectrum(s2.r2c(canvas, a=1.0), a=1.0)
        d0 = s2.readout(canvas, x, a=1.0)
        s2.close()
        for r in range(world):
            assert np.array_equal(probe[r][4], n0)
            sel = n0 > 0
            np.testing.assert_allclose(probe[r][3][sel], p0[sel], rtol=2e-5)
        pid = np.concatenate([q[0] for q in probe])
        pd = np.concatenate([q[1] for q in probe])[np.argsort(pid)]
        assert np.abs(pd - d0).max() < 2e-5 * np.abs(d0).max(), np.abs(pd - d0).max()
        # the snapshot: same blocks and attributes, same particles once both are ordered by id
        mine, refd = os.path.join(tmp, "mine"), os.path.join(tmp, "ref")
        assert sorted(os.listdir(os.path.join(mine, "1"))) == sorted(os.listdir(os.path.join(refd, "1")))
        assert open(os.path.join(mine, "1", "attr-v2")).read() == open(os.path.join(refd, "1", "attr-v2")).read()
        rd = lambda top, name, dt, nm: np.fromfile(os.path.join(top, "1", name, "000000"), dtype=dt).reshape(-1, nm)
        ia, ib = rd(mine, "ID", np.uint64, 1)[:, 0], rd(refd, "ID", np.uint64, 1)[:, 0]
        assert np.array_equal(np.sort(ia), np.sort(ib)) and len(ia) == nc ** 3
        oa, ob = np.argsort(ia), np.argsort(ib)
        dd = np.abs(rd(mine, "Position", np.float32, 3)[oa].astype(np.float64) - rd(refd, "Position", np.float32, 3)[ob])
        assert np.minimum(dd, L - dd).max() < 1e-4
        vb = rd(refd, "Velocity", np.float32, 3)[ob]
        assert np.abs(rd(mine, "Velocity", np.float32, 3)[oa] - vb).max() < 1e-4 * np.abs(vb).max()
        # sorted by id on several ranks == the one-rank reference file row by row (its store is still in id order)
        srt = os.path.join(tmp, "mine_sorted")
        assert np.array_equal(ib, np.arange(nc ** 3, dtype=np.uint64))
        assert open(os.path.join(srt, "1", "ID", "000000"), "rb").read() == open(os.path.join(refd, "1", "ID", "000000"), "rb").read()
        assert open(os.path.join(srt, "1", "ID", "header")).read() == open(os.path.join(refd, "1", "ID", "header")).read()
        dd = np.abs(rd(srt, "Position", np.float32, 3).astype(np.float64) - rd(refd, "Position", np.float32, 3))
        assert np.minimum(dd, L - dd).max() < 1e-4
        assert np.abs(rd(srt, "Velocity", np.float32, 3) - rd(refd, "Velocity", np.float32, 3)).max() < 1e-4 * np.abs(vb).max()
        shutil.rmtree(tmp, ignore_errors=True)
    # row N1 on several ranks: the device IC chain (every rank fills its ky-slab of the white noise) against the reference's
    kw = dict(nc=nc, boxsize=64.0, pm_nc_factor=2, force_mode="cola", growth_mode="LCDM")
    pk_path = os.path.join(ROOT, "tests", "golden", "powerspec.txt")
    tab = np.loadtxt(pk_path)
    g = Solver(np_alloc_factor=3.0, **kw)
    g.setup_ic(2024, tab[:, 0], tab[:, 1], 0.1, remove_variance=True)
    out = [None] * world
    dist.all_gather_object(out, (g.get_column("id"), g.get_column("x"), g.get_column("dx1"), g.get_column("dx2")))
    g.close()
    if rank == 0:
        s = ref.Session(np_alloc_factor=2.0, **kw)
        dk, _, _ = s.ic_deltak(2024, open(pk_path).read(), remove_variance=True)
        s.setup_lpt(dk, 0.1)
        w = s.get_particles()
        s.close()
        ids = np.concatenate([o[0] for o in out])
        assert np.array_equal(np.sort(ids), np.sort(w["id"]))
        order, ro = np.argsort(ids), np.argsort(w["id"])
        x, d1, d2 = (np.concatenate([o[i] for o in out])[order] for i in (1, 2, 3))
        dd = np.abs(x - w["x"][ro])
        ic_err = np.minimum(dd, 64.0 - dd).max()
        assert ic_err < 1e-5, ic_err
        assert np.abs(d1 - w["dx1"][ro]).max() < 1e-5 * np.abs(w["dx1"]).max() and np.abs(d2 - w["dx2"][ro]).max() < 1e-4 * np.abs(w["dx2"]).max()
        print("MP_EXTRAS_OK ranks=%d PGD position error %.3g Mpc/h, IC position error %.3g Mpc/h" % (world, err, ic_err))
    # row N4 on several ranks: the quadratic and a Lanczos window, whose reach beyond the slab (1 + 2 and 2 + 3 mesh planes) is kept
    # in a halo block and exchanged with both neighbours (host/gravity.c, csrc/comm.cu) where the reference moves ghost particles;
    # one kick-drift-kick cycle (two force evaluations) against the reference, like the one-GPU case of first_gpu_run_cases.py
    short = np.linspace(0.1, 1.0, 4)[:2]
    errs = {}
    for painter, support in (("quad", 3), ("lanczos", 6)):
        kw = dict(nc=nc, boxsize=L, pm_nc_factor=2, force_mode="fastpm", growth_mode="LCDM", painter=painter, painter_support=support)
        box = [None]
        if rank == 0:
            s = ref.Session(np_alloc_factor=2.0, **kw)
            dk, _, _ = s.ic_deltak(11, open(pk_path).read())
            s.setup_lpt(dk, short[0])
            s.evolve(short)
            box = [(dk, s.get_particles())]
            s.close()
        dist.broadcast_object_list(box, src=0)
        dk, want = box[0]
        g = Solver(np_alloc_factor=3.0, **kw)
        g.setup_lpt(dk, short[0])
        g.evolve(short)
        out = [None] * world
        dist.all_gather_object(out, (g.get_column("id"), g.get_column("x"), g.get_column("v")))
        g.close()
        if rank == 0:
            ids = np.concatenate([o[0] for o in out])
            assert np.array_equal(np.sort(ids), np.sort(want["id"]))
            order, ro = np.argsort(ids), np.argsort(want["id"])
            x, v = (np.concatenate([o[i] for o in out])[order] for i in (1, 2))
            dd = np.abs(np.mod(x, L) - np.mod(want["x"][ro], L))
            errs[painter] = np.minimum(dd, L - dd).max()
            assert errs[painter] < 2e-5, (painter, errs[painter])
            verr = np.abs(v - want["v"][ro]).max() / np.abs(want["v"]).max()
            assert verr < 1e-4, (painter, verr)
    if rank == 0:
        print("MP_WINDOWS_OK ranks=%d one kick-drift-kick: quadratic %.3g, Lanczos(6) %.3g Mpc/h from the reference" % (world, errs["quad"], errs["lanczos"]))


if __name__ == "__main__":
    main()
