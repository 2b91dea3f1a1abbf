"""Worker for the multi-process tests (launched with torch.distributed.run, one process per rank).

mode "cpu":  world_size-2 gloo test of the host plumbing (no GPU): callback collectives, slab ownership.
mode "gpu":  x-slab run on one GPU per rank from the committed fixture; rank 0 gathers the particles by id and checks
             them against the reference fixture (tests/golden/small_run.npz).
"""
import os
import sys
import ctypes as C
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


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
        if rank == 0:
            print("MP_CPU_OK")
        return
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    lib = _lib.require_device(local)
    multigpu.init_comm(lib)
    from fastpm_b200.solver import Solver
    fx = np.load(os.path.join(ROOT, "tests", "golden", "small_run.npz"))
    L = 32.0
    os.environ.setdefault("FASTPM_B200_MIGRATE_FRAC", "1.0")
    g = Solver(nc=16, boxsize=L, pm_nc_factor=2, force_mode="fastpm", growth_mode="LCDM", np_alloc_factor=3.0)
    g.setup_lpt(fx["delta_k"], 0.1)
    g.evolve(fx["steps"])
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
        print("MP_GPU_OK ranks=%d max position error %.3g Mpc/h, np per rank %s" % (world, err, [len(o[0]) for o in out]))
    g.close()
    dist.barrier()


if __name__ == "__main__":
    main()
