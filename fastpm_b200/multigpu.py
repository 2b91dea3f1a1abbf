"""One process per GPU: joins the C library's communicator to torch.distributed and runs the multi-GPU bench.

torch.distributed (NCCL on GPUs, gloo on CPU) is the plumbing: rendezvous, host scalars, counts, CUDA-IPC handles.
Mesh planes and particles move GPU to GPU inside the library's own kernels (csrc/comm.cu, csrc/fft*.cu).
"""
import ctypes as C
import os

import numpy as np

ALLREDUCE = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p)
ALLGATHER = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p)

_keep = []


def make_callbacks(group=None):
    """ctypes callbacks implementing host all-reduce (sum/min/max on f64 or i64) and all-gather of bytes."""
    import torch
    import torch.distributed as dist
    ops = {0: dist.ReduceOp.SUM, 1: dist.ReduceOp.MIN, 2: dist.ReduceOp.MAX}
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")

    def allreduce(buf, count, is_int64, op, userdata):
        ct = C.c_int64 if is_int64 else C.c_double
        arr = np.ctypeslib.as_array(C.cast(buf, C.POINTER(ct)), (count,))
        t = torch.from_numpy(arr.copy()).to(dev)
        dist.all_reduce(t, op=ops[op], group=group)
        arr[:] = t.cpu().numpy()

    def allgather(send, nbytes, recv, userdata):
        world = dist.get_world_size(group)
        src = np.ctypeslib.as_array(C.cast(send, C.POINTER(C.c_uint8)), (nbytes,))
        t = torch.from_numpy(src.copy()).to(dev)
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t, group=group)
        dst = np.ctypeslib.as_array(C.cast(recv, C.POINTER(C.c_uint8)), (nbytes * world,))
        for r, o in enumerate(outs):
            dst[r * nbytes:(r + 1) * nbytes] = o.cpu().numpy()

    a, g = ALLREDUCE(allreduce), ALLGATHER(allgather)
    _keep.extend([a, g])
    return a, g


def init_comm(lib):
    """Call once per process after torch.distributed.init_process_group and before creating a Solver."""
    import torch.distributed as dist
    host_group = dist.new_group(backend="gloo")          # host scalars and IPC handles: CPU tensors, no stream sync
    a, g = make_callbacks(host_group)
    lib.fastpm_b200_comm_init.argtypes = [C.c_int, C.c_int, ALLREDUCE, ALLGATHER, C.c_void_p]
    lib.fastpm_b200_comm_init(dist.get_rank(), dist.get_world_size(), a, g, None)


def bench_main(args):
    import json
    import time
    import torch
    import torch.distributed as dist
    from . import _lib
    from .solver import Solver, ForceEvent
    import bench as B

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    lib = _lib.require_device(local)
    init_comm(lib)

    nc, Bf, K, W = args.nc, args.pm_nc_factor, args.steps, args.warmup
    N = nc * Bf
    Np = nc ** 3
    k_tab, p_tab = B.read_pk()
    ts = np.linspace(0.1, 1.0, K)
    alloc = float(os.environ.get("FASTPM_B200_ALLOC_FACTOR", "1.25"))
    g = Solver(nc=nc, boxsize=float(nc), pm_nc_factor=Bf, force_mode=args.mode, growth_mode="LCDM", np_alloc_factor=alloc)
    B.setup_ic(g, k_tab, p_tab, ts[0])
    meta0 = g.meta
    np0 = g.np
    cola = args.mode == "cola"
    cols_in = ["x", "v", "id"] + (["dx1", "dx2"] if cola else [])
    cols_out = ["x", "v", "id"]
    itemsize = dict(x=24, v=12, id=8, dx1=12, dx2=12)
    host = {}
    cap = int(np0 * alloc) + 1
    for c in set(cols_in + cols_out):                    # pinned host buffers: this rank's slab of the initial state
        ptr = lib.fpm_host_alloc_pinned(cap * itemsize[c])
        if not ptr:
            raise RuntimeError("pinned host allocation failed: " + lib.fpm_last_error().decode())
        host[c] = ptr
    for c in cols_in:
        _lib.check(lib.fpm_memcpy_d2h(host[c], g.column_ptr(c), np0 * itemsize[c]), "save IC")

    def restore():
        g.set_np(np0)
        for c in cols_in:
            _lib.check(lib.fpm_memcpy_h2d(g.column_ptr(c), host[c], np0 * itemsize[c]), "restore IC")
        g.set_meta(meta0["a_x"], meta0["a_v"], meta0["M0"])

    spectra = []

    def on_force_after(solver_ptr, event_ptr, userdata):
        ev = C.cast(event_ptr, C.POINTER(ForceEvent)).contents
        spectra.append(g.powerspectrum_of(ev.pm, ev.delta_k))
        return 0

    g.add_handler("FORCE", 1, on_force_after)
    if W >= 1:                                           # warm-up on the first entries of the same table, then the ICs again
        g.evolve(ts[:max(W, 2)])
        restore()
        spectra.clear()

    # ---- device-resident timed run: barrier + device sync on both sides, CUDA events on the library stream, max over ranks
    timer = C.c_void_p()
    _lib.check(lib.fpm_timer_create(C.byref(timer)))
    lib.fpm_prof_reset()
    lib.fpm_prof_enable(1)
    launches0 = int(lib.fpm_kernel_launch_count())
    nvl0 = (C.c_uint64 * 4)()
    lib.fpm_comm_byte_counts.argtypes = [C.c_void_p]
    lib.fpm_comm_byte_counts(nvl0)
    sampler = B.ClockSampler(local) if rank == 0 else None
    _lib.check(lib.fpm_sync())
    torch.cuda.synchronize()
    dist.barrier()
    lib.fpm_timer_start(timer)
    g.evolve(ts)
    lib.fpm_timer_stop(timer)
    ms = C.c_double()
    _lib.check(lib.fpm_timer_elapsed_ms(timer, C.byref(ms)))
    _lib.check(lib.fpm_sync())
    torch.cuda.synchronize()
    dist.barrier()
    clocks = sampler.stop() if sampler else None
    launches = int(lib.fpm_kernel_launch_count()) - launches0
    nvl1 = (C.c_uint64 * 4)()
    lib.fpm_comm_byte_counts(nvl1)
    nvl = [int(b) - int(a) for a, b in zip(nvl0, nvl1)]
    lib.fpm_prof_enable(0)
    counts = (C.c_int64 * len(B.KCLASSES))()
    totals = (C.c_double * len(B.KCLASSES))()
    _lib.check(lib.fpm_prof_get(counts, totals, len(B.KCLASSES)))
    t = torch.tensor([ms.value], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                 # max over ranks of the device time
    t_evolve = float(t.item()) / 1e3
    np_local = torch.tensor([g.np], dtype=torch.int64, device="cuda")
    dist.all_reduce(np_local)
    stages = {nm: {"launches": int(counts[i]), "ms": round(float(totals[i]), 3)} for i, nm in enumerate(B.KCLASSES)}

    # ---- end to end: every rank copies its slab of the initial state from pinned host memory to its GPU, evolves, and copies
    #      x, v, id of the particles it ends up with back to pinned host memory; wall clock between barriers, max over ranks
    dist.barrier()
    t0 = time.perf_counter()
    restore()
    g.evolve(ts)
    n_local = g.np
    for c in cols_out:
        _lib.check(lib.fpm_memcpy_d2h(host[c], g.column_ptr(c), n_local * itemsize[c]), "result d2h")
    _lib.check(lib.fpm_sync())
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e = float(te.item())
    h2d = sum(Np * itemsize[c] for c in cols_in)
    d2h = sum(Np * itemsize[c] for c in cols_out) + K * 3 * (N // 2) * 8
    xh = np.ctypeslib.as_array(C.cast(host["x"], C.POINTER(C.c_double)), (3 * n_local,))
    finite = bool(np.isfinite(xh[:: max(1, xh.size // 100000)]).all())
    fin = torch.tensor([1 if finite else 0], dtype=torch.int64, device="cuda")
    dist.all_reduce(fin, op=dist.ReduceOp.MIN)
    # the same fingerprints the one-GPU line prints: P(k) bins of the last step (already reduced over ranks by the library) and the
    # id-weighted position checksum, summed over the slabs
    idh = np.ctypeslib.as_array(C.cast(host["id"], C.POINTER(C.c_uint64)), (n_local,))
    chk = torch.tensor(B.position_checksum(xh, idh, float(nc)), dtype=torch.float64, device="cuda")
    dist.all_reduce(chk)

    S_local = 4.0 * N * N * (N + 2) / world
    peak, peak_src = B.measured_peak()
    tile = stages["fft_tile"]
    fftz = stages["fft_z"]
    ntr = max(1, fftz["launches"])
    # two strided passes per transform, each reads and writes this rank's share S_local once; the transposing one is launched in
    # chunks when it is staged (csrc/fft.cu staged_transpose), so bytes per LAUNCH = all the passes' bytes / all the launches
    tile_bytes = 2 * ntr * 2 * S_local
    avg_ms = tile["ms"] / max(1, tile["launches"])
    bytes_per_launch = tile_bytes / max(1, tile["launches"])
    achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    t_tr = (tile["ms"] + fftz["ms"]) / ntr
    line = {
        "metric": B.METRIC, "value": Np * K / t_evolve, "unit": B.UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": 1e3 * t_evolve / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 mesh / f64 positions", "data": "synthetic", "config": dict(B.workload_config(args), parallelism="x-slabs x%d" % world),
        "e2e": {"value": Np * K / t_e2e, "unit": B.UNIT, "h2d_bytes_per_step": int(h2d // K), "d2h_bytes_per_step": int(d2h // K),
                "seconds": round(t_e2e, 4)},
        "gpu_launches": launches, "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "fft_tma_kernel (strided FFT pass, per rank; the slab-transposing pass runs as chunk launches into a local staging mesh that the copy engines push to the peers)",
                     "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": None,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_ms": round(avg_ms, 4)},
        "fft": {"gbs_6S_per_gpu": round(6 * S_local / (t_tr * 1e-3) / 1e9, 1) if t_tr > 0 else 0.0, "ms_per_transform": round(t_tr, 4), "transforms": ntr},
        "host_collectives": "shared memory (host/shmcoll.c)" if lib.fastpm_b200_host_collectives_shared() else "launcher callbacks (torch.distributed)",
        "stages_rank0": stages, "np_total_after": int(np_local.item()), "result_finite": bool(fin.item() == 1),
        "pk_bins": [float(v) for v in spectra[-1][1][:8]] if spectra else None,
        "x_checksum": [float(v) for v in chk.tolist()],
        # bytes rank 0 moved over NVLink in the timed run, counted where the transfers are issued (nvidia-smi's NVLink counters read
        # N/A on these boxes); a slab transpose moves S/G * (G-1)/G out of every GPU
        "nvlink_rank0": {"transpose_push_bytes_per_transform": (nvl[0] + nvl[3]) / max(1, ntr), "expected_bytes_per_transform": S_local * (world - 1) / world,
                         "halo_bytes_per_step": nvl[1] / K, "migration_bytes_per_step": nvl[2] / K},
    }
    if rank == 0:
        B.emit(line)
    g.close()
    dist.barrier()
    return 0
