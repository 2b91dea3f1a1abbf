#!/usr/bin/env python
"""Runs one BASELINE.json configuration on the GPUs of this job and prints a JSON record (not the bench contract: no warm-up pass,
no host-buffer leg -- a capacity / correctness run).

    torchrun --nproc-per-node G scripts/run_config.py --nc 512 --pm-nc-factor "0:1,0.5:3" --steps 40 --mode fastpm      # configs[3]
    torchrun --nproc-per-node 8 scripts/run_config.py --nc 2048 --pm-nc-factor 2 --steps 20 --mode pm --alloc 1.15       # configs[4]

Record: ms per step (CUDA events on the library stream, max over ranks), particles/s, particles before / after (none lost in
migration), finite positions, the first P(k) bins of the last step, the id-weighted position checksum, per-class kernel times of
rank 0, which FFT paths served the run, device memory high-water mark."""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nc", type=int, required=True)
    ap.add_argument("--pm-nc-factor", default="2")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--mode", default="pm")
    ap.add_argument("--alloc", type=float, default=1.25)
    ap.add_argument("--a0", type=float, default=0.1)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import bench as B
    from fastpm_b200 import _lib
    from fastpm_b200.solver import Solver, ForceEvent
    if world > 1:
        import torch
        import torch.distributed as dist
        from fastpm_b200 import multigpu
        torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
        lib = _lib.require_device(local)
        multigpu.init_comm(lib)
    else:
        lib = _lib.require_device(local)
    pmf = args.pm_nc_factor
    factor = [tuple(float(v) for v in pr.split(":")) for pr in pmf.split(",")] if ":" in pmf else int(pmf)
    nc, K = args.nc, args.steps
    k_tab, p_tab = B.read_pk()
    ts = np.linspace(args.a0, 1.0, K)
    t0 = time.perf_counter()
    g = Solver(nc=nc, boxsize=float(nc), pm_nc_factor=factor, force_mode=args.mode, growth_mode="LCDM", np_alloc_factor=args.alloc if world > 1 else 1.0)
    g.setup_ic(B.IC_SEED, k_tab, p_tab, ts[0])
    _lib.check(lib.fpm_sync())
    t_ic = time.perf_counter() - t0
    np0 = g.np
    spectra, meshes = [], []

    def on_force_after(solver_ptr, event_ptr, userdata):
        ev = C.cast(event_ptr, C.POINTER(ForceEvent)).contents
        spectra.append(g.powerspectrum_of(ev.pm, ev.delta_k))
        meshes.append(len(spectra[-1][0]) * 2)
        return 0

    g.add_handler("FORCE", 1, on_force_after)
    timer = C.c_void_p()
    _lib.check(lib.fpm_timer_create(C.byref(timer)))
    lib.fpm_prof_reset()
    lib.fpm_prof_enable(1)
    cnt = (C.c_uint64 * 11)()
    lib.fpm_path_counts.argtypes = [C.c_void_p, C.c_int]
    lib.fpm_path_counts(cnt, 11)
    before = [int(v) for v in cnt]
    _lib.check(lib.fpm_sync())
    if world > 1:
        dist.barrier()
    lib.fpm_timer_start(timer)
    g.evolve(ts)
    lib.fpm_timer_stop(timer)
    ms = C.c_double()
    _lib.check(lib.fpm_timer_elapsed_ms(timer, C.byref(ms)))
    lib.fpm_prof_enable(0)
    lib.fpm_path_counts(cnt, 11)
    used = [int(v) - b for v, b in zip(cnt, before)]
    counts = (C.c_int64 * len(B.KCLASSES))()
    totals = (C.c_double * len(B.KCLASSES))()
    _lib.check(lib.fpm_prof_get(counts, totals, len(B.KCLASSES)))
    stages = {nm: {"launches": int(counts[i]), "ms": round(float(totals[i]), 1)} for i, nm in enumerate(B.KCLASSES) if counts[i]}
    # checksums from device columns in slices (no big host buffers)
    n_local = g.np
    chk = np.zeros(3)
    finite = True
    step = 1 << 24
    xbuf = np.empty((min(step, max(n_local, 1)), 3))
    ibuf = np.empty(min(step, max(n_local, 1)), dtype=np.uint64)
    for i in range(0, n_local, step):
        m = min(step, n_local - i)
        _lib.check(lib.fastpm_b200_store_get_column(g.cdm, 1 << 1, xbuf.ctypes.data, i, m), "x")
        _lib.check(lib.fastpm_b200_store_get_column(g.cdm, 1 << 8, ibuf.ctypes.data, i, m), "id")
        finite = finite and bool(np.isfinite(xbuf[:m]).all())
        chk += np.array(B.position_checksum(xbuf[:m], ibuf[:m], float(nc)))
    free_b, total_b = C.c_size_t(), C.c_size_t()
    lib.fpm_device_mem_info(C.byref(free_b), C.byref(total_b))
    t_ms, n_tot, fin, used_dev = ms.value, n_local, 1 if finite else 0, (total_b.value - free_b.value) / 1e9
    if world > 1:
        import torch
        v = torch.tensor([t_ms, used_dev], dtype=torch.float64, device="cuda")
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        t_ms, used_dev = float(v[0]), float(v[1])
        w = torch.tensor([n_tot, fin], dtype=torch.int64, device="cuda")
        dist.all_reduce(w[:1]); dist.all_reduce(w[1:], op=dist.ReduceOp.MIN)
        n_tot, fin = int(w[0]), int(w[1])
        c = torch.tensor(chk, dtype=torch.float64, device="cuda")
        dist.all_reduce(c)
        chk = c.cpu().numpy()
    if rank == 0:
        print(json.dumps({
            "config": {"nc": nc, "pm_nc_factor": pmf, "steps": K, "mode": args.mode, "gpus": world, "alloc_factor": args.alloc},
            "ms_per_step": t_ms / K, "particles_per_s": nc ** 3 * K / (t_ms * 1e-3), "ic_seconds": round(t_ic, 2),
            "np_expected": nc ** 3, "np_total_after": n_tot, "np_rank0_before_after": [np0, n_local], "result_finite": bool(fin),
            "mesh_per_step": meshes, "pk_bins_last": [float(v) for v in spectra[-1][1][:8]], "x_checksum": [float(v) for v in chk],
            "stages_rank0": stages, "paths": dict(zip(["fft_tma", "fft_tile_generic", "fft_zrow", "fft_z_generic", "fft_tma_multi", "paint_bricks",
                                                       "readout_bricks", "pk_rows", "staged_transpose", "paint_tiles", "readout_tiles"], used)),
            "device_gb_in_use_max": round(used_dev, 1)}))
    g.close()
    if world > 1:
        dist.barrier()


if __name__ == "__main__":
    main()
