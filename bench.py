#!/usr/bin/env python
"""bench.py -- PM-step throughput of the B200 path, with roofline, CPU baseline and end-to-end numbers.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--nc NC]

Workload (BASELINE.json): nc^3 particles on a (2 nc)^3 force mesh, COLA, K PM steps
(time_step = linspace(0.1, 1, K): K force evaluations and K-1 kick-drift-kick cycles, the reference's own
meaning of "K steps", tests/standard.lua:19-22), synthetic 2LPT initial conditions at a = 0.1 (the reference's Gadget-scheme
white noise for seed 100 coloured by the reference's linear P(k) table, generated on the device), P(k) handler on (one measurement + device->host read per force evaluation).

metric  particles/s = nc^3 * K / t_evolve   (src/fastpm.c:368-370 times exactly fastpm_solver_evolve)
value   t_evolve measured with CUDA events around fastpm_solver_evolve, particle state resident in HBM
e2e     the same call with HOST particle buffers: pinned host -> device copy of x, v, dx1, dx2, id, the evolve,
        and the device -> host copy of x, v, id, all inside the timed region
The reference arm (--impl reference) and the cpu_baseline object time the reference's own sources
(oracle/_ref, built from /root/reference against the shims in oracle/shims) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pm_step_particles_per_second"
UNIT = "particles/s"
KCLASSES = ["paint", "readout", "fft_tile", "fft_z", "kick", "drift", "kspace", "pk", "summary", "other",
            "memset", "barrier", "halo", "migrate", "push"]


def read_pk():
    import numpy as np
    tab = np.loadtxt(os.path.join(ROOT, "tests", "golden", "powerspec.txt"))
    return tab[:, 0].copy(), tab[:, 1].copy()


def measured_traffic(nmesh):
    """DRAM bytes (read + written) per launch of the dominant kernel -- the strided FFT tile pass -- at this mesh size, averaged over the
    launches of the committed ncu capture (profiles/r02_dram_traffic_n2048.csv, written by the round-2 summary script from the
    .ncu-rep files: dram__bytes_read.sum + dram__bytes_write.sum per launch), or None when no capture at this size is committed."""
    import csv
    try:
        tot, n = 0.0, 0
        with open(os.path.join(ROOT, "profiles", "r02_dram_traffic_n2048.csv")) as f:
            for row in csv.DictReader(f):
                if row["kernel"].startswith("fft_tma_kernel") and int(row["nmesh"]) == int(nmesh):
                    tot += float(row["dram_read_bytes"]) + float(row["dram_write_bytes"])
                    n += 1
        return tot / n if n else None
    except Exception:
        return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [v.strip() for v in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_run(nc, B, mode, K, W, threads):
    """The reference's own CPU implementation (oracle/_ref) on a bounded sample of the workload."""
    os.environ["OMP_NUM_THREADS"] = str(threads)
    import numpy as np
    from oracle import ref
    if not ref.available():
        return None
    pk_text = open(os.path.join(ROOT, "tests", "golden", "powerspec.txt")).read()
    s = ref.Session(nc=nc, boxsize=float(nc), pm_nc_factor=B, force_mode=mode, growth_mode="LCDM", np_alloc_factor=2.0)
    dk, _, _ = s.ic_deltak(100, pk_text)
    s.setup_lpt(dk, 0.1)
    p0 = s.get_particles()
    ts = np.linspace(0.1, 1.0, K)
    if W >= 2:
        s.evolve(ts[:W])
        s.set_particles(p0["x"], v=p0["v"], id=p0["id"], dx1=p0.get("dx1"), dx2=p0.get("dx2"), meta=p0["meta"])
    c0 = ref.clock_table()
    t = s.evolve(ts)
    c1 = ref.clock_table()
    p1 = s.get_particles()
    recs = s.records()
    s.close()
    # the reference's own named clocks over the timed evolve (fastpm_clock_stat, prof.c:144)
    phases = {}
    for key, val in c1.items():
        d = val - c0.get(key, 0.0)
        if d > 0:
            phases[key.split(":")[1]] = round(phases.get(key.split(":")[1], 0.0) + d, 2)
    return dict(value=nc ** 3 * K / t, seconds=t, np=nc ** 3, phases=phases,
                pk_bins=[float(v) for v in recs[-1]["p"][:8]] if recs else None,
                x_checksum=position_checksum(p1["x"], p1["id"], float(nc)))


def position_checksum(x, ids, L):
    """Order-independent fingerprint of a particle set: sum_i w(id_i) * (x_i mod L) per component with w in (0, 1] a fixed hash
    of the id.  Any decomposition (1, 2, 4, 8 slabs) of the same run gives the same three numbers up to float summation order."""
    import numpy as np
    tot = np.zeros(3)
    n = len(ids)
    step = 1 << 24
    xs = np.asarray(x).reshape(-1, 3)
    for i in range(0, n, step):
        w = ((np.asarray(ids[i:i + step], dtype=np.uint64) * np.uint64(2654435761)) % np.uint64(1 << 20)).astype(np.float64)
        w = (w + 1.0) / float(1 << 20)
        tot += (np.mod(xs[i:i + step], L) * w[:, None]).sum(axis=0)
    return [float(v) for v in tot]


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    nc_s = args.ref_nc if args.ref_nc > 0 else 256     # C1 (nc = 256, 512^3 mesh) is the CPU-runnable case
    r = cpu_reference_run(nc_s, args.pm_nc_factor, args.mode, args.steps, args.warmup, threads)
    if r is None:
        emit({"impl": "reference", "unavailable": "oracle/_ref/libfastpm_ref.so is not built"})
        return 0
    sample = "nc=%d^3 particles, %d^3 mesh, %s, %d steps: the reference's own sources (oracle/_ref, shimmed FFT/GSL/MPI), 1 rank x %d threads; a bounded sample of the nc=%d workload of the GPU arm" % (
        nc_s, nc_s * args.pm_nc_factor, args.mode, args.steps, threads, args.nc)
    ran = argparse.Namespace(**vars(args))
    ran.nc = nc_s                                       # `config` describes what this arm RAN, not what the GPU arm runs
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 mesh / f64 positions", "data": "synthetic",
        "config": dict(workload_config(ran), gpu_arm_nc=args.nc, note="CPU arm: bounded sample at nc=%d, the GPU arm runs nc=%d" % (nc_s, args.nc)),
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample,
                         "phase_seconds": r["phases"]},
        "pk_bins": r["pk_bins"], "x_checksum": r["x_checksum"],
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


def default_nc(args):
    """nc = 1024 (the configuration BASELINE.json's metric is quoted on) needs 156 GB of HBM on one GPU (2 meshes of 34.9 GB +
    80 B per particle) and, for the end-to-end leg, 73 GB of pinned host memory; otherwise the largest power-of-two case that fits."""
    if args.impl == "reference":
        return 1024
    world = max(1, int(os.environ.get("WORLD_SIZE", "1")), args.gpus)
    try:
        with open("/proc/meminfo") as f:
            avail_kb = [int(l.split()[1]) for l in f if l.startswith("MemAvailable")][0]
    except Exception:
        avail_kb = 0
    try:
        import ctypes as C
        from fastpm_b200 import _lib
        lib = _lib.require_device(int(os.environ.get("LOCAL_RANK", "0")))
        free_b, total_b = C.c_size_t(), C.c_size_t()
        lib.fpm_device_mem_info(C.byref(free_b), C.byref(total_b))
        dev_free = free_b.value
    except Exception:
        dev_free = 0
    need_dev = (2 * 34.9e9 + 80.0 * 1024 ** 3 * (1.0 if world == 1 else 1.3)) / world + 4e9
    need_host = 73e9 / world + 16e9
    return 1024 if (dev_free >= need_dev and avail_kb * 1024.0 >= need_host) else 512


def workload_config(args):
    return {"workload": "nc=%d^3 particles, %d^3 mesh (B=%d), %s, %d PM steps linspace(0.1,1,%d), 2LPT ICs, P(k) each step" % (
        args.nc, args.nc * args.pm_nc_factor, args.pm_nc_factor, args.mode.upper(), args.steps, args.steps),
        "nc": args.nc, "nmesh": args.nc * args.pm_nc_factor, "boxsize_mpc_h": float(args.nc), "force_mode": args.mode,
        "l2": "mesh (%.1f GB) and particle columns exceed the 126 MB L2; no flush needed" % (
            4.0 * (args.nc * args.pm_nc_factor) ** 2 * (args.nc * args.pm_nc_factor + 2) / 1e9)}


IC_SEED = 100


def setup_ic(g, k_tab, p_tab, a0):
    """The reference's own initial conditions for seed 100 (Gadget-scheme RANLUX white noise coloured by the P(k) table, 2LPT at a0:
    src/fastpm.c:415-545), generated on the device -- the same particles the reference arm starts from at equal nc.
    FASTPM_B200_BENCH_IC=philox selects the counter-based white noise of round 1 instead (not comparable with the reference)."""
    if os.environ.get("FASTPM_B200_BENCH_IC", "gadget") == "philox":
        g.setup_synthetic_ic(IC_SEED, k_tab, p_tab, a0)
        return "philox"
    g.setup_ic(IC_SEED, k_tab, p_tab, a0)
    return "gadget"


def run_ours(args):
    import ctypes as C
    import numpy as np
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 or world > 1:
        from fastpm_b200 import multigpu
        return multigpu.bench_main(args)
    from fastpm_b200 import _lib
    from fastpm_b200.solver import Solver, ForceEvent
    lib = _lib.require_device(int(os.environ.get("LOCAL_RANK", "0")))
    nc, B, K, W = args.nc, args.pm_nc_factor, args.steps, args.warmup
    N = nc * B
    Np = nc ** 3
    k_tab, p_tab = read_pk()
    ts = np.linspace(0.1, 1.0, K)

    g = Solver(nc=nc, boxsize=float(nc), pm_nc_factor=B, force_mode=args.mode, growth_mode="LCDM", np_alloc_factor=1.0)
    setup_ic(g, k_tab, p_tab, ts[0])
    meta0 = g.meta
    cola = args.mode == "cola"
    cols_in = ["x", "v", "id"] + (["dx1", "dx2"] if cola else [])
    cols_out = ["x", "v", "id"]
    itemsize = dict(x=24, v=12, id=8, dx1=12, dx2=12)
    dtypes = dict(x=np.float64, v=np.float32, id=np.uint64, dx1=np.float32, dx2=np.float32)
    host = {}
    for c in cols_in:                                   # pinned host copies of the initial state
        ptr = lib.fpm_host_alloc_pinned(Np * itemsize[c])
        if not ptr:
            raise RuntimeError("pinned host allocation failed: " + lib.fpm_last_error().decode())
        n_el = Np * itemsize[c] // np.dtype(dtypes[c]).itemsize
        host[c] = (ptr, np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtypes[c]))), (n_el,)))
        _lib.check(lib.fpm_memcpy_d2h(ptr, g.column_ptr(c), Np * itemsize[c]), "save IC")

    def restore():
        for c in cols_in:
            _lib.check(lib.fpm_memcpy_h2d(g.column_ptr(c), host[c][0], Np * itemsize[c]), "restore IC")
        g.set_meta(meta0["a_x"], meta0["a_v"], meta0["M0"])

    spectra = []
    # state of the end-to-end arm (below): which force evaluation is running, whether the uploads have been fenced
    e2e = dict(on=False, nforce=0, nforce_total=-1)

    def on_force_after(solver_ptr, event_ptr, userdata):      # the reference's write_powerspectrum handler, src/fastpm.c:1711-1776
        ev = C.cast(event_ptr, C.POINTER(ForceEvent)).contents
        spectra.append(g.powerspectrum_of(ev.pm, ev.delta_k))
        e2e["nforce"] += 1
        if e2e["on"] and e2e["nforce"] == e2e["nforce_total"]:
            # the deposit of the LAST force evaluation is done: positions (wrapped) and ids are final, they go down on the copy
            # stream while the inverse transforms and the gather of this evaluation run
            for c in ("x", "id"):
                _lib.check(lib.fpm_memcpy_d2h_async(host[c][0], g.column_ptr(c), Np * itemsize[c]), "result d2h (async)")
        return 0

    g.add_handler("FORCE", 1, on_force_after)

    if W >= 1:                                          # warm-up: the first max(W,2) entries of the same table
        g.evolve(ts[:max(W, 2)])
        restore()

    # ---- device-resident timed run
    timer = C.c_void_p()
    _lib.check(lib.fpm_timer_create(C.byref(timer)))
    lib.fpm_prof_reset()
    lib.fpm_prof_enable(1)
    launches0 = int(lib.fpm_kernel_launch_count())
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    _lib.check(lib.fpm_sync())
    t_wall = time.perf_counter()
    lib.fpm_timer_start(timer)
    e2e["nforce"] = 0
    g.evolve(ts)
    e2e["nforce_total"] = e2e["nforce"]
    lib.fpm_timer_stop(timer)
    ms = C.c_double()
    _lib.check(lib.fpm_timer_elapsed_ms(timer, C.byref(ms)))
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop()
    launches = int(lib.fpm_kernel_launch_count()) - launches0
    lib.fpm_prof_enable(0)
    counts = (C.c_int64 * len(KCLASSES))()
    totals = (C.c_double * len(KCLASSES))()
    _lib.check(lib.fpm_prof_get(counts, totals, len(KCLASSES)))
    stages = {nm: {"launches": int(counts[i]), "ms": round(float(totals[i]), 3)} for i, nm in enumerate(KCLASSES)}
    # launch by launch (issue order) for the last full step: which pass of which transform costs what
    ncap = 4096
    lcls = (C.c_int32 * ncap)()
    lms = (C.c_double * ncap)()
    nl = min(int(lib.fpm_prof_get_launches(lcls, lms, ncap)), ncap)
    per_step = max(1, nl // K)
    trace = [[KCLASSES[lcls[i]], round(lms[i], 3)] for i in range(max(0, nl - per_step), nl)]
    t_evolve = ms.value / 1e3
    value = Np * K / t_evolve

    # ---- end-to-end run: host buffers in, host buffers out.  The PCIe traffic overlaps the first and the last force evaluation
    # (FASTPM_B200_BENCH_E2E_OVERLAP=0: everything up, run, everything down): x goes up first, the other columns follow on the
    # copy stream while the first force evaluation -- which reads only x -- runs (fpm_copy_fence_before_update); x and id go down during
    # the inverse transforms and the gather of the last one (on_force_after); v follows when the run is over.
    overlap = os.environ.get("FASTPM_B200_BENCH_E2E_OVERLAP", "1") != "0"
    _lib.check(lib.fpm_sync())
    t0 = time.perf_counter()
    if overlap:
        e2e.update(on=True, nforce=0)
        _lib.check(lib.fpm_memcpy_h2d(g.column_ptr("x"), host["x"][0], Np * itemsize["x"]), "restore IC")
        for c in cols_in:
            if c != "x":
                _lib.check(lib.fpm_memcpy_h2d_async(g.column_ptr(c), host[c][0], Np * itemsize[c]), "restore IC (async)")
        _lib.check(lib.fpm_copy_fence_before_update(), "fence")      # the first kick waits (on the device) for these uploads
        g.set_meta(meta0["a_x"], meta0["a_v"], meta0["M0"])
        g.evolve(ts)
        assert e2e["nforce"] == e2e["nforce_total"]
        e2e["on"] = False
        _lib.check(lib.fpm_memcpy_d2h(host["v"][0], g.column_ptr("v"), Np * itemsize["v"]), "result d2h")
        _lib.check(lib.fpm_copy_wait(), "copy wait")
    else:
        restore()
        g.evolve(ts)
        for c in cols_out:
            _lib.check(lib.fpm_memcpy_d2h(host[c][0], g.column_ptr(c), Np * itemsize[c]), "result d2h")
    _lib.check(lib.fpm_sync())
    t_e2e = time.perf_counter() - t0
    h2d = sum(Np * itemsize[c] for c in cols_in)
    d2h = sum(Np * itemsize[c] for c in cols_out) + K * 3 * (N // 2) * 8
    x_final = host["x"][1]
    finite = bool(np.isfinite(x_final[:: max(1, x_final.size // 100000)]).all())

    # ---- roofline of the dominant kernel (the strided FFT tile pass: 4 of the 6 passes of every transform)
    S = 4.0 * N * N * (N + 2)                           # bytes of one padded real mesh
    peak, peak_src = measured_peak()
    tile = stages["fft_tile"]
    avg_ms = tile["ms"] / max(1, tile["launches"])
    achieved = 2 * S / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    fftz = stages["fft_z"]
    n_transforms = max(1, fftz["launches"])
    t_transform_ms = (tile["ms"] + fftz["ms"]) / n_transforms
    fft_gbs = 6 * S / (t_transform_ms * 1e-3) / 1e9 if t_transform_ms > 0 else 0.0

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": 1e3 * t_evolve / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 mesh / f64 positions", "data": "synthetic", "config": workload_config(args),
        "e2e": {"value": Np * K / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d // K, "d2h_bytes_per_step": d2h // K,
                "seconds": round(t_e2e, 4),
                "copies": "x up, then v / id / dx1 / dx2 on a copy stream during the first force evaluation; x and id down during the last one, v after it" if overlap else "all columns up, run, all columns down"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "fft_tile_kernel (strided y/x FFT pass)", "achieved": round(achieved, 1), "peak": peak,
                     "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": measured_traffic(N), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": 2 * S, "avg_launch_ms": round(avg_ms, 4)},
        "fft": {"gbs_6S": round(fft_gbs, 1), "frac_of_peak": round(fft_gbs / peak, 4), "ms_per_transform": round(t_transform_ms, 4),
                "transforms": n_transforms},
        "stages": stages, "last_step_launches": trace,
        "evolve_wall_s": round(t_wall, 4), "result_finite": finite,
        "pk_bins": [float(v) for v in spectra[-1][1][:8]] if spectra else None,
        "x_checksum": position_checksum(host["x"][1], host["id"][1], float(nc)),
        "np_total_after": int(g.np),
    }
    if os.environ.get("FASTPM_B200_TILE_STATS"):
        ts4 = (C.c_uint64 * 4)()
        lib.fpm_tile_stats.argtypes = [C.c_void_p]
        lib.fpm_tile_stats(ts4)
        line["tile_stats"] = {"paint_particles_on_global_path": int(ts4[0]), "paint_ctas_without_tile": int(ts4[1]),
                              "readout_particles_on_global_path": int(ts4[2]), "readout_ctas_without_tile": int(ts4[3]),
                              "note": "summed over every deposit / gather since the library was loaded (warm-up, timed and end-to-end runs)"}
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        ref_nc = args.ref_nc if args.ref_nc > 0 else 128
        r = cpu_reference_run(ref_nc, B, args.mode, min(K, args.ref_steps), 0, threads)
        if r is not None:
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "reference",
                                    "sample": "reference sources (oracle/_ref, shimmed FFT/GSL/MPI), 1 rank x %d threads: nc=%d^3, %d^3 mesh, %s, %d steps, %.1f s" % (
                                        threads, ref_nc, ref_nc * B, args.mode, min(K, args.ref_steps), r["seconds"]),
                                    "phase_seconds": r["phases"]}
        else:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": threads, "kind": "reference", "sample": "oracle/_ref not built"}
    if rank == 0:
        emit(line)
    g.close()
    return 0


_REAL_STDOUT = None


def emit(line):
    """Prints the one JSON line on the process's original stdout (fd 1 itself is redirected to stderr in main(), so that
    banners printed by native libraries -- e.g. NCCL's version line -- cannot end up next to it)."""
    data = (json.dumps(line) + "\n").encode()
    fd = os.environ.get("FASTPM_B200_BENCH_STDOUT_FD")       # set by main(); also seen by `import bench` from fastpm_b200.multigpu
    if fd is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(int(fd), data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.environ["FASTPM_B200_BENCH_STDOUT_FD"] = str(_REAL_STDOUT)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nc", type=int, default=int(os.environ.get("FASTPM_B200_BENCH_NC", "0")),
                    help="particles per side; 0 = BASELINE.json's nc=1024 (2048^3 mesh) when device and host memory allow, else 512")
    ap.add_argument("--pm-nc-factor", type=int, default=2)
    ap.add_argument("--mode", default="cola", choices=["cola", "pm", "fastpm"])
    ap.add_argument("--ref-nc", type=int, default=0, help="particle grid of the bounded CPU sample; 0 = 256 (BASELINE C1) for --impl reference, 128 for the cpu_baseline object of the GPU arm")
    ap.add_argument("--ref-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.steps < 2:
        args.steps = 2
    if args.nc <= 0:
        args.nc = default_nc(args)
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
