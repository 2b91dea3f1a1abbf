"""ctypes loader for libfastpm_b200.so (the C ABI in include/fastpm_b200.h).

The product path is the CUDA library only: importing this module without the built
library, or calling into it without a CUDA device, raises -- there is no CPU fallback.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfastpm_b200.so")

# every symbol include/fastpm_b200.h declares; tests check that the library exports all of them
ABI_SYMBOLS = [
    "fpm_last_error", "fpm_version", "fpm_device_init", "fpm_device_count", "fpm_device_mem_info", "fpm_debug_state",
    "fpm_malloc", "fpm_free", "fpm_host_alloc_pinned", "fpm_host_free_pinned",
    "fpm_memcpy_h2d", "fpm_memcpy_d2h", "fpm_memcpy_h2d_async", "fpm_memcpy_d2h_async", "fpm_copy_fence", "fpm_copy_wait", "fpm_copy_fence_before_update", "fpm_memcpy_d2d", "fpm_memset", "fpm_sync",
    "fpm_timer_create", "fpm_timer_start", "fpm_timer_stop", "fpm_timer_elapsed_ms", "fpm_timer_destroy",
    "fpm_kernel_launch_count", "fpm_path_counts", "fpm_comm_byte_counts", "fpm_prof_enable", "fpm_prof_reset", "fpm_prof_get", "fpm_prof_get_launches",
    "fpm_mesh_create", "fpm_mesh_destroy", "fpm_mesh_info", "fpm_mesh_ktables_host",
    "fpm_paint", "fpm_readout", "fpm_readout3", "fpm_readout_pack3", "fpm_paint_window", "fpm_readout_window", "fpm_paint_window_ex", "fpm_readout_window_ex", "fpm_window_halo_planes", "fpm_particle_grid_hint", "fpm_tile_stats", "fpm_r2c", "fpm_r2c_ws", "fpm_c2r", "fpm_c2r_ws", "fpm_fft_set_generic", "fpm_transfer_for_kernel",
    "fpm_apply_transfer", "fpm_apply_decic", "fpm_decic_defer", "fpm_decic_cancel", "fpm_sync_deferred", "fpm_apply_pgd_transfer", "fpm_apply_radial", "fpm_remove_variance", "fpm_apply_axis_factors", "fpm_scale", "fpm_divide", "fpm_muladd", "fpm_set_mode",
    "fpm_induce_correlation", "fpm_fill_gaussian_gadget", "fpm_fill_whitenoise", "fpm_powerspectrum", "fpm_powerspectrum_sums", "fpm_cross_powerspectrum_sums",
    "fpm_kick", "fpm_drift", "fpm_update_fused", "fpm_pgd_shift", "fpm_wrap", "fpm_wrap_paint", "fpm_wrap_check", "fpm_shift_positions", "fpm_cast_f64_to_f32", "fpm_particle_grid_hint_get", "fpm_id_order_counts", "fpm_subsample_mask", "fpm_mask_scan", "fpm_compact_rows", "fpm_gather_rows", "fpm_permute_by_id", "fpm_fill_rand", "fpm_summary", "fpm_fill_grid", "fpm_lpt_evolve",
]


class FpmTransfer(C.Structure):
    _fields_ = [("active", C.c_int32), ("potorder", C.c_int32), ("negate", C.c_int32), ("ngrad", C.c_int32),
                ("graddir", C.c_int32 * 2), ("gradorder", C.c_int32), ("zero_selfconj", C.c_int32),
                ("scale", C.c_double)]


class FastPMB200Error(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (no device is touched yet)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FastPMB200Error(
            "libfastpm_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C fastpm_b200/csrc`; there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i64, dbl, i32, sz = C.c_void_p, C.c_int64, C.c_double, C.c_int, C.c_size_t
    lib.fpm_last_error.restype = C.c_char_p
    lib.fpm_version.restype = C.c_char_p
    lib.fpm_malloc.restype = vp
    lib.fpm_malloc.argtypes = [sz]
    lib.fpm_free.argtypes = [vp]
    lib.fpm_host_alloc_pinned.restype = vp
    lib.fpm_host_alloc_pinned.argtypes = [sz]
    lib.fpm_host_free_pinned.argtypes = [vp]
    lib.fpm_memcpy_h2d.argtypes = [vp, vp, sz]
    lib.fpm_memcpy_d2h.argtypes = [vp, vp, sz]
    lib.fpm_memcpy_h2d_async.argtypes = [vp, vp, sz]
    lib.fpm_fill_rand.argtypes = [vp, i64, i32]
    lib.fpm_cast_f64_to_f32.argtypes = [vp, vp, i64]
    lib.fpm_id_order_counts.argtypes = [vp, i64, C.c_uint64, vp]
    lib.fpm_permute_by_id.argtypes = [vp, vp, vp, i64, C.c_uint64, i32]
    lib.fpm_subsample_mask.argtypes = [vp, vp, dbl, i64, vp]
    lib.fpm_mask_scan.argtypes = [vp, i64, vp, vp]
    lib.fpm_compact_rows.argtypes = [vp, vp, vp, vp, i64, i32]
    lib.fpm_gather_rows.argtypes = [vp, vp, vp, i64, i32]
    lib.fpm_shift_positions.argtypes = [vp, i64, dbl, dbl, dbl]
    lib.fpm_memcpy_d2h_async.argtypes = [vp, vp, sz]
    lib.fpm_memcpy_d2d.argtypes = [vp, vp, sz]
    lib.fpm_memset.argtypes = [vp, i32, sz]
    lib.fpm_device_mem_info.argtypes = [C.POINTER(sz), C.POINTER(sz)]
    lib.fpm_timer_create.argtypes = [C.POINTER(vp)]
    lib.fpm_timer_start.argtypes = [vp]
    lib.fpm_timer_stop.argtypes = [vp]
    lib.fpm_timer_elapsed_ms.argtypes = [vp, C.POINTER(dbl)]
    lib.fpm_timer_destroy.argtypes = [vp]
    lib.fpm_kernel_launch_count.restype = C.c_uint64
    lib.fpm_prof_get.argtypes = [vp, vp, i32]
    lib.fpm_prof_get_launches.argtypes = [vp, vp, i32]
    lib.fpm_mesh_create.restype = vp
    lib.fpm_mesh_create.argtypes = [i32, dbl, i32, i32]
    lib.fpm_mesh_destroy.argtypes = [vp]
    lib.fpm_mesh_info.argtypes = [vp, vp]
    lib.fpm_mesh_ktables_host.argtypes = [vp, vp]
    lib.fpm_paint.argtypes = [vp, vp, vp, i64, dbl, vp, vp, i32]
    lib.fpm_readout.argtypes = [vp, vp, vp, i64, vp, i32, dbl]
    lib.fpm_readout3.argtypes = [vp, vp, vp, vp, vp, i64, vp]
    lib.fpm_paint_window.argtypes = [vp, i32, i32, vp, vp, i64, dbl, vp, vp, i32]
    lib.fpm_readout_window.argtypes = [vp, i32, i32, vp, vp, i64, vp, i32]
    lib.fpm_paint_window_ex.argtypes = [vp, i32, i32, i32, vp, vp, vp, i64, dbl, vp, vp, i32]
    lib.fpm_readout_window_ex.argtypes = [vp, i32, i32, i32, vp, vp, vp, i64, vp, i32]
    lib.fpm_window_halo_planes.argtypes = [i32, i32, vp, vp]
    lib.fpm_r2c.argtypes = [vp, vp, vp, dbl]
    lib.fpm_c2r.argtypes = [vp, vp, vp, C.POINTER(FpmTransfer)]
    lib.fpm_transfer_for_kernel.argtypes = [i32, i32, i32, C.POINTER(FpmTransfer)]
    lib.fpm_apply_transfer.argtypes = [vp, vp, vp, C.POINTER(FpmTransfer)]
    lib.fpm_apply_decic.argtypes = [vp, vp, vp]
    lib.fpm_decic_defer.argtypes = [vp, vp]
    lib.fpm_decic_cancel.argtypes = [vp]
    lib.fpm_scale.argtypes = [vp, vp, sz, dbl]
    lib.fpm_apply_pgd_transfer.argtypes = [vp, vp, vp, dbl, dbl, dbl]
    lib.fpm_apply_radial.argtypes = [vp, vp, vp, i32, dbl]
    lib.fpm_remove_variance.argtypes = [vp, vp]
    lib.fpm_apply_axis_factors.argtypes = [vp, vp, vp, vp]
    lib.fpm_divide.argtypes = [vp, vp, sz, dbl]
    lib.fpm_r2c_ws.argtypes = [vp, vp, vp, vp, dbl]
    lib.fpm_c2r_ws.argtypes = [vp, vp, vp, vp, C.POINTER(FpmTransfer)]
    lib.fpm_muladd.argtypes = [vp, vp, vp, sz, i32]
    lib.fpm_set_mode.argtypes = [vp, vp, i32, i32, i32, C.c_float, C.c_float]
    lib.fpm_induce_correlation.argtypes = [vp, vp, vp, vp, i32]
    lib.fpm_fill_whitenoise.argtypes = [vp, vp, C.c_uint64]
    lib.fpm_fill_gaussian_gadget.argtypes = [vp, vp, i32]
    lib.fpm_powerspectrum.argtypes = [vp, vp, i32, vp, vp, vp]
    lib.fpm_powerspectrum_sums.argtypes = [vp, vp, i32, vp]
    lib.fpm_cross_powerspectrum_sums.argtypes = [vp, vp, vp, vp]
    lib.fpm_kick.argtypes = [vp, vp, vp, vp, vp, i64, i32, dbl, dbl, dbl, dbl, dbl]
    lib.fpm_drift.argtypes = [vp, vp, vp, vp, vp, i64, i32, dbl, dbl, dbl, dbl, dbl]
    lib.fpm_update_fused.argtypes = [vp, vp, vp, vp, vp, i64, i32, vp]
    lib.fpm_pgd_shift.argtypes = [vp, vp, i64, dbl, dbl]
    lib.fpm_wrap.argtypes = [vp, i64, dbl]
    lib.fpm_wrap_paint.argtypes = [vp, vp, vp, i64, dbl, vp, vp, i32]
    lib.fpm_summary.argtypes = [vp, i32, i32, i64, vp]
    lib.fpm_fill_grid.argtypes = [vp, vp, vp, i32, i32, i64, dbl, dbl]
    lib.fpm_lpt_evolve.argtypes = [vp, vp, vp, vp, i64, dbl, dbl, dbl, dbl]
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        raise FastPMB200Error("%s failed: %s" % (what or "fastpm_b200 call", load().fpm_last_error().decode()))


_device = None


def require_device(device=None):
    """Initialise the CUDA device; raises if there is none (the product has no CPU path).

    device=None keeps the device this process already initialised, else FASTPM_B200_DEVICE / LOCAL_RANK / 0
    (the same rule as libfastpm_init in csrc/host/support.c)."""
    global _device
    lib = load()
    if device is None:
        if _device is not None:
            return lib
        device = int(os.environ.get("FASTPM_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    if _device != int(device):
        check(lib.fpm_device_init(int(device)), "fpm_device_init")
        _device = int(device)
    return lib
