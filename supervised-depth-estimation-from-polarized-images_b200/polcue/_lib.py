"""ctypes binding of libpolcue.so (the C ABI declared in include/polcue.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C csrc``.  There is no CPU or
pure-PyTorch fallback: if the shared object is missing, importing an op raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("POLCUE_LIB") or os.path.join(_HERE, "libpolcue.so")   # POLCUE_LIB: tuning builds only

OK, EINVAL, ENOMEM, ERANGE, E2BIG = 0, -22, -12, -34, -7


class PolcueError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        super().__init__(f"{where} failed with code {code}: {error_string(code)}")


_lib = None

_u8p, _f32p, _f64p, _vp = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p   # raw device/host addresses
_SIGNATURES = {
    # name: (restype, argtypes)
    "polcue_version": (C.c_char_p, []),
    "polcue_error_string": (C.c_char_p, [C.c_int]),
    "polcue_launch_count": (C.c_ulonglong, []),
    "polcue_debug_stencil_tma_launches": (C.c_ulonglong, []),
    "polcue_lut_create": (C.c_int, [C.c_double, C.POINTER(C.c_void_p)]),
    "polcue_lut_destroy": (None, [C.c_void_p]),
    "polcue_lut_set_trig": (C.c_int, [C.c_void_p, C.c_int]),
    "polcue_lut_host_build": (C.c_int, [C.c_double, C.POINTER(C.c_void_p)]),
    "polcue_lut_cells": (C.c_int, [C.c_void_p, C.c_int]),
    "polcue_lut_steep": (C.c_int, [C.c_void_p, C.c_int, _f64p]),
    "polcue_lut_knots": (C.c_int, [C.c_void_p, C.c_int, _f64p, _f64p, C.c_int]),
    "polcue_lut_eval_host": (C.c_int, [C.c_void_p, C.c_int, _f32p, C.c_size_t, _f32p]),
    "polcue_split_pol": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "polcue_fused_mosaic_u8": (C.c_int, [_u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, _u8p, _f32p, _f32p, _f32p, _vp]),
    "polcue_fused_stats_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "polcue_fused_mosaic_stats_u8": (C.c_int, [_u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, _u8p, _f32p, _f32p, _f32p, _vp, _f64p, _vp]),
    "polcue_fused_superpixel_u8": (C.c_int, [_u8p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_void_p, _u8p, _f32p, _f32p, _f32p, _vp]),
    "polcue_fused_planes_u8": (C.c_int, [_u8p, _u8p, _u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, _f32p, _f32p, _f32p, _vp]),
    "polcue_resize_plan_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "polcue_resize_plan_host_build": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "polcue_resize_plan_destroy": (None, [C.c_void_p]),
    "polcue_resize_plan_coeffs": (C.c_int, [C.c_void_p, C.c_int, _vp, _vp, C.c_size_t]),
    "polcue_debug_resize_force_bytes": (C.c_int, [C.c_int]),
    "polcue_debug_resize_pass_times": (C.c_int, [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "polcue_resize_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int]),
    "polcue_resize_lanczos_u8": (C.c_int, [C.c_void_p, _u8p, C.c_int, _u8p, _u8p, _u8p, _vp]),
    "polcue_loader_front_end_u8": (C.c_int, [C.c_void_p, _u8p, _u8p, _u8p, _u8p, C.c_int, _u8p, C.c_void_p, _u8p, _u8p, _f32p, _f32p,
                                            _f32p, C.POINTER(C.c_float), _f32p, _vp]),
    "polcue_loader_front_end_u8_host": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _u8p, _u8p, _u8p, _u8p, C.c_int, _u8p, C.c_void_p, _u8p,
                                                 _f32p, _f32p, C.POINTER(C.c_float), _f32p, C.c_int]),
    "polcue_fused_mosaic_u8_host": (C.c_int, [_u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, _f32p, _f32p, _f32p, C.c_int]),
    "polcue_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "polcue_host_alloc_on": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t, C.c_int]),
    "polcue_host_numa_node": (C.c_int, [C.c_int]),
    "polcue_host_free": (C.c_int, [C.c_void_p]),
    "polcue_xolp_stack_u8": (C.c_int, [_u8p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), _f32p, _f32p, _vp]),
    "polcue_xolp_stack_f32": (C.c_int, [_f32p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), _f32p, _f32p, _vp]),
    "polcue_xolp_planes_u8": (C.c_int, [_u8p, _u8p, _u8p, _u8p, C.c_int, C.c_int, C.c_int, _f32p, _f32p, _vp]),
    "polcue_normals_from_xolp_f32": (C.c_int, [_f32p, C.c_int, C.c_int, C.c_int, C.c_void_p, _f32p, _vp]),
    "polcue_rho_diffuse_f32": (C.c_int, [_f32p, C.c_size_t, C.c_void_p, _f32p, _vp]),
    "polcue_rho_spec_f32": (C.c_int, [_f32p, C.c_size_t, C.c_void_p, _f32p, _f32p, _vp]),
    "polcue_calc_normals_f32": (C.c_int, [_f32p, _f32p, C.c_int, C.c_size_t, _f32p, _vp]),
    "polcue_stokes_channel_f32": (C.c_int, [_f32p, _u8p, C.c_int, C.c_int, _f32p, _f32p, _f32p, _vp]),
    "polcue_calc_normals_channel_f32": (C.c_int, [_f32p, _f32p, _u8p, C.c_size_t, C.c_float, _f32p, _vp]),
    "polcue_depth_to_normals_f32": (C.c_int, [_f32p, _f32p, C.c_int, C.c_int, C.c_int, _f32p, _vp]),
    "polcue_normals_loss_workspace_bytes": (C.c_size_t, []),
    "polcue_normals_loss_fwd_f32": (C.c_int, [_f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, _vp, _f64p, _f32p, _vp]),
    "polcue_normals_loss_bwd_f32": (C.c_int, [_f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, _f64p, _f32p, _f32p, _vp]),
    "polcue_supervised_losses_fwd_f32": (C.c_int, [_f32p, _f32p, _f32p, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, _vp, _f64p, _f32p, _vp]),
    "polcue_supervised_losses_bwd_f32": (C.c_int, [_f32p, _f32p, _f32p, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, _f64p, _f32p, _f32p,
                                                  _f32p, _vp]),
    "polcue_depth_errors_workspace_bytes": (C.c_size_t, []),
    "polcue_depth_errors_f32": (C.c_int, [_f32p, _f32p, C.c_size_t, _vp, _f64p, _f32p, _vp]),
    "polcue_depth_errors_images_f32": (C.c_int, [_f32p, _f32p, _u8p, C.c_int, C.c_size_t, C.c_float, C.c_float, C.c_int,
                                                  _f64p, _f32p, _vp]),
    "polcue_masked_median_scale_f32": (C.c_int, [_f32p, _f32p, _u8p, C.c_int, C.c_size_t, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int,
                                                _f32p, _f32p, _vp]),
    "polcue_depth_errors_images_scaled_f32": (C.c_int, [_f32p, _f32p, _u8p, C.c_int, C.c_size_t, C.c_float, C.c_float, C.c_int, C.c_int,
                                                       C.c_int, _f32p, _f64p, _f32p, _vp]),
    "polcue_depth_errors_groups_f32": (C.c_int, [_f32p, _f32p, _u8p, C.c_int, C.c_size_t, C.c_float, C.c_float, C.POINTER(C.c_int),
                                                  C.c_int, _f64p, _f32p, _vp]),
    "polcue_eval_pass_f32": (C.c_int, [_f32p, _f32p, _u8p, _f32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.POINTER(C.c_int),
                                      C.c_int, _f32p, _f64p, _f32p, _f64p, _vp]),
    "polcue_eval_pass_peer_f32": (C.c_int, [_f32p, _f32p, _u8p, _f32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.POINTER(C.c_int),
                                           C.c_int, _f32p, _f64p, _f32p, _f64p, _vp, _f64p, _vp]),
    "polcue_peer_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(_vp), _vp]),
    "polcue_peer_connect": (C.c_int, [_vp, _vp]),
    "polcue_peer_allreduce_f64": (C.c_int, [_vp, _f64p, C.c_int, _f64p, _vp]),
    "polcue_peer_status": (C.c_int, [_vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]),
    "polcue_peer_destroy": (C.c_int, [_vp]),
    "polcue_channel_stats_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_size_t]),
    "polcue_channel_stats_f32": (C.c_int, [_f32p, C.c_int, C.c_int, C.c_size_t, _vp, _f64p, _vp]),
}
EXPORTS = tuple(_SIGNATURES)


def lib():
    """The loaded library; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C supervised-depth-estimation-from-polarized-images_b200/csrc`. polcue has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)       # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def error_string(code):
    return lib().polcue_error_string(int(code)).decode()


def check(code, where):
    if code != OK:
        raise PolcueError(code, where)


def launch_count():
    return int(lib().polcue_launch_count())
