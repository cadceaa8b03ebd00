"""ctypes binding of libmulut_b200.so (the C ABI in include/mulut.h).

There is no CPU fallback: if the shared library is missing, loading fails
loudly and every product entry point raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MULUT_B200_LIB selects an instrumented side build (tools/bn_timing.py); the product never sets it
LIB_PATH = os.environ.get("MULUT_B200_LIB") or os.path.join(_HERE, "libmulut_b200.so")

OK, E_BAD_MODE, E_BAD_ARG, E_CUDA, E_NOMEM, E_LUT_SMALL = 0, -1, -2, -3, -4, -5
KERNEL_AUTO, KERNEL_GENERIC, KERNEL_TILED, KERNEL_TILED_QUAD, KERNEL_TILED_CELL, KERNEL_TILED_BINNED = -1, 0, 1, 1, 2, 3

PROF_KINDS = {"generic_stage": 0, "generic_last": 1, "smem_stage": 2, "combine": 3, "last_tiled": 4,
              "bin_hist": 5, "last_binned": 6, "bin_orphans": 7, "fused_stage": 8}

GB_VARIANTS = {
    "ldg_u8": 0, "ldg_u32": 1, "ldg_u128": 2, "quad_cell64": 3, "lds_u8": 4,
    "pair_cell64": 5, "oct_cell128": 6, "cpasync_cell64": 7, "lds_u32": 8,
    "quad_cell256_3rows": 9, "quad_cell256_4sect": 10, "bulk_cell256": 11, "bulk_rows64x3": 12,
}

# every symbol include/mulut.h declares: (restype, argtypes)
_c = ctypes
SYMBOLS = {
    "mulut_version": (_c.c_int, []),
    "mulut_last_error": (_c.c_char_p, []),
    "mulut_create": (_c.c_int, [_c.POINTER(_c.c_void_p), _c.c_int, _c.c_int, _c.c_char_p, _c.c_int, _c.c_int,
                                _c.POINTER(_c.c_void_p), _c.c_int]),
    "mulut_destroy": (_c.c_int, [_c.c_void_p]),
    "mulut_set_kernel": (_c.c_int, [_c.c_void_p, _c.c_int]),
    "mulut_reserve": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int]),
    "mulut_sr_infer_u8": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                     _c.c_void_p]),
    "mulut_sr_infer_u8_host": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int,
                                          _c.c_int]),
    "mulut_sr_infer_u8_host_async": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int,
                                          _c.c_int]),
    "mulut_host_copy_probe_async": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int,
                                          _c.c_int]),
    "mulut_sr_host_sync": (_c.c_int, [_c.c_void_p]),
    "mulut_launch_count": (_c.c_longlong, [_c.c_void_p]),
    "mulut_profile_enable": (_c.c_int, [_c.c_void_p, _c.c_int]),
    "mulut_profile_read": (_c.c_int, [_c.c_void_p, _c.c_int, _c.POINTER(_c.c_double), _c.POINTER(_c.c_longlong)]),
    "mulut_interp_pass_f64": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                         _c.c_int, _c.c_int, _c.c_char, _c.c_void_p, _c.c_void_p]),
    "mulut_interp_fwd_f32": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_char, _c.c_void_p, _c.c_int, _c.c_int,
                                        _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p]),
    "mulut_interp_bwd_f32": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_char, _c.c_void_p, _c.c_int, _c.c_int,
                                        _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p,
                                        _c.c_void_p, _c.c_void_p]),
    "mulut_stage_workspace_bytes": (_c.c_size_t, [_c.c_int, _c.c_int, _c.c_int]),
    "mulut_stage_fwd_f32": (_c.c_int, [_c.POINTER(_c.c_void_p), _c.c_int, _c.c_char_p, _c.c_int, _c.c_int, _c.c_int,
                                       _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_float, _c.c_float,
                                       _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    "mulut_stage_bwd_f32": (_c.c_int, [_c.POINTER(_c.c_void_p), _c.c_int, _c.c_char_p, _c.c_int, _c.c_int, _c.c_int,
                                       _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_float, _c.c_float,
                                       _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.POINTER(_c.c_void_p), _c.c_void_p,
                                       _c.c_void_p]),
    "mulut_adam_step_f32": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_void_p,
                                       _c.c_float, _c.c_float, _c.c_float, _c.c_float, _c.c_void_p, _c.c_void_p]),
    "mulut_mse_head_fwd_f32": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_float, _c.c_void_p, _c.c_void_p,
                                          _c.c_void_p]),
    "mulut_mse_head_bwd_f32": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_float, _c.c_void_p, _c.c_void_p,
                                          _c.c_void_p]),
    "mulut_eval_psnr_ssim_y_u8": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p,
                                             _c.POINTER(_c.c_double), _c.c_void_p]),
    "mulut_plan_bins": (_c.c_int, [_c.POINTER(_c.c_ulonglong), _c.c_longlong, _c.c_int, _c.c_ulonglong, _c.c_int,
                                   _c.POINTER(_c.c_int), _c.POINTER(_c.c_uint)]),
    "mulut_host_alloc": (_c.c_void_p, [_c.c_size_t]),
    "mulut_host_free": (_c.c_int, [_c.c_void_p]),
    "mulut_gather_bench": (_c.c_int, [_c.c_int, _c.c_int, _c.c_size_t, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                      _c.POINTER(_c.c_double)]),
}

_lib = None


class MulutError(RuntimeError):
    pass


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise MulutError(
                "libmulut_b200.so not found at {}: build it with `python -m mulut_b200.build` "
                "(there is no CPU fallback)".format(LIB_PATH))
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)        # AttributeError if the ABI drifted
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error() -> str:
    msg = lib().mulut_last_error()
    return msg.decode(errors="replace") if msg else ""


def check(rc: int) -> None:
    """Map C status codes onto the exception types the reference raises."""
    if rc == OK:
        return
    msg = last_error()
    if rc == E_BAD_MODE:
        raise ValueError(msg)                 # "Mode {} not implemented." (sr/4_test_lut.py:52-54)
    if rc == E_LUT_SMALL:
        raise IndexError(msg)                 # numpy IndexError in the reference
    if rc == E_BAD_ARG:
        raise ValueError(msg)
    if rc == E_NOMEM:
        raise MemoryError(msg)
    raise MulutError(msg or "mulut error {}".format(rc))
