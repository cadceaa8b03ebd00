"""ctypes binding of oracle/libmulut_oracle.so (TEST INFRASTRUCTURE ONLY).

Used by tests/ (full-size parity), __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by mulut_b200.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmulut_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mulut_oracle.c")
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libmulut_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.mulut_oracle_sr_u8.restype = ctypes.c_int
        _lib.mulut_oracle_sr_u8.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.c_int,
            ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]
        _lib.mulut_oracle_max_threads.restype = ctypes.c_int
    return _lib


def max_threads() -> int:
    return int(lib().mulut_oracle_max_threads())


def sr_u8(frames, luts: dict, stages: int, modes: str, scale: int, interval: int = 4, nthreads: int = 0):
    """frames: uint8 (N,H,W,C) or (H,W,C); returns uint8 (N,H*scale,W*scale,C)."""
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    squeeze = frames.ndim == 3
    if squeeze:
        frames = frames[None]
    N, H, W, C = frames.shape
    modes = "".join(modes)
    tabs = []
    for s in range(stages):
        cols = scale * scale if s + 1 == stages else 1
        for m in modes:
            t = np.ascontiguousarray(np.asarray(luts["s{}_{}".format(s + 1, m)]).reshape(-1, cols), dtype=np.int8)
            tabs.append(t)
    ptrs = (ctypes.c_void_p * len(tabs))(*[t.ctypes.data for t in tabs])
    out = np.empty((N, H * scale, W * scale, C), dtype=np.uint8)
    rc = lib().mulut_oracle_sr_u8(frames.ctypes.data, out.ctypes.data, N, H, W, C, stages,
                                  modes.encode(), scale, interval, ptrs, nthreads)
    if rc == -1:
        bad = [m for m in modes if m not in "sdy"][0]
        raise ValueError("Mode {} not implemented.".format(bad))
    if rc != 0:
        raise RuntimeError("mulut_oracle_sr_u8 failed: {}".format(rc))
    return out[0] if squeeze else out
