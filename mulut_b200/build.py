"""Builds libmulut_b200.so (in-tree) with nvcc for sm_100a.

    python -m mulut_b200.build [--force] [--verbose]

The shared library is the product's only compute path; there is no CPU
fallback.  The build needs no GPU (nvcc cross-compiles).
"""
from __future__ import annotations

import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OBJ_DIR = os.path.join(_HERE, "csrc", "build")
LIB_PATH = os.path.join(_HERE, "libmulut_b200.so")
SOURCES = ["capi.cu", "tma.cu", "infer_generic.cu", "infer_tiled.cu", "infer_stage1.cu", "infer_binned.cu", "interp_f32.cu", "eval_metrics.cu", "gather_bench.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith(".cuh")) + [os.path.join("..", "..", "include", "mulut.h")]   # every object depends on every header
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isfile(cand) or cand == "nvcc"):
            return cand
    return "nvcc"


def _stale(target: str, deps) -> bool:
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), lib_path: str = LIB_PATH) -> str:
    """defines: extra -D macros (e.g. ["MULUT_BN_TIMING"]) for an instrumented side build; such a
    build goes to its own object directory and `lib_path`, never over the product library."""
    global OBJ_DIR
    obj_dir_saved = OBJ_DIR
    if defines:
        OBJ_DIR = os.path.join(obj_dir_saved, "_".join(defines).lower())
    try:
        return _build(force, verbose, ["-D" + d for d in defines], lib_path)
    finally:
        OBJ_DIR = obj_dir_saved


def _build(force: bool, verbose: bool, extra, lib_path: str) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdrs = [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS] + [os.path.abspath(__file__)]
    objs = []
    nvcc = _nvcc()
    env = dict(os.environ)
    # the image's CC/CXX wrappers are fine for host code; let nvcc pick its default host compiler
    procs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [sp] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", sp, "-o", obj]
            procs.append((src, subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("== nvcc {} ==\n{}\n".format(src, out.decode(errors="replace")))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed (see log above)")
    if force or procs or _stale(lib_path, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib_path] + objs
        subprocess.check_call(cmd, env=env)
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
