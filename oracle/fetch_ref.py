"""Recipe: stage the UNMODIFIED reference modules of the hot path under oracle/_ref/ (TEST INFRASTRUCTURE).

    python -m oracle.fetch_ref          # build container only: needs /root/reference

oracle/_ref/ is git-ignored (nothing of the reference enters the history) but not gpurun-ignored, so the
staged files travel to the GPU box with the snapshot - like a compiled oracle/_ref/*.so would for a C
reference.  The reference is pure Python: "building" it is copying the five modules its two hot-path
entry points import, byte for byte (the recipe verifies the copies against the originals):

    sr/4_test_lut.py   FourSimplexInterpFaster + the eltr._worker loop      (numpy CPU path)
    sr/model.py        MuLUT with InterpTorchBatch / forward                 (torch path, CPU or CUDA)
    common/{option,utils,network}.py   what those two import

Consumers (only the places the oracle may run): bench.py's `cpu_baseline.numpy_ref` leg times the numpy
path on the box's host cores; bench.py's `finetune.reference_module` leg times model.MuLUT on the B200
through ATen - the bar K4 is measured against.  oracle/ref_import.py loads the modules from
/root/reference when that exists and from oracle/_ref otherwise.
"""
from __future__ import annotations

import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("MULUT_REFERENCE_ROOT", "/root/reference")
FILES = ["sr/4_test_lut.py", "sr/model.py", "common/option.py", "common/utils.py", "common/network.py"]


def staged() -> bool:
    return all(os.path.isfile(os.path.join(DEST, f)) for f in FILES)


def fetch(verbose: bool = True) -> bool:
    """Copy the modules (no-op without the reference tree).  Returns staged()."""
    if not os.path.isfile(os.path.join(SRC, FILES[0])):
        if verbose:
            print("oracle.fetch_ref: {} not present; staged copy {}".format(SRC, "found" if staged() else "absent"))
        return staged()
    for f in FILES:
        dst = os.path.join(DEST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        src = os.path.join(SRC, f)
        if not (os.path.isfile(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
            os.chmod(dst, 0o644)
        assert filecmp.cmp(src, dst, shallow=False), f
    if verbose:
        print("oracle.fetch_ref: {} files staged under {}".format(len(FILES), DEST))
    return True


if __name__ == "__main__":
    sys.exit(0 if fetch() else 1)
