"""Import the UNMODIFIED reference from /root/reference (build container only).

TEST INFRASTRUCTURE.  /root/reference does not exist on the GPU box, so this
module is used only (a) by oracle/make_golden.py to generate the committed
fixtures under tests/golden/ and (b) by CPU tests that are skipped when the
reference tree is absent.  Nothing is copied from the reference: the modules
are loaded from where they lie.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np

REF_ROOT = os.environ.get("MULUT_REFERENCE_ROOT", "/root/reference")
# the GPU box has no /root/reference: there the byte-identical staged copy made by `python -m oracle.fetch_ref`
# (git-ignored, shipped with the snapshot) stands in for the modules of the hot path
_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
if not os.path.isfile(os.path.join(REF_ROOT, "sr", "4_test_lut.py")) and \
        os.path.isfile(os.path.join(_STAGED, "sr", "4_test_lut.py")):
    REF_ROOT = _STAGED


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "sr", "4_test_lut.py"))


def full_tree() -> bool:
    """The whole reference repository (data, results, models), not just the staged hot-path modules."""
    return os.path.isdir(os.path.join(REF_ROOT, "models", "sr_x2sdy"))


_cache = {}


def _load(name: str, path: str):
    if name in _cache:
        return _cache[name]
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)       # for `common.*`
    sr = os.path.join(REF_ROOT, "sr")
    if sr not in sys.path:
        sys.path.insert(0, sr)
    old_argv = sys.argv
    sys.argv = ["x"]
    try:
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.argv = old_argv
    _cache[name] = mod
    return mod


def test_lut_module():
    """sr/4_test_lut.py (file name starts with a digit -> importlib)."""
    return _load("ref_4_test_lut", os.path.join(REF_ROOT, "sr", "4_test_lut.py"))


def model_module():
    """sr/model.py (MuLUT with InterpTorchBatch)."""
    return _load("ref_sr_model", os.path.join(REF_ROOT, "sr", "model.py"))


def utils_module():
    """common/utils.py (modcrop, _rgb2ycbcr, PSNR, cal_ssim: the reference's report metrics)."""
    return _load("ref_common_utils", os.path.join(REF_ROOT, "common", "utils.py"))


def transfer_module():
    """sr/2_transfer_to_lut.py (the LUT producer; only its grid-enumeration helpers are used)."""
    return _load("ref_2_transfer_to_lut", os.path.join(REF_ROOT, "sr", "2_transfer_to_lut.py"))


def ref_pipeline(img_u8, luts, stages, modes, scale, interval=4):
    """The reference's own stage/mode/rotation loop (4_test_lut.py:279-306)
    re-driven around the reference's FourSimplexInterpFaster.  ``luts`` maps
    "s{stage}_{mode}" -> int8 array; converted the way 4_test_lut.py:333 does."""
    fn = test_lut_module().FourSimplexInterpFaster
    img_lr = np.asarray(img_u8).astype(np.float32)
    if img_lr.ndim == 2:
        img_lr = np.expand_dims(img_lr, axis=2)
        img_lr = np.concatenate([img_lr, img_lr, img_lr], axis=2)
    for s in range(stages):
        pred = 0
        if (s + 1) == stages:
            upscale = scale
            avg_factor, bias = len(modes), 0
        else:
            upscale = 1
            avg_factor, bias = len(modes) * 4, 127
        for mode in modes:
            key = "s{}_{}".format(s + 1, mode)
            lut = np.asarray(luts[key]).astype(np.float32).reshape(-1, upscale * upscale)
            pad = (0, 2) if mode in ["d", "y"] else (0, 1)
            for r in [0, 1, 2, 3]:
                rot = np.rot90(img_lr, r)
                h, w, _ = rot.shape
                img_in = np.pad(rot, (pad, pad, (0, 0)), mode="edge").transpose((2, 0, 1))
                pred = pred + fn(lut, img_in, h, w, interval, 4 - r, upscale=upscale, mode=mode)
        img_lr = np.clip((pred / avg_factor) + bias, 0, 255)
        img_lr = img_lr.transpose((1, 2, 0))
        img_lr = np.round(np.clip(img_lr, 0, 255))
        img_lr = img_lr.astype(np.uint8) if (s + 1) == stages else img_lr.astype(np.float32)
    return img_lr


def shipped_luts():
    d = os.path.join(REF_ROOT, "models", "sr_x2sdy")
    luts = {}
    for s in (1, 2):
        for m in "sdy":
            luts["s{}_{}".format(s, m)] = np.load(
                os.path.join(d, "LUT_ft_x4_4bit_int8_s{}_{}.npy".format(s, m))).reshape(-1, 1 if s == 1 else 16)
    return luts
