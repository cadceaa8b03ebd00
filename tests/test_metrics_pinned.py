"""Pins the report metrics (BT.601 luma, PSNR with border shave, 11x11 Gaussian SSIM) to the reference:
tests/golden/set5_metrics.json holds what the reference's OWN common/utils.py:28-101 computes for its
golden Set5 results against the Set5 HR images (oracle/make_golden.py:set5_hr_metrics), including the
known answer printed at sr/4_test_lut.py:343 (`AVG LUT PSNR: 30.61 SSIM: 0.8655`; the mean SSIM is
0.86556, which today's numpy prints as 0.8656).  CPU: the host definitions in mulut_b200.metrics; GPU: the
on-device kernel mulut_eval_psnr_ssim_y_u8 and the CLI's printed summary."""
import json
import os

import numpy as np
import pytest

from conftest import GOLD
from oracle import ref_import as R

NAMES = ["baby", "bird", "butterfly", "head", "woman"]


@pytest.fixture(scope="module")
def set5_hr():
    return dict(np.load(os.path.join(GOLD, "set5_hr.npz")))


@pytest.fixture(scope="module")
def ref_metrics():
    return json.load(open(os.path.join(GOLD, "set5_metrics.json")))


def test_known_answer_of_the_reference(ref_metrics):
    assert ref_metrics["_printed"].startswith("Dataset Set5 | AVG LUT PSNR: 30.61 SSIM: 0.865")
    assert abs(ref_metrics["_mean"]["psnr"] - 30.61) < 0.005
    assert abs(ref_metrics["_mean"]["ssim"] - 0.8655) < 0.0001


@pytest.mark.parametrize("name", NAMES)
def test_host_metrics_match_reference_numbers(name, set5, set5_hr, ref_metrics):
    from mulut_b200.metrics import PSNR, cal_ssim, modcrop, rgb2ycbcr
    gt = modcrop(set5_hr["hr_" + name], 4)
    sr = set5["sr_" + name]
    assert gt.shape == sr.shape
    y_gt, y_out = rgb2ycbcr(gt)[:, :, 0], rgb2ycbcr(sr)[:, :, 0]
    r = ref_metrics[name]
    assert abs(y_gt.sum() - r["y_gt_sum"]) < 1e-4 * 1e-3 * y_gt.size        # per-pixel luma agrees to ~1e-7
    assert abs(PSNR(y_gt, y_out, 4) - r["psnr"]) < 1e-5
    assert abs(cal_ssim(y_gt, y_out) - r["ssim"]) < 1e-10


@pytest.mark.skipif(not R.available(), reason="reference tree not present")
def test_host_metrics_match_reference_live():
    """Random and ragged images, live against common/utils.py (build container only)."""
    from mulut_b200 import metrics as M
    U = R.utils_module()
    rng = np.random.default_rng(3)
    for (H, W, shave) in [(64, 64, 4), (37, 91, 2), (13, 200, 0), (121, 77, 3)]:
        gt = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        img = np.clip(gt.astype(np.int32) + rng.integers(-20, 21, gt.shape), 0, 255).astype(np.uint8)
        assert (M.modcrop(gt, 4) == U.modcrop(gt, 4)).all()
        a, b = M.rgb2ycbcr(gt), U._rgb2ycbcr(gt)
        assert np.abs(a - b).max() < 1e-10
        y1, y2 = b[:, :, 0], U._rgb2ycbcr(img)[:, :, 0]
        assert abs(M.PSNR(y1, y2, shave) - U.PSNR(y1, y2, shave)) < 1e-6
        assert abs(M.cal_ssim(y1, y2) - U.cal_ssim(y1, y2)) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_device_metrics_match_reference_numbers(name, set5, set5_hr, ref_metrics):
    import torch
    from mulut_b200.metrics import modcrop, psnr_ssim_device
    gt = np.ascontiguousarray(modcrop(set5_hr["hr_" + name], 4))
    p, s = psnr_ssim_device(torch.from_numpy(gt).cuda(), torch.from_numpy(set5["sr_" + name]).cuda(), 4)
    r = ref_metrics[name]
    assert abs(p - r["psnr"]) < 1e-4, (p, r["psnr"])       # the reference averages float32 squares in float32
    assert abs(s - r["ssim"]) < 1e-9, (s, r["ssim"])


@pytest.mark.gpu
def test_cli_prints_the_reference_known_answer(tmp_path, capsys, set5, set5_hr, ref_metrics, gold_dir):
    """sr/4_test_lut.py's flags end to end on the real Set5 tree (LR inputs + HR images from the fixtures):
    bit-exact PNGs, the reference's per-image PSNR / SSIM and its printed summary line."""
    from PIL import Image
    from mulut_b200.cli import test_lut
    root = tmp_path / "data" / "Set5"
    (root / "HR").mkdir(parents=True)
    (root / "LR_bicubic" / "X4").mkdir(parents=True)
    for n in NAMES:
        Image.fromarray(set5["lr_" + n]).save(root / "LR_bicubic" / "X4" / (n + ".png"))
        Image.fromarray(set5_hr["hr_" + n]).save(root / "HR" / (n + ".png"))
    res = test_lut.main(["--stages", "2", "--modes", "sdy", "-e", os.path.join(gold_dir, "luts_x4"),
                         "--testDir", str(tmp_path / "data"), "--resultRoot", str(tmp_path / "results")])
    out = capsys.readouterr().out
    assert ref_metrics["_printed"] in out, out
    assert "AVG LUT PSNR: 30.61 SSIM: 0.865" in out
    for i, n in enumerate(NAMES):
        got = np.array(Image.open(tmp_path / "results" / "luts_x4" / "Set5" / "X4" / (n + "_LUT_ft_4bit.png")))
        assert (got == set5["sr_" + n]).all(), n
        assert abs(res["Set5"][i, 0] - ref_metrics[n]["psnr"]) < 1e-4
        assert abs(res["Set5"][i, 1] - ref_metrics[n]["ssim"]) < 1e-9
