"""On-device PSNR / SSIM (mulut_eval_psnr_ssim_y_u8) on synthetic and ragged images against the host
definitions in mulut_b200.metrics.  Those host definitions - and the device kernel itself - are pinned to
the reference's common/utils.py:42-101 by tests/test_metrics_pinned.py (reference-computed Set5 numbers,
plus a live comparison when /root/reference is present)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _host(gt, img, shave):
    from mulut_b200.metrics import PSNR, cal_ssim, rgb2ycbcr
    y_gt, y_img = rgb2ycbcr(gt)[:, :, 0], rgb2ycbcr(img)[:, :, 0]
    return PSNR(y_gt, y_img, shave), cal_ssim(y_gt, y_img)


@pytest.mark.parametrize("shape,shave", [((64, 64), 4), ((37, 91), 2), ((11, 11), 0), ((200, 333), 4), ((16, 500), 3)])
def test_device_metrics_match_host(shape, shave):
    import torch
    from mulut_b200.metrics import psnr_ssim_device
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    H, W = shape
    yy, xx = np.mgrid[0:H, 0:W]
    base = (127 + 90 * np.sin(xx / 9.0) * np.cos(yy / 7.0))[..., None] + rng.normal(0, 12, (H, W, 3))
    gt = np.clip(base, 0, 255).astype(np.uint8)
    img = np.clip(gt.astype(np.int32) + rng.integers(-9, 10, gt.shape), 0, 255).astype(np.uint8)
    p_ref, s_ref = _host(gt, img, shave)
    p, s = psnr_ssim_device(torch.from_numpy(gt).cuda(), torch.from_numpy(img).cuda(), shave)
    assert abs(p - p_ref) < 1e-4, (p, p_ref)          # numpy averages the float32 squares in float32
    assert abs(s - s_ref) < 1e-9, (s, s_ref)


def test_device_metrics_edge_cases():
    import torch
    from mulut_b200.metrics import psnr_ssim_device
    a = torch.full((32, 40, 3), 77, dtype=torch.uint8, device="cuda")
    p, s = psnr_ssim_device(a, a.clone(), 4)
    assert p == float("inf") and abs(s - 1.0) < 1e-12
    small = torch.zeros((8, 8, 3), dtype=torch.uint8, device="cuda")      # smaller than the SSIM window
    p, s = psnr_ssim_device(small, small, 0)
    assert np.isnan(s)
    with pytest.raises(TypeError):
        psnr_ssim_device(a.cpu(), a.cpu())
    with pytest.raises(ValueError):
        psnr_ssim_device(a, a[:, :, :2])
