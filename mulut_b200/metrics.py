"""Evaluation helpers used by the CLI (host side, numpy): BT.601 luma, PSNR with a
border shave and the 11x11 Gaussian SSIM.  Same definitions as the reference's
common/utils.py:28-101 so the printed PSNR/SSIM agree; not on the hot path."""
from __future__ import annotations

import numpy as np

_BT601 = np.array([[65.481, 128.553, 24.966],
                   [-37.797, -74.203, 112.0],
                   [112.0, -93.786, -18.214]]) / 255.0
_OFFSET = np.array([16.0, 128.0, 128.0])


def modcrop(image: np.ndarray, modulo: int) -> np.ndarray:
    h, w = image.shape[:2]
    return image[: h - h % modulo, : w - w % modulo]


def rgb2ycbcr(img: np.ndarray) -> np.ndarray:
    """(H,W,3) RGB in [0,255] -> YCbCr, ITU-R BT.601 studio swing."""
    return np.asarray(img, dtype=np.float64) @ _BT601.T + _OFFSET


def PSNR(y_true, y_pred, shave_border: int = 4) -> float:
    d = np.asarray(y_pred, dtype=np.float32) - np.asarray(y_true, dtype=np.float32)
    if shave_border > 0:
        d = d[shave_border:-shave_border, shave_border:-shave_border]
    rmse = np.sqrt(np.mean(d * d))
    return float(20.0 * np.log10(255.0 / rmse))


def _gauss_window(size: int = 11, sigma: float = 1.5) -> np.ndarray:
    ax = np.arange(size) - (size - 1) / 2.0
    g = np.exp(-(ax * ax) / (2 * sigma * sigma))
    g /= g.sum()
    return np.outer(g, g)


def cal_ssim(img1, img2) -> float:
    from scipy import signal
    win = _gauss_window()
    c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    a, b = np.float64(img1), np.float64(img2)
    conv = lambda z: signal.convolve2d(z, win, "valid")
    mu1, mu2 = conv(a), conv(b)
    s1, s2, s12 = conv(a * a) - mu1 * mu1, conv(b * b) - mu2 * mu2, conv(a * b) - mu1 * mu2
    m = ((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 * mu1 + mu2 * mu2 + c1) * (s1 + s2 + c2))
    return float(m.mean())


def psnr_ssim_device(gt, img, shave_border: int = 4):
    """(PSNR, SSIM) on the BT.601 luma of two RGB uint8 CUDA tensors (H,W,3), computed on the GPU
    (mulut_eval_psnr_ssim_y_u8): the same definitions as PSNR(rgb2ycbcr(.)[..., 0]) and cal_ssim above."""
    import ctypes
    import torch
    from . import _lib
    if not (isinstance(gt, torch.Tensor) and isinstance(img, torch.Tensor) and gt.is_cuda and img.is_cuda):
        raise TypeError("psnr_ssim_device expects CUDA uint8 tensors (no CPU fallback; use PSNR / cal_ssim on the host)")
    if gt.dtype != torch.uint8 or img.dtype != torch.uint8 or gt.shape != img.shape or gt.dim() != 3 or gt.shape[2] != 3:
        raise ValueError("expected two uint8 (H,W,3) tensors of the same shape")
    g, s = gt.contiguous(), img.contiguous()
    work = torch.empty(32, dtype=torch.uint8, device=g.device)
    out = (ctypes.c_double * 2)()
    stream = ctypes.c_void_p(torch.cuda.current_stream(g.device).cuda_stream)
    with torch.cuda.device(g.device):
        _lib.check(_lib.lib().mulut_eval_psnr_ssim_y_u8(g.data_ptr(), s.data_ptr(), g.shape[0], g.shape[1],
                                                        int(shave_border), work.data_ptr(), out, stream))
    return float(out[0]), float(out[1])
