"""The writers of the on-disk LUT format.

* :func:`get_input_tensor` / :func:`get_mode_input_tensor` enumerate the L^4 sampling grid exactly as
  the reference's `sr/2_transfer_to_lut.py:12-66` does: row  a*L^3 + b*L^2 + c*L + d  of the result is
  the 2x2 patch [[a, b], [c, d]] (values 0, q, 2q, ..., 255 - the last grid point is 255, not 256), placed
  on the mode's 3x3 footprint for the d / y modes.
* :func:`transfer_to_lut` runs a caller-supplied network over the grid in chunks and writes
  `LUT_x{scale}_{interval}bit_int8_s{stage}_{mode}.npy` with `round(clamp(out, -1, 1) * 127).int8`
  (`:85-116`).  The convolutional MuLUT networks themselves are out of scope: `net_fn`
  is whatever produced them - `net_fn(batch [n,1,k,k] in [0,1], stage, mode) -> [n,1,r,r]`.
* The fine-tuning exporter (`sr/3_finetune_lut.py:162-169`) is `mulut_b200.model.MuLUT.export_luts`.
"""
from __future__ import annotations

import os
from typing import Callable, Iterable

import numpy as np
import torch


def get_input_tensor(interval: int = 4, device="cpu") -> torch.Tensor:
    """[L^4, 1, 2, 2] float32 in [0, 1]; 2_transfer_to_lut.py:12-42."""
    base = torch.arange(0, 257, 2 ** interval, device=device)
    base[-1] -= 1
    L = base.numel()
    a, b, c, d = torch.meshgrid(base, base, base, base, indexing="ij")      # a slowest, d fastest
    grid = torch.stack([a, b, c, d], dim=-1).reshape(L ** 4, 1, 2, 2)
    return grid.float() / 255.0


def get_mode_input_tensor(input_tensor: torch.Tensor, mode: str) -> torch.Tensor:
    """Scatter the 2x2 taps onto the mode's footprint; 2_transfer_to_lut.py:45-66."""
    if mode == "s":
        return input_tensor
    if mode == "d":
        pos = ((0, 0), (0, 2), (2, 0), (2, 2))
    elif mode == "y":
        pos = ((0, 0), (1, 1), (1, 2), (2, 1))
    else:
        raise ValueError("Mode {} not implemented.".format(mode))
    out = torch.zeros((input_tensor.shape[0], input_tensor.shape[1], 3, 3), dtype=input_tensor.dtype,
                      device=input_tensor.device)
    for k, (y, x) in enumerate(pos):
        out[:, :, y, x] = input_tensor[:, :, k // 2, k % 2]
    return out


def quantize_lut(batch_output: torch.Tensor) -> np.ndarray:
    """round(clamp(out, -1, 1) * 127) as int8; 2_transfer_to_lut.py:104-105."""
    return torch.round(torch.clamp(batch_output, -1, 1) * 127).cpu().numpy().astype(np.int8)


def lut_file_name(scale: int, interval: int, stage: int, mode: str) -> str:
    return "LUT_x{}_{}bit_int8_s{}_{}.npy".format(scale, interval, stage, mode)     # :110-111


def transfer_to_lut(net_fn: Callable, stages: int, modes: Iterable[str], scale: int, interval: int = 4,
                    exp_dir: str = None, device="cpu", chunks: int = 100) -> dict:
    """The main loop of 2_transfer_to_lut.py:85-116.  Returns {"s{stage}_{mode}": int8 array
    [L^4, 1, r, r]} and, when exp_dir is given, saves each table under the reference's file name."""
    luts = {}
    for s in range(stages):
        stage = s + 1
        for mode in modes:
            x = get_input_tensor(interval, device)
            if mode != "s":
                x = get_mode_input_tensor(x, mode)
            B = max(1, x.size(0) // chunks)
            outs = []
            with torch.no_grad():
                for b0 in range(0, x.size(0), B):
                    outs.append(quantize_lut(net_fn(x[b0:b0 + B], stage, mode)))
            res = np.concatenate(outs, 0)
            luts["s{}_{}".format(stage, mode)] = res
            if exp_dir is not None:
                np.save(os.path.join(exp_dir, lut_file_name(scale, interval, stage, mode)), res)
    return luts
