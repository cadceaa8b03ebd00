"""fp32 torch oracle for the differentiable twin (TEST INFRASTRUCTURE ONLY).

Restates, with plain torch ops and autograd, what the reference computes in
    MuLUT.InterpTorchBatch   /root/reference/sr/model.py:69-287
    MuLUT.round_func         /root/reference/sr/model.py:59-67
    MuLUT.forward            /root/reference/sr/model.py:289-312
The reference enumerates 24 strict-inequality cases; here the four fractions are
sorted descending with ties broken "higher tap index first" (the order the
reference's case cascade resolves to) and the simplex is walked.
Gradients come from torch autograd, so this is the checker for K2/K3.

Pinned by tests/test_oracle_pinned.py against fixtures generated from the
reference itself (oracle/make_golden.py -> tests/golden/finetune_*.npz), and,
when /root/reference is present, against the reference module directly.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

MODE_TAPS = {
    "s": ((0, 0), (0, 1), (1, 0), (1, 1)),
    "d": ((0, 0), (0, 2), (2, 0), (2, 2)),
    "y": ((0, 0), (1, 1), (1, 2), (2, 1)),
}
MODE_PAD = {"s": 1, "d": 2, "y": 2}


def round_ste(x):
    return x + (torch.round(x) - x).detach()


def interp_torch_batch(weight, upscale, mode, img_in, bd, interval=4):
    """model.py:69-287.  weight (n_rows, up^2) fp32 parameter; img_in (B,C,h+bd,w+bd)."""
    if mode not in MODE_TAPS:
        raise ValueError("Mode {} not implemented.".format(mode))
    _, _, H, W = img_in.shape
    h, w = H - bd, W - bd
    q = 2 ** interval
    L = 2 ** (8 - interval) + 1
    wq = torch.clamp(round_ste(weight * 127), -127, 127)          # model.py:74-76
    taps = [img_in[:, :, dy:dy + h, dx:dx + w] for dy, dx in MODE_TAPS[mode]]
    m = [torch.floor_divide(t, q).to(torch.int64) for t in taps]
    f = [t % q for t in taps]                                       # differentiable w.r.t. img_in
    strides = [L * L * L, L * L, L, 1]
    v0 = m[0] * strides[0] + m[1] * strides[1] + m[2] * strides[2] + m[3] * strides[3]
    # stable descending sort over taps listed d,c,b,a  => ties keep the higher tap first
    fr = torch.stack([f[3], f[2], f[1], f[0]], dim=0)
    st = torch.tensor([strides[3], strides[2], strides[1], strides[0]], device=img_in.device)
    fs, idx = torch.sort(fr.detach(), dim=0, descending=True, stable=True)
    fs = torch.gather(fr, 0, idx)                                   # keep the autograd path
    ss = st[idx]
    verts = [v0]
    for k in range(4):
        verts.append(verts[-1] + ss[k])
    wts = [q - fs[0], fs[0] - fs[1], fs[1] - fs[2], fs[2] - fs[3], fs[3]]
    out = 0
    for k in range(5):
        p = wq[verts[k].reshape(-1)].reshape(*verts[k].shape, upscale, upscale)
        out = out + wts[k][..., None, None] * p
    B, C = out.shape[0], out.shape[1]
    out = out.permute(0, 1, 2, 4, 3, 5).reshape(B, C, h * upscale, w * upscale)
    return out / q


def mulut_forward(weights: dict, x, stages, modes, upscale, interval=4):
    """model.py:289-312 with weights["s{stage}_{mode}"]; x in [0,1]."""
    x = x * 255.0
    for s in range(stages):
        pred = 0
        stage = s + 1
        if stage == stages:
            avg_factor, bias, scale = len(modes), 0, upscale
        else:
            avg_factor, bias, scale = len(modes) * 4, 127, 1
        for mode in modes:
            pad = MODE_PAD[mode]
            weight = weights["s{}_{}".format(stage, mode)]
            for r in range(4):
                xin = F.pad(torch.rot90(x, r, [2, 3]), (0, pad, 0, pad), mode="replicate")
                o = interp_torch_batch(weight, scale, mode, xin, pad, interval)
                pred = round_ste(pred + torch.rot90(o, (4 - r) % 4, [2, 3]))
        x = round_ste(torch.clamp(pred / avg_factor + bias, 0, 255))
    return x / 255.0
