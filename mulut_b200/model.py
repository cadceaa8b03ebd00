"""LUT-aware fine-tuning model: drop-in for the reference's `model.MuLUT`
(sr/model.py:39-312) with `InterpTorchBatch` backed by the sm_100a kernels.

Same constructor, parameter names (`weight_s{stage}_{mode}`), file-name rule
(`LUT_x{r}_{interval}bit_int8_s{stage}_{mode}.npy`, model.py:53-54), forward
semantics (per-rotation BPDA rounding, model.py:305-309) and `InterpTorchBatch`
signature.  Autograd is supplied by a torch.autograd.Function that calls
mulut_interp_fwd_f32 / mulut_interp_bwd_f32 through the C ABI; there is no
ATen fallback (CPU tensors raise).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib

mode_pad_dict = {"s": 1, "d": 2, "y": 2, "e": 3, "h": 3, "o": 3}    # model.py:12


class _InterpFunction(torch.autograd.Function):
    """K2/K3: forward and backward of MuLUT.InterpTorchBatch (model.py:69-287)."""

    @staticmethod
    def forward(ctx, weight, img_in, upscale, mode, bd, interval):
        if not (weight.is_cuda and img_in.is_cuda):
            raise RuntimeError("mulut_b200 InterpTorchBatch needs CUDA tensors (no CPU fallback)")
        w = weight.detach().contiguous().float()
        x = img_in.detach().contiguous().float()
        B, C, Hp, Wp = x.shape
        h, wd = Hp - bd, Wp - bd
        out = torch.empty((B, C, h * upscale, wd * upscale), dtype=torch.float32, device=x.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().mulut_interp_fwd_f32(w.data_ptr(), w.shape[0], upscale, mode.encode(), x.data_ptr(),
                                                       B, C, h, wd, bd, interval, out.data_ptr(), stream))
        ctx.save_for_backward(w, x)
        ctx.cfg = (upscale, mode, bd, interval, B, C, h, wd)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        w, x = ctx.saved_tensors
        upscale, mode, bd, interval, B, C, h, wd = ctx.cfg
        g = grad_out.contiguous().float()
        gw = torch.zeros_like(w) if ctx.needs_input_grad[0] else None
        gx = torch.zeros_like(x) if ctx.needs_input_grad[1] else None
        stream = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().mulut_interp_bwd_f32(
                w.data_ptr(), w.shape[0], upscale, mode.encode(), x.data_ptr(), B, C, h, wd, bd, interval,
                g.data_ptr(), gw.data_ptr() if gw is not None else None,
                gx.data_ptr() if gx is not None else None, stream))
        return gw, gx, None, None, None, None


class _StageFunction(torch.autograd.Function):
    """K4: one whole stage of MuLUT.forward (model.py:296-310) - 4*len(modes) interpolation
    passes with their per-pass rounding, the average, bias, clamp and final rounding - as
    ONE kernel per direction (mulut_stage_fwd_f32 / mulut_stage_bwd_f32)."""

    @staticmethod
    def forward(ctx, x, upscale, modes, avg, bias, interval, grad_targets, *weights):
        # grad_targets: None, or one fp32 buffer per weight (e.g. the views of a FlatGradBucket) that the backward
        # kernel accumulates INTO directly - autograd then gets None for the weights, which saves the zero-filled
        # temporaries and the six accumulation kernels of the usual `.grad += returned gradient` (40 us per step)
        if not (x.is_cuda and all(w.is_cuda for w in weights)):
            raise RuntimeError("mulut_b200 fused stage needs CUDA tensors (no CPU fallback)")
        ctx.grad_targets = grad_targets
        xs = x.detach().contiguous().float()
        ws = [w.detach().contiguous().float() for w in weights]
        B, C, h, wd = xs.shape
        out = torch.empty((B, C, h * upscale, wd * upscale), dtype=torch.float32, device=xs.device)
        mask = torch.empty(out.shape, dtype=torch.uint8, device=xs.device)
        # quantised int8 rows + clamp flags of this stage's tables: written by the forward, re-read by the backward
        qws = torch.empty(_lib.lib().mulut_stage_workspace_bytes(len(ws), ws[0].shape[0], upscale), dtype=torch.uint8,
                          device=xs.device)
        ptrs = (ctypes.c_void_p * len(ws))(*[w.data_ptr() for w in ws])
        stream = ctypes.c_void_p(torch.cuda.current_stream(xs.device).cuda_stream)
        with torch.cuda.device(xs.device):
            _lib.check(_lib.lib().mulut_stage_fwd_f32(ptrs, len(ws), modes.encode(), ws[0].shape[0], upscale, interval,
                                                      xs.data_ptr(), B, C, h, wd, float(avg), float(bias),
                                                      out.data_ptr(), mask.data_ptr(), qws.data_ptr(), stream))
        ctx.save_for_backward(xs, mask, qws, *ws)
        ctx.cfg = (upscale, modes, float(avg), float(bias), interval)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        xs, mask, qws, *ws = ctx.saved_tensors
        upscale, modes, avg, bias, interval = ctx.cfg
        B, C, h, wd = xs.shape
        g = grad_out.contiguous().float()
        gx = torch.zeros_like(xs) if ctx.needs_input_grad[0] else None
        direct = ctx.grad_targets is not None
        if direct:
            gws = [t if ctx.needs_input_grad[7 + i] else None for i, t in enumerate(ctx.grad_targets)]
        else:
            gws = [torch.zeros_like(w) if ctx.needs_input_grad[7 + i] else None for i, w in enumerate(ws)]
        ptrs = (ctypes.c_void_p * len(ws))(*[w.data_ptr() for w in ws])
        gptrs = (ctypes.c_void_p * len(ws))(*[(t.data_ptr() if t is not None else None) for t in gws])
        stream = ctypes.c_void_p(torch.cuda.current_stream(xs.device).cuda_stream)
        with torch.cuda.device(xs.device):
            _lib.check(_lib.lib().mulut_stage_bwd_f32(ptrs, len(ws), modes.encode(), ws[0].shape[0], upscale, interval,
                                                      xs.data_ptr(), B, C, h, wd, avg, bias, g.data_ptr(),
                                                      mask.data_ptr(), qws.data_ptr(), gptrs,
                                                      gx.data_ptr() if gx is not None else None, stream))
        return (gx, None, None, None, None, None, None) + (tuple(None for _ in gws) if direct else tuple(gws))


class _MseHead(torch.autograd.Function):
    """`F.mse_loss(x / 255.0, label)` (sr/model.py:312 + sr/3_finetune_lut.py:132) as one kernel per direction
    (mulut_mse_head_fwd_f32 / _bwd_f32): ATen needs six kernels over the 9.4 M outputs of a cfg-4 step for it."""

    @staticmethod
    def forward(ctx, x, label):
        if not (x.is_cuda and label.is_cuda):
            raise RuntimeError("mulut_b200 loss head needs CUDA tensors (no CPU fallback)")
        if x.shape != label.shape:
            raise ValueError("prediction and label shapes differ")
        xs, lb = x.detach().contiguous().float(), label.detach().contiguous().float()
        loss = torch.empty((), dtype=torch.float32, device=xs.device)
        work = torch.empty(16, dtype=torch.uint8, device=xs.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(xs.device).cuda_stream)
        with torch.cuda.device(xs.device):
            _lib.check(_lib.lib().mulut_mse_head_fwd_f32(xs.data_ptr(), lb.data_ptr(), xs.numel(), 1.0 / 255.0,
                                                         work.data_ptr(), loss.data_ptr(), stream))
        ctx.save_for_backward(xs, lb)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        xs, lb = ctx.saved_tensors
        gl = grad_loss.detach().contiguous().float()
        gx = torch.empty_like(xs)
        stream = ctypes.c_void_p(torch.cuda.current_stream(xs.device).cuda_stream)
        with torch.cuda.device(xs.device):
            _lib.check(_lib.lib().mulut_mse_head_bwd_f32(xs.data_ptr(), lb.data_ptr(), xs.numel(), 1.0 / 255.0,
                                                         gl.data_ptr(), gx.data_ptr(), stream))
        return gx, None


def mse_head(x_stage, label):
    """mean((x_stage / 255 - label)^2) with its gradient, fused (x_stage: the last stage's output, 0..255)."""
    return _MseHead.apply(x_stage, label)


def fused_stage(x, weights, upscale, modes, avg, bias, interval=4, check_inputs=False, grad_targets=None):
    """x' = round(clamp(sum-with-per-pass-rounding / avg + bias, 0, 255)) for one stage.

    PRECONDITION: `x` is integer-valued (0..255 as float32) - what MuLUT.forward feeds every stage
    ((k/255)*255 is exact in fp32 when the division is IEEE - a GPU `x / 255` through a reciprocal can be one
    ulp off - and every stage output is rounded).  K4 works on the integer grid: it rounds every sample to
    the nearest integer, where the reference's float floor_divide / % (model.py:123-131) keeps a fraction.
    check_inputs=True rejects inputs further than 1e-3 from an integer (one device synchronisation; not
    inside a CUDA-graph capture); the un-fused path (MuLUT(fused=False), K2/K3) accepts any float input."""
    modes = "".join(modes)
    if check_inputs and not torch.cuda.is_current_stream_capturing():
        if bool(((x.detach() - torch.round(x.detach())).abs() > 1e-3).any()):
            raise ValueError("fused_stage: the input is not integer-valued; use MuLUT(fused=False) for fractional inputs")
    for m in modes:
        if m not in ("s", "d", "y"):
            raise ValueError("Mode {} not implemented.".format(m))
    if grad_targets is not None:
        for w, t in zip(weights, grad_targets):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.shape == w.shape):
                raise ValueError("grad_targets must be contiguous fp32 CUDA buffers shaped like the weights")
    return _StageFunction.apply(x, int(upscale), modes, float(avg), float(bias), int(interval), grad_targets, *weights)


def interp_torch_batch(weight, upscale, mode, img_in, bd, interval=4):
    if mode not in ("s", "d", "y"):
        raise ValueError("Mode {} not implemented.".format(mode))      # model.py:119-121
    return _InterpFunction.apply(weight, img_in, int(upscale), mode, int(bd), int(interval))


class MuLUT(nn.Module):
    """PyTorch version of MuLUT for LUT-aware fine-tuning (kernel-backed)."""

    def __init__(self, lut_folder, stages, modes, upscale=4, interval=4, luts=None, fused=True, check_inputs=False):
        super().__init__()
        # fused=True: every stage is one K4 kernel per direction; fused=False keeps the reference's
        # loop of 4*len(modes) InterpTorchBatch calls per stage (each a K2/K3 kernel) - same results for the
        # inputs the reference flow produces (x on the k/255 grid).  The fused kernels REQUIRE integer-valued
        # stage inputs (see fused_stage); check_inputs=True verifies that on every call (debug: it synchronises)
        self.fused = fused
        self.check_inputs = check_inputs
        self._grad_targets = None            # see accumulate_grads_into()
        self._stage_grads_cb = None          # see on_stage_grads()
        self.interval = interval
        self.upscale = upscale
        self.modes = modes
        self.stages = stages
        for s in range(stages):
            stage = s + 1
            scale = upscale if stage == stages else 1
            for mode in modes:
                key = "s{}_{}".format(stage, mode)
                if luts is not None:
                    arr = np.asarray(luts[key])
                else:
                    arr = np.load(os.path.join(
                        lut_folder, "LUT_x{}_{}bit_int8_s{}_{}.npy".format(upscale, interval, stage, mode)))
                lut_arr = arr.reshape(-1, scale * scale).astype(np.float32) / 127.0
                self.register_parameter(name="weight_" + key, param=nn.Parameter(torch.Tensor(lut_arr)))

    @staticmethod
    def round_func(input):
        """BPDA: round forward, identity backward (model.py:59-67)."""
        return input + (torch.round(input) - input).detach()

    def accumulate_grads_into(self, enabled=True):
        """Fused path only: let the K4 backward accumulate the LUT gradients straight into each parameter's existing
        `.grad` buffer (e.g. the views of a `FlatGradBucket`) instead of returning them to autograd.  The `.grad`
        tensors must exist, be contiguous fp32 and stay the same objects (zero them in place between steps).
        Hooks on the parameters do not see these gradients - which is why this is opt-in."""
        self._grad_targets = bool(enabled)

    def on_stage_grads(self, callback):
        """Fused path with `accumulate_grads_into`: call `callback(stage)` (1-based) during backward as soon as the LUT
        gradients of `stage` are complete, i.e. right after that stage's K4 backward has been launched - a data-parallel
        step starts the all-reduce of the last stage's tables (16 of the 17 MB) under the first stage's backward.
        The first stage reports nothing (its gradients are complete when backward returns).  None removes it."""
        self._stage_grads_cb = callback

    def stage_parameters(self, stage):
        return [getattr(self, "weight_s{}_{}".format(stage, mode)) for mode in self.modes]

    def InterpTorchBatch(self, weight, upscale, mode, img_in, bd):
        return interp_torch_batch(weight, upscale, mode, img_in, bd, self.interval)

    def forward_loss(self, x, label):
        """`F.mse_loss(self.forward(x), label)` with the final `/ 255` and the loss fused into one kernel per direction
        (the fused path's training step; same value and gradients up to fp32 summation order)."""
        if not self.fused:
            return F.mse_loss(self.forward(x), label)
        return mse_head(self._stages(x), label)

    def forward(self, x):
        return self._stages(x) / 255.0

    def _stages(self, x):
        x = x * 255.0
        modes, stages = self.modes, self.stages
        for s in range(stages):
            pred = 0
            stage = s + 1
            if stage == stages:
                avg_factor, bias = len(modes), 0
                scale = self.upscale
            else:
                avg_factor, bias = len(modes) * 4, 127
                scale = 1
            if self.fused:
                for mode in modes:
                    if mode not in ("s", "d", "y"):
                        raise ValueError("Mode {} not implemented.".format(mode))
                weights = [getattr(self, "weight_s{}_{}".format(stage, mode)) for mode in modes]
                targets = None
                if self._grad_targets and torch.is_grad_enabled() and all(w.grad is not None for w in weights):
                    targets = [w.grad for w in weights]
                    if self._stage_grads_cb is not None and x.requires_grad:
                        # fires when the gradient of this stage's INPUT exists: the stage's backward kernel is launched
                        x.register_hook(lambda g, st=stage, cb=self._stage_grads_cb: (cb(st), None)[1])
                x = fused_stage(x, weights, scale, modes, avg_factor, bias, self.interval, self.check_inputs, targets)
                continue
            for mode in modes:
                pad = mode_pad_dict[mode]
                weight = getattr(self, "weight_s{}_{}".format(stage, mode))
                for r in [0, 1, 2, 3]:
                    xin = F.pad(torch.rot90(x, r, [2, 3]), (0, pad, 0, pad), mode="replicate")
                    pred = pred + torch.rot90(self.InterpTorchBatch(weight, scale, mode, xin, pad), (4 - r) % 4, [2, 3])
                    pred = self.round_func(pred)
            x = self.round_func(torch.clamp((pred / avg_factor) + bias, 0, 255))
        return x

    # -- the on-disk format's writer (3_finetune_lut.py:162-169) -------------------
    def export_luts(self, exp_dir=None):
        """round(clip(w,-1,1)*127).int8 per table; optionally saved as
        LUT_ft_x{r}_{interval}bit_int8_s{stage}_{mode}.npy."""
        out = {}
        for s in range(self.stages):
            for mode in self.modes:
                key = "s{}_{}".format(s + 1, mode)
                w = getattr(self, "weight_" + key).detach().cpu().numpy()
                lut = np.round(np.clip(w, -1, 1) * 127).astype(np.int8)
                out[key] = lut
                if exp_dir is not None:
                    np.save(os.path.join(exp_dir, "LUT_ft_x{}_{}bit_int8_s{}_{}.npy".format(
                        self.upscale, self.interval, s + 1, mode)), lut)
        return out
