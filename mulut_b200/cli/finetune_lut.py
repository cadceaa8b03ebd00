"""`python -m mulut_b200.cli.finetune_lut ...` (or under torchrun, one rank per GPU)

Drop-in for the reference's step 3 (sr/3_finetune_lut.py:68-172): same flags,
Adam(lr0, betas .9/.999, eps 1e-8) with the cosine LambdaLR lr0 -> lr1
(:85-95), MSE loss, final LUT export (:162-169).  Differences: the interpolation
runs in the sm_100a kernels; multi-GPU is real data parallelism (one NCCL
all-reduce of the flat 17 MB LUT-gradient buffer per step) instead of the
reference's non-functional `--gpuNum` branch (:156-157); `--synthetic` trains on
seeded random patches because the DIV2K set is not shipped.
"""
from __future__ import annotations

import math
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

from .. import dist as mdist
from ..model import MuLUT
from ..options import TrainOptions


def lr_lambda(total_iter: int, lr0: float, lr1: float):
    """3_finetune_lut.py:88-94."""
    if lr1 < 0:
        return lambda x: ((1 + math.cos(x * math.pi / total_iter)) / 2) * 0.8 + 0.2
    lr_b = lr1 / lr0
    lr_a = 1 - lr_b
    return lambda x: ((1 + math.cos(x * math.pi / total_iter)) / 2) * lr_a + lr_b


def synthetic_batch(batch: int, crop: int, scale: int, seed: int, device):
    """im [B,1,crop,crop], lb [B,1,crop*r,crop*r] in [0,1] on the k/255 grid
    (the shapes data.Provider yields, sr/data.py:35-39)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    im = torch.randint(0, 256, (batch, 1, crop, crop), generator=g).float() / 255.0
    lb = torch.randint(0, 256, (batch, 1, crop * scale, crop * scale), generator=g).float() / 255.0
    return im.to(device), lb.to(device)


class GraphedStep:
    """One training step of 3_finetune_lut.py:129-136 (zero_grad -> forward -> MSE -> backward ->
    LUT-gradient all-reduce -> Adam) captured ONCE in a CUDA graph and replayed: the step is ~10
    kernels of 0.2-1 ms, so the ~60 launches and the autograd bookkeeping of an eager step cost as
    much host time as the GPU needs.  The learning rate lives in a device tensor (Adam capturable),
    the batch in static buffers; the warm-up iterations CUDA graphs need are rolled back."""

    def __init__(self, model_G, im_shape, lb_shape, lr0, weight_decay=0.0):
        import copy
        device = next(model_G.parameters()).device
        self.model = model_G
        self.params = [p for p in model_G.parameters() if p.requires_grad]
        self.bucket = mdist.FlatGradBucket(self.params)
        self.opt = mdist.FusedAdam(self.bucket, lr0, betas=(0.9, 0.999), eps=1e-8, weight_decay=weight_decay)
        self.im = torch.zeros(im_shape, device=device)
        self.lb = torch.zeros(lb_shape, device=device)
        saved = [p.detach().clone() for p in self.params]
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(3):
                self._body()
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        with torch.no_grad():                    # roll the warm-up back: parameters and Adam state
            for p, q in zip(self.params, saved):
                p.copy_(q)
            self.opt.reset_state()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._body()

    def _body(self):
        self.bucket.zero_()
        loss = F.mse_loss(self.model(self.im), self.lb)
        loss.backward()
        self.bucket.all_reduce_mean()
        self.opt.step()
        return loss.detach()

    def __call__(self, im, lb, lr: float):
        self.im.copy_(im)
        self.lb.copy_(lb)
        self.opt.set_lr(lr)
        self.graph.replay()
        return self.loss.clone()                  # device scalar (the graph's output buffer is reused); .item() synchronises


def finetune_steps(model_G: MuLUT, steps: int, batch: int, crop: int, seed: int, lr0: float, lr1: float,
                   total_iter: int, weight_decay: float = 0.0, batches=None, graph: bool = False):
    """The training loop body of 3_finetune_lut.py:118-136 on this rank's share.
    graph=True replays the step as one CUDA graph (same arithmetic; needs a fixed batch shape)."""
    if graph:
        rank, world, _ = mdist.env_rank_world()
        device = next(model_G.parameters()).device
        r = model_G.upscale
        C = 1 if batches is None else batches[0][0].shape[1]
        b0 = batch if batches is None else batches[0][0].shape[0]
        model_G.train()
        gs = GraphedStep(model_G, (b0, C, crop, crop), (b0, C, crop * r, crop * r), lr0, weight_decay)
        lam = lr_lambda(total_iter, lr0, lr1)
        losses = []
        for i in range(steps):
            im, lb = batches[i] if batches is not None else synthetic_batch(batch, crop, r, seed + i * world + rank, device)
            losses.append(gs(im, lb, lr0 * lam(i)))     # LambdaLR: lr of step i is lr0 * lambda(i)
        return [float(l) for l in torch.stack(losses).cpu()] if losses else []
    rank, world, _ = mdist.env_rank_world()
    device = next(model_G.parameters()).device
    params = [p for p in model_G.parameters() if p.requires_grad]
    bucket = mdist.FlatGradBucket(params)
    opt_G = torch.optim.Adam(params, lr=lr0, betas=(0.9, 0.999), eps=1e-8, weight_decay=weight_decay)
    sched = torch.optim.lr_scheduler.LambdaLR(opt_G, lr_lambda=lr_lambda(total_iter, lr0, lr1))
    losses = []
    model_G.train()
    for i in range(steps):
        if batches is not None:
            im, lb = batches[i]
        else:
            im, lb = synthetic_batch(batch, crop, model_G.upscale, seed + i * world + rank, device)
        bucket.zero_()
        pred = model_G(im)
        loss = F.mse_loss(pred, lb)
        loss.backward()
        bucket.all_reduce_mean()
        opt_G.step()
        sched.step()
        losses.append(float(loss.item()))
    return losses


def main(argv=None):
    opt = TrainOptions().parse(argv)
    rank, world, local = mdist.init_process_group()
    device = torch.device("cuda", local if world > 1 else opt.device)
    torch.cuda.set_device(device)
    modes = [m for m in opt.modes]
    model_G = MuLUT(lut_folder=opt.expDir, stages=opt.stages, modes=modes, upscale=opt.scale,
                    interval=opt.interval).to(device)
    if not opt.synthetic:
        raise SystemExit("only --synthetic training data is available in this build "
                         "(DIV2K is not shipped; see DESIGN.md, out of scope: data providers)")
    per_rank = max(1, opt.batchSize // world)
    st = time.time()
    losses = finetune_steps(model_G, opt.totalIter, per_rank, opt.cropSize, 0, opt.lr0, opt.lr1, opt.totalIter,
                            opt.weightDecay, graph=not getattr(opt, "eager", False))
    if rank == 0:
        print("{} | Iter:{:6d}, loss:{:.3e}, rT:{:.4f}".format(opt.expDir, opt.totalIter, np.mean(losses[-100:]),
                                                          (time.time() - st) / max(1, opt.totalIter)))
        model_G.export_luts(opt.expDir)
        print("Finetuned LUT saved to {}".format(opt.expDir))
    if world > 1:
        # communicators captured in a CUDA graph can hang in their destructor: leave without tearing them down
        import os
        torch.cuda.synchronize()
        torch.distributed.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main(sys.argv[1:])
