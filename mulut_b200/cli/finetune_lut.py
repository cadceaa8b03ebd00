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


def finetune_steps(model_G: MuLUT, steps: int, batch: int, crop: int, seed: int, lr0: float, lr1: float,
                   total_iter: int, weight_decay: float = 0.0, batches=None):
    """The training loop body of 3_finetune_lut.py:118-136 on this rank's share."""
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
                            opt.weightDecay)
    if rank == 0:
        print("{} | Iter:{:6d}, loss:{:.3e}, rT:{:.4f}".format(opt.expDir, opt.totalIter, np.mean(losses[-100:]),
                                                          (time.time() - st) / max(1, opt.totalIter)))
        model_G.export_luts(opt.expDir)
        print("Finetuned LUT saved to {}".format(opt.expDir))


if __name__ == "__main__":
    main(sys.argv[1:])
