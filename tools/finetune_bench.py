#!/usr/bin/env python
"""cfg4 benchmark: one LUT fine-tuning step (MuLUT.forward + mse_loss + backward +
LUT-gradient all-reduce + Adam), batch 256 of 48x48 patches, x4 sdy 2-stage, shipped
LUTs.  One process per GPU (torchrun); the batch is split across ranks.

    python tools/finetune_bench.py [--batch 256] [--steps 50] [--eager] [--smooth]

(The plain-ATen restatement of the reference's torch path is timed by tests/aten_restatement_timing.py:
the oracle is test infrastructure and is not imported from here.)
Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def shipped_luts():
    d = os.path.join(ROOT, "tests", "golden", "luts_x4")
    return {"s{}_{}".format(s, m): np.load(os.path.join(d, "LUT_ft_x4_4bit_int8_s{}_{}.npy".format(s, m))).reshape(
        -1, 1 if s == 1 else 16) for s in (1, 2) for m in "sdy"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256, help="GLOBAL batch (split across ranks)")
    ap.add_argument("--crop", type=int, default=48)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--eager", action="store_true", help="eager launches with per-phase CUDA events instead of the CUDA-graph step")
    ap.add_argument("--smooth", action="store_true", help="low-frequency patches (neighbouring pixels share LUT rows, like natural images) instead of uniform noise")
    ap.add_argument("--loop", action="store_true", help="reference-style loop of 24 InterpTorchBatch calls (K2/K3) instead of the fused stages (K4)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    from mulut_b200.dist import FlatGradBucket
    from mulut_b200.cli.finetune_lut import lr_lambda, synthetic_batch
    from mulut_b200.model import MuLUT

    dev = torch.device("cuda", local)
    luts = shipped_luts()
    net = MuLUT(None, 2, ["s", "d", "y"], upscale=4, interval=4, luts=luts, fused=not args.loop).to(dev)
    params = list(net.parameters())
    bucket = FlatGradBucket(params)
    opt = torch.optim.Adam(params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=lr_lambda(200000, 1e-3, 1e-4))
    per_rank = args.batch // world
    im, lb = synthetic_batch(per_rank, args.crop, 4, 1000 + rank, dev)
    if args.smooth:
        g = torch.Generator(device="cpu").manual_seed(7 + rank)
        coarse = torch.randint(0, 256, (per_rank, 1, args.crop // 8 + 2, args.crop // 8 + 2), generator=g).float()
        im = torch.round(F.interpolate(coarse, scale_factor=8, mode="bilinear", align_corners=False)[..., :args.crop, :args.crop]
                         .clamp(0, 255)).div(255.0).to(dev).contiguous()

    def step(timers=None):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        bucket.zero_()
        ev[0].record()
        pred = net(im)
        loss = F.mse_loss(pred, lb)
        ev[1].record()
        loss.backward()
        ev[2].record()
        bucket.all_reduce_mean()
        ev[3].record()
        opt.step()
        sched.step()
        ev[4].record()
        if timers is not None:
            torch.cuda.synchronize()
            for i, k in enumerate(("fwd", "bwd", "allreduce", "adam")):
                timers[k] += ev[i].elapsed_time(ev[i + 1])
        return loss

    timers = {"fwd": 0.0, "bwd": 0.0, "allreduce": 0.0, "adam": 0.0}
    if args.eager:
        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            loss = step(timers)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.steps
    else:
        # the product path: the whole step replayed as one CUDA graph, timed on the device
        from mulut_b200.cli.finetune_lut import GraphedStep
        gs = GraphedStep(net, tuple(im.shape), tuple(lb.shape), 1e-3)
        for _ in range(args.warmup):
            gs(im, lb, 1e-3)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            loss = gs(im, lb, 1e-3)
        e1.record()
        torch.cuda.synchronize()
        dt = e0.elapsed_time(e1) * 1e-3 / args.steps
        loss = loss.clone()
        del gs                                   # drop the captured graph (it holds NCCL work) before the group goes away
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res = {"workload": "cfg4 finetune step, global batch {} of {}x{} patches ({}), x4 sdy 2-stage".format(
        args.batch, args.crop, args.crop, "smooth" if args.smooth else "uniform noise"),
           "n_gpus": world, "ms_per_step": float(t.item()) * 1e3, "patches_per_s": args.batch / float(t.item()),
           "mode": "eager" if args.eager else "cuda graph",
           "breakdown_ms": {k: v / args.steps for k, v in timers.items()} if args.eager else None, "loss": float(loss.item()),
           "allreduce_bytes": bucket.flat.numel() * 4,
           # 60 + 960 fp32 scatter-adds per patch pixel in the backward (issued as 60 scalar + 240 four-wide reds)
           "lut_grad_scatter_adds_per_step": args.batch * args.crop * args.crop * 1020,
           "scatter_adds_per_s": args.batch * args.crop * args.crop * 1020 / float(t.item())}

    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)          # NCCL communicators that were captured in a CUDA graph can hang in their destructor


if __name__ == "__main__":
    main()
