#!/usr/bin/env python
"""cfg4 benchmark: one LUT fine-tuning step (3_finetune_lut.py:129-136: zero_grad, MuLUT.forward, mse_loss,
backward, LUT-gradient all-reduce, Adam), GLOBAL batch 256 of 48x48 patches split across the ranks, x4 sdy
2-stage, the reference's shipped LUTs.  One process per GPU (torchrun).

    python tools/finetune_bench.py [--batch 256] [--steps 30] [--smooth] [--no-reference]

`finetune_block()` is what bench.py puts into its JSON line under "finetune" at every N:
  * ms_per_step: the product path - the whole step replayed as ONE CUDA graph, CUDA events, max over ranks;
  * breakdown_ms: each phase captured as its own CUDA graph and timed the same way (zero-grad memset,
    forward + MSE, backward, all-reduce, Adam); `sum` vs `ms_per_step` shows what the phases hide of each other;
  * allreduce: the 17.04 MB NCCL all-reduce alone (bus bandwidth);
  * scatter_adds_per_s: SURVEY 8(d)'s 1020 fp32 LUT-gradient scatter-adds per patch pixel;
  * reference_module: the UNMODIFIED reference `model.MuLUT` (oracle/_ref/sr/model.py staged by
    `python -m oracle.fetch_ref`) running the same step on the same GPU through ATen - the bar the kernels
    are measured against (rank 0, N = 1 only: the reference has no multi-GPU path).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def shipped_luts():
    d = os.path.join(ROOT, "tests", "golden", "luts_x4")
    return {"s{}_{}".format(s, m): np.load(os.path.join(d, "LUT_ft_x4_4bit_int8_s{}_{}.npy".format(s, m))).reshape(
        -1, 1 if s == 1 else 16) for s in (1, 2) for m in "sdy"}


def _capture(fn, dev):
    """Warm `fn` up on a side stream, then capture it into a CUDA graph."""
    import torch
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(2):
            fn()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn()
    return g, out


def _time_graph(g, steps, warmup, world, dist, dev):
    """ms per replay: CUDA events on the launching stream, barrier on both sides, max over ranks."""
    import torch
    for _ in range(warmup):
        g.replay()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        g.replay()
    e1.record()
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def finetune_block(rank, world, local, dist=None, steps=30, warmup=10, batch=256, crop=48, smooth=False,
                   reference_fn=None, clock_sampler=None, phases_wanted=True):
    """reference_fn(luts, batch, crop, dev, steps) -> dict: the reference-module leg (bench.py owns it: it runs code
    staged under oracle/_ref, and only bench.py's baseline legs may execute the oracle tree)."""
    import torch
    import torch.nn.functional as F
    from mulut_b200.cli.finetune_lut import GraphedStep, synthetic_batch
    from mulut_b200.model import MuLUT

    dev = torch.device("cuda", local)
    luts = shipped_luts()
    net = MuLUT(None, 2, ["s", "d", "y"], upscale=4, interval=4, luts=luts, fused=True).to(dev)
    per_rank = max(1, batch // world)
    im, lb = synthetic_batch(per_rank, crop, 4, 1000 + rank, dev)
    if smooth:
        g = torch.Generator(device="cpu").manual_seed(7 + rank)
        coarse = torch.randint(0, 256, (per_rank, 1, crop // 8 + 2, crop // 8 + 2), generator=g).float()
        im = torch.round(F.interpolate(coarse, scale_factor=8, mode="bilinear", align_corners=False)[..., :crop, :crop]
                         .clamp(0, 255)).div(255.0).to(dev).contiguous()
    gs = GraphedStep(net, tuple(im.shape), tuple(lb.shape), 1e-3)
    gs.im.copy_(im)
    gs.lb.copy_(lb)
    gs.opt.set_lr(1e-3)
    sampler = clock_sampler() if (clock_sampler and rank == 0) else None
    if sampler:
        sampler.start()
    ms_full = _time_graph(gs.graph, steps, warmup, world, dist, dev)
    loss = float(gs.loss.item())
    clocks = sampler.stop() if sampler else None

    if not phases_wanted:
        return {"workload": "the same step on low-frequency patches (neighbouring pixels share LUT rows, like natural "
                            "images: the LUT-gradient atomics collide on a few hundred hot rows)",
                "ms_per_step": ms_full, "patches_per_s": batch / (ms_full * 1e-3), "loss": loss}, gs
    # ---- the phases, each as its own graph (same buffers, same kernels); the step itself starts the all-reduce of
    # the last stage's tables under the first stage's backward (GraphedStep), the phases are timed one after another ----
    overlapped = gs.model._stage_grads_cb is not None
    gs.model.on_stage_grads(None)
    def p_zero():
        gs.bucket.zero_()

    def p_fwd():
        with torch.no_grad():
            return gs.model.forward_loss(gs.im, gs.lb)

    def p_fwd_bwd():
        gs.bucket.zero_()
        l = gs.model.forward_loss(gs.im, gs.lb)
        l.backward()
        return l.detach()

    def p_allreduce():
        gs.bucket.all_reduce_mean()

    def p_adam():
        gs.opt.step()

    phases = {}
    graphs = []
    for name, fn in (("zero_grad", p_zero), ("forward_mse", p_fwd), ("zero+forward+backward", p_fwd_bwd),
                     ("allreduce", p_allreduce), ("adam", p_adam)):
        if name == "allreduce" and world == 1:
            phases[name] = 0.0
            continue
        g, _ = _capture(fn, dev)
        graphs.append(g)
        phases[name] = _time_graph(g, steps, 3, world, dist, dev)
    bwd = phases["zero+forward+backward"] - phases["zero_grad"] - phases["forward_mse"]
    breakdown = {"zero_grad": phases["zero_grad"], "forward_mse": phases["forward_mse"], "backward": bwd,
                 "allreduce": phases["allreduce"], "adam": phases["adam"]}
    breakdown["sum"] = sum(breakdown.values())
    nbytes = gs.bucket.flat.numel() * 4
    ar = None
    if world > 1:
        # ring/tree-independent "bus bandwidth" of an all-reduce: 2 (N-1)/N x bytes / time
        ar = {"bytes": nbytes, "ms": phases["allreduce"],
              "busbw_GBps": 2.0 * (world - 1) / world * nbytes / (phases["allreduce"] * 1e-3) / 1e9,
              "note": "includes the division by the world size (one elementwise kernel over 17 MB)"}
    n_adds = batch * crop * crop * 1020
    res = {
        "workload": "cfg4: LUT finetune step (3_finetune_lut.py:129-136), GLOBAL batch {} of {}x{} patches ({}), x4 sdy "
                    "2-stage, shipped LUTs".format(batch, crop, crop, "smooth" if smooth else "uniform noise"),
        "n_gpus": world, "per_rank_batch": per_rank, "scaling": "strong (global batch fixed)",
        "ms_per_step": ms_full, "patches_per_s": batch / (ms_full * 1e-3), "mode": "one CUDA graph per step",
        "steps": steps, "warmup": warmup, "breakdown_ms": breakdown,
        "breakdown_how": "each phase captured as its own CUDA graph and replayed {} times between CUDA events; backward = "
                         "(zero+forward+backward) - zero_grad - forward_mse".format(steps),
        "allreduce": ar, "allreduce_bytes": nbytes, "allreduce_overlaps_backward": overlapped,
        "lut_grad_scatter_adds_per_step": n_adds, "scatter_adds_per_s": n_adds / (ms_full * 1e-3),
        "scatter_adds_per_s_backward_only": n_adds / (bwd * 1e-3) if bwd > 0 else None,
        "loss": loss, "clocks": clocks,
    }
    del graphs
    if reference_fn is not None and world == 1 and rank == 0:
        try:
            ref = {"B{}".format(b): reference_fn(luts, b, crop, dev, 3) for b in (32, batch)}
            r = ref["B{}".format(batch)]
            if "ms_per_step" in r:
                ref["speedup_at_B{}".format(batch)] = r["ms_per_step"] / ms_full
            # the same module on the host cores: a bounded sample, batch 16, one step
            ref["cpu_B16"] = reference_fn(luts, 16, crop, torch.device("cpu"), 1)
            res["reference_module"] = ref
        except Exception as e:                                 # the bar is a report, never a reason to lose the line
            res["reference_module"] = {"error": repr(e)[:300]}
    return res, gs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256, help="GLOBAL batch (split across ranks)")
    ap.add_argument("--crop", type=int, default=48)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--smooth", action="store_true", help="low-frequency patches (neighbouring pixels share LUT rows, like natural images) instead of uniform noise")
    ap.add_argument("--no-reference", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    sys.path.insert(0, ROOT)
    from bench import ClockSampler, reference_module_step
    res, gs = finetune_block(rank, world, local, dist, args.steps, args.warmup, args.batch, args.crop, args.smooth,
                             None if args.no_reference else reference_module_step, lambda: ClockSampler(local))
    if rank == 0:
        print(json.dumps(res), flush=True)
    gs.close()
    del gs                                       # the captured graphs hold NCCL work: drop them before the group
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
