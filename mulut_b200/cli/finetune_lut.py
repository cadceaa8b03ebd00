"""`python -m mulut_b200.cli.finetune_lut ...` (or under torchrun, one rank per GPU)

Drop-in for the reference's step 3 (sr/3_finetune_lut.py:68-172): same flags,
Adam(lr0, betas .9/.999, eps 1e-8) with the cosine LambdaLR lr0 -> lr1
(:85-95), MSE loss, final LUT export (:162-169).  Differences: the interpolation
runs in the sm_100a kernels; multi-GPU is real data parallelism (one NCCL
all-reduce of the flat 17 MB LUT-gradient buffer per step) instead of the
reference's non-functional `--gpuNum` branch (:156-157); `--synthetic` trains on
seeded random patches because the DIV2K set is not shipped (any iterator of (im, lb) batches can be
passed to `finetune_steps(batches=...)`); `--displayStep` prints the running loss (:142-149),
`--valStep` runs `valid_steps` (:23-65, whole-image C=3 inference through the torch path) when the
benchmark folders exist under `--valDir`, and `--saveStep` additionally exports the LUTs and the
optimiser state every N steps so a long run can resume with `--startIter` (the reference saves once,
at the end).
"""
from __future__ import annotations

import math
import os
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


def load_benchmarks(val_dir: str, scale: int, datasets=("Set5", "Set14", "B100", "Urban100", "Manga109")):
    """{dataset: [(name, lr uint8 HWC, hr uint8 HWC)]} for the benchmark folders that exist under `val_dir`
    (layout of sr/data.py:127-168: <dataset>/HR/*.png and <dataset>/LR_bicubic/X<scale>/*.png)."""
    import os
    from PIL import Image
    from ..metrics import modcrop
    out = {}
    for ds in datasets:
        folder = os.path.join(val_dir, ds, "HR")
        if not os.path.isdir(folder):
            continue
        items = []
        for f in sorted(os.listdir(folder)):
            hr = modcrop(np.array(Image.open(os.path.join(folder, f))), scale)
            lr = np.array(Image.open(os.path.join(val_dir, ds, "LR_bicubic/X%d" % scale, f)))
            if hr.ndim == 2:
                hr = np.stack([hr] * 3, axis=2)
            if lr.ndim == 2:
                lr = np.stack([lr] * 3, axis=2)
            items.append((f[:-4], np.ascontiguousarray(lr[:, :, :3]), np.ascontiguousarray(hr[:, :, :3])))
        out[ds] = items
    return out


def valid_steps(model_G, benchmarks, scale: int, it: int, out_dir=None, log=print):
    """sr/3_finetune_lut.py:23-65: whole images ([1,3,H,W], any non-square size) through MuLUT.forward,
    round(clip(pred*255)) -> uint8, PSNR / SSIM on the BT.601 luma (on the device:
    mulut_eval_psnr_ssim_y_u8).  Returns {dataset: (mean PSNR, mean SSIM)}."""
    import os
    from ..metrics import psnr_ssim_device
    device = next(model_G.parameters()).device
    res = {}
    was_training = model_G.training
    with torch.no_grad():
        model_G.eval()
        for ds, items in benchmarks.items():
            psnrs, ssims = [], []
            for name, lr, hr in items:
                im = torch.from_numpy(lr.astype(np.float32) / 255.0).permute(2, 0, 1)[None].to(device)
                pred = model_G(im) * 255.0
                pred = torch.round(torch.clamp(pred[0].permute(1, 2, 0), 0, 255)).to(torch.uint8).contiguous()
                p, s = psnr_ssim_device(pred, torch.from_numpy(hr).to(device), scale)
                psnrs.append(p)
                ssims.append(s)
                if out_dir is not None:
                    from PIL import Image
                    os.makedirs(os.path.join(out_dir, ds), exist_ok=True)
                    Image.fromarray(pred.cpu().numpy()).save(os.path.join(out_dir, ds, "{}_lutft.png".format(name)))
            res[ds] = (float(np.mean(psnrs)), float(np.mean(ssims)))
            log("Iter {} | Dataset {} | AVG PSNR: {:02f}, AVG: SSIM: {:04f}".format(it, ds, res[ds][0], res[ds][1]))
    model_G.train(was_training)
    return res


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
        if getattr(model_G, "fused", False):
            model_G.accumulate_grads_into(True)      # K4's backward adds straight into the bucket's views
            if mdist.env_rank_world()[1] > 1 and os.environ.get("MULUT_AR_OVERLAP", "0") == "1":
                # opt-in: a stage's tables are consecutive in the bucket, so the last stage's 16 MB can be reduced while
                # stage 1 still runs its backward.  Measured: 0.844 -> 0.830 ms on 2 GPUs, no
                # change on 4, 0.400 -> 0.425 ms on 8 (the second, 1 MB all-reduce pays a full collective latency
                # after the backward) - hence off unless asked for
                ranges = {st: self.bucket.range_of(model_G.stage_parameters(st)) for st in range(1, model_G.stages + 1)}
                bucket = self.bucket                 # (not `self`: the model must not keep the captured graph alive)
                model_G.on_stage_grads(lambda st: bucket.begin_range(*ranges[st]))
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

    def close(self):
        """Detach from the model and drop the captured graph (it holds NCCL work: it must go before the process group)."""
        model, self.model = getattr(self, "model", None), None
        if model is not None and getattr(model, "_stage_grads_cb", None) is not None:
            model.on_stage_grads(None)
        self.graph = None

    def __del__(self):
        self.close()

    def _body(self):
        self.bucket.zero_()
        # pred = model(im); loss = F.mse_loss(pred, lb)  (3_finetune_lut.py:130-132) - with K4 the `/ 255` and the loss
        # are one kernel per direction (MuLUT.forward_loss)
        loss = self.model.forward_loss(self.im, self.lb) if hasattr(self.model, "forward_loss") else \
            F.mse_loss(self.model(self.im), self.lb)
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
                   total_iter: int, weight_decay: float = 0.0, batches=None, graph: bool = False,
                   start_iter: int = 0, display_step: int = 0, on_step=None, log=print, state=None):
    """The training loop body of 3_finetune_lut.py:118-136 on this rank's share.
    graph=True replays the step as one CUDA graph (same arithmetic; needs a fixed batch shape).
    batches: None (seeded synthetic patches) or any iterable of (im, lb) CUDA batches - a list, a generator,
    a DataLoader wrapper.  display_step: every N steps the mean loss of the last N is synchronised, logged
    and the stored device scalars are dropped.  on_step(i, optimiser): called after step i (1-based, like
    the reference's loop) - the CLI hangs validation and checkpoints on it.  state: optimiser state to resume
    from (see `checkpoint`).  Returns the per-step losses as floats."""
    rank, world, _ = mdist.env_rank_world()
    device = next(model_G.parameters()).device
    r = model_G.upscale
    it = iter(batches) if batches is not None else None
    first = next(it) if it is not None else None
    lam = lr_lambda(total_iter, lr0, lr1)
    model_G.train()
    losses, pending = [], []

    def flush(i, show=True):
        if pending:
            vals = [float(v) for v in torch.stack(pending).cpu()]
            losses.extend(vals)
            pending.clear()
            if show and display_step and rank == 0:
                log("Iter:{:6d}, loss:{:.3e}".format(i, float(np.mean(vals))))

    def batch_of(i):
        nonlocal first
        if it is None:
            return synthetic_batch(batch, crop, r, seed + i * world + rank, device)
        if first is not None:
            b, first = first, None
            return b
        return next(it)

    if graph:
        C = 1 if first is None else first[0].shape[1]
        b0 = batch if first is None else first[0].shape[0]
        hw = (crop, crop) if first is None else tuple(first[0].shape[2:])
        gs = GraphedStep(model_G, (b0, C) + hw, (b0, C, hw[0] * r, hw[1] * r), lr0, weight_decay)
        if state is not None:
            gs.opt.load_state(state)
        for k in range(steps):
            i = start_iter + k
            im, lb = batch_of(i)
            pending.append(gs(im, lb, lr0 * lam(i)))       # LambdaLR: lr of step i is lr0 * lambda(i)
            if display_step and (i + 1) % display_step == 0:
                flush(i + 1)
            elif len(pending) >= 1024:                      # never hold more than ~1k device scalars
                flush(i + 1, show=False)
            if on_step is not None:
                on_step(i + 1, gs.opt)
        flush(start_iter + steps)
        gs.close()                                          # the captured graph (it holds NCCL work) goes before the group
        del gs
        return losses
    params = [p for p in model_G.parameters() if p.requires_grad]
    bucket = mdist.FlatGradBucket(params)
    opt_G = torch.optim.Adam(params, lr=lr0, betas=(0.9, 0.999), eps=1e-8, weight_decay=weight_decay)
    sched = torch.optim.lr_scheduler.LambdaLR(opt_G, lr_lambda=lambda x: lam(x + start_iter))
    for k in range(steps):
        i = start_iter + k
        im, lb = batch_of(i)
        bucket.zero_()
        pred = model_G(im)
        loss = F.mse_loss(pred, lb)
        loss.backward()
        bucket.all_reduce_mean()
        opt_G.step()
        sched.step()
        pending.append(loss.detach())
        if display_step and (i + 1) % display_step == 0:
            flush(i + 1)
        elif len(pending) >= 1024:
            flush(i + 1, show=False)
        if on_step is not None:
            on_step(i + 1, None)
    flush(start_iter + steps)
    return losses


def checkpoint(model_G: MuLUT, opt, exp_dir: str, it: int) -> None:
    """LUT export (the reference's .npy format) + parameters and Adam moments for --startIter."""
    import os
    model_G.export_luts(exp_dir)
    blob = {"iter": it, "params": {n: p.detach().cpu() for n, p in model_G.named_parameters()}}
    if opt is not None:
        blob["adam"] = opt.state()
    torch.save(blob, os.path.join(exp_dir, "Finetune_{:06d}.pth".format(it)))


def main(argv=None):
    import os
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
    state = None
    if opt.startIter > 0:                     # resume: parameters + Adam moments written by --saveStep
        blob = torch.load(os.path.join(opt.expDir, "Finetune_{:06d}.pth".format(opt.startIter)), map_location=device)
        with torch.no_grad():
            for n, p in model_G.named_parameters():
                p.copy_(blob["params"][n])
        state = blob.get("adam")
    benchmarks = load_benchmarks(opt.valDir, opt.scale) if (rank == 0 and opt.valStep > 0 and os.path.isdir(opt.valDir)) else {}

    def on_step(i, optimiser):
        if rank != 0:
            return
        if benchmarks and (i % opt.valStep == 0 or i == 1):           # 3_finetune_lut.py:152-157
            valid_steps(model_G, benchmarks, opt.scale, i, os.path.join(opt.expDir, "val"))
        if opt.saveStep > 0 and i % opt.saveStep == 0:
            checkpoint(model_G, optimiser, opt.expDir, i)

    per_rank = max(1, opt.batchSize // world)
    st = time.time()
    n_steps = max(0, opt.totalIter - opt.startIter)
    losses = finetune_steps(model_G, n_steps, per_rank, opt.cropSize, 0, opt.lr0, opt.lr1, opt.totalIter,
                            opt.weightDecay, graph=not getattr(opt, "eager", False), start_iter=opt.startIter,
                            display_step=opt.displayStep, on_step=on_step, state=state)
    if rank == 0:
        print("{} | Iter:{:6d}, loss:{:.3e}, rT:{:.4f}".format(opt.expDir, opt.totalIter, np.mean(losses[-100:]) if losses else float("nan"),
                                                          (time.time() - st) / max(1, n_steps)))
        model_G.export_luts(opt.expDir)
        print("Finetuned LUT saved to {}".format(opt.expDir))
    if world > 1:
        # the step graph (which captured the all-reduce) was destroyed inside finetune_steps: the group can go
        torch.cuda.synchronize()
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return losses


if __name__ == "__main__":
    main(sys.argv[1:])
