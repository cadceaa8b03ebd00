"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs,
gloo in the CPU tests).

* Inference shards FRAMES contiguously across ranks and needs no collective
  (SURVEY.md 8e): every rank holds its own LUT replica; outputs are
  byte-identical to a single-GPU run.
* Fine-tuning is data-parallel: the six LUT gradient tables live in ONE flat
  fp32 buffer (4 259 571 elements = 17.04 MB for x4 sdy 2-stage) so that a
  single all-reduce per step suffices; MSE is a mean, so the sum is divided by
  the world size.
"""
from __future__ import annotations

import os
from typing import List, Tuple

import torch
import torch.distributed as dist


def env_rank_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_process_group(backend: str = None) -> Tuple[int, int, int]:
    """Initialise from torchrun's environment (no-op for a single process)."""
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) share of `n_items` for `rank`: the first
    n_items % world ranks get one extra item."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_rows_with_halo(H: int, rank: int, world: int, halo: int) -> Tuple[int, int, int, int]:
    """Strip sharding of ONE tall frame: returns (row_begin, row_end, load_begin,
    load_end): the rows this rank owns and the rows it must read (halo = 2 rows per
    stage on each side; clamped at the frame borders, where the kernel's own
    coordinate clamp supplies the reference's edge replication)."""
    b, e = shard_range(H, rank, world)
    return b, e, max(0, b - halo), min(H, e + halo)


class FlatGradBucket:
    """Keeps all parameters' gradients as views of one flat buffer so one
    all-reduce covers them (replaces nothing in the reference: its multi-GPU
    finetune is unimplemented, 3_finetune_lut.py:156-157)."""

    ALIGN = 4                                # elements: every view starts on a 16-byte boundary (K4 scatters into the
                                             # views with 16-byte vector reductions when it accumulates in place)

    def __init__(self, params: List[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        self.offsets = []
        n = 0
        for p in self.params:
            self.offsets.append(n)
            n += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        ref = self.params[0]
        self.flat = torch.zeros(n, dtype=ref.dtype, device=ref.device)
        for p, off in zip(self.params, self.offsets):
            p.grad = self.flat[off:off + p.numel()].view_as(p)
        self._pending = []                   # (lo, hi, work) of the ranges begin_range() has in flight

    def range_of(self, params) -> Tuple[int, int]:
        """[lo, hi) of the flat buffer covering `params` - they must be consecutive bucket entries."""
        idx = sorted(next(i for i, q in enumerate(self.params) if q is p) for p in params)
        if idx != list(range(idx[0], idx[0] + len(idx))):
            raise ValueError("range_of: parameters are not consecutive in the bucket")
        last = idx[-1]
        hi = self.offsets[last + 1] if last + 1 < len(self.offsets) else self.flat.numel()
        return self.offsets[idx[0]], hi

    def begin_range(self, lo: int, hi: int) -> None:
        """Start the all-reduce of flat[lo:hi] NOW (its gradients are final) so it runs under the rest of the
        backward pass; `all_reduce_mean` reduces whatever was not begun, waits and divides.  Every rank must
        begin the same ranges in the same order."""
        if dist.is_initialized() and dist.get_world_size() > 1 and hi > lo:
            work = dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True)
            self._pending.append((lo, hi, work))

    def zero_(self) -> None:
        self._pending = []
        self.flat.zero_()
        # re-attach in case an optimizer dropped the views (set_to_none=True)
        for p, off in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != self.flat[off:off + p.numel()].data_ptr():
                p.grad = self.flat[off:off + p.numel()].view_as(p)

    def packed(self) -> torch.Tensor:
        """The gradients concatenated WITHOUT the alignment padding (parameter order)."""
        return torch.cat([self.flat[off:off + p.numel()] for p, off in zip(self.params, self.offsets)])

    def all_reduce_mean(self) -> None:
        if dist.is_initialized() and dist.get_world_size() > 1:
            pos = 0
            for lo, hi, _ in sorted(self._pending, key=lambda t: t[0]) + [(self.flat.numel(), 0, None)]:
                if lo > pos:                                   # a gap nobody began: reduce it here
                    dist.all_reduce(self.flat[pos:lo], op=dist.ReduceOp.SUM)
                pos = max(pos, hi)
            for _, _, work in self._pending:
                work.wait()
            self._pending = []
            self.flat.div_(dist.get_world_size())


class FusedAdam:
    """torch.optim.Adam(params, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay) of
    3_finetune_lut.py:85-87 as ONE kernel over flat buffers (mulut_adam_step_f32): the parameters
    are re-pointed at slices of one flat tensor (their values and names are unchanged), the
    gradients are the FlatGradBucket's buffer, the two moment buffers are flat as well.  The
    learning rate and the step counter are device scalars, so the step is CUDA-graph capturable."""

    def __init__(self, bucket: "FlatGradBucket", lr: float, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0):
        self.bucket, self.betas, self.eps, self.weight_decay = bucket, betas, float(eps), float(weight_decay)
        ref = bucket.flat
        if not ref.is_cuda:
            raise RuntimeError("FusedAdam needs CUDA parameters (no CPU fallback)")
        self.flat_param = torch.zeros_like(ref)          # same layout as the gradient bucket (padding stays zero)
        with torch.no_grad():
            for p, off in zip(bucket.params, bucket.offsets):
                n = p.numel()
                self.flat_param[off:off + n].copy_(p.detach().reshape(-1))
                p.data = self.flat_param[off:off + n].view_as(p)
        self.exp_avg = torch.zeros_like(ref)
        self.exp_avg_sq = torch.zeros_like(ref)
        self.lr = torch.tensor(float(lr), device=ref.device)
        self.step_count = torch.zeros((), device=ref.device)

    def set_lr(self, lr: float) -> None:
        self.lr.fill_(float(lr))

    def step(self) -> None:
        import ctypes
        from . import _lib
        dev = self.flat_param.device
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().mulut_adam_step_f32(
                self.flat_param.data_ptr(), self.bucket.flat.data_ptr(), self.exp_avg.data_ptr(),
                self.exp_avg_sq.data_ptr(), self.flat_param.numel(), self.lr.data_ptr(), self.betas[0], self.betas[1],
                self.eps, self.weight_decay, self.step_count.data_ptr(), stream))

    def state(self) -> dict:
        """Moments, step counter and learning rate (host copies) - what a resumed run needs."""
        return {"exp_avg": self.exp_avg.detach().cpu(), "exp_avg_sq": self.exp_avg_sq.detach().cpu(),
                "step": float(self.step_count.item()), "lr": float(self.lr.item())}

    def load_state(self, st: dict) -> None:
        self.exp_avg.copy_(st["exp_avg"]); self.exp_avg_sq.copy_(st["exp_avg_sq"])
        self.step_count.fill_(float(st["step"])); self.lr.fill_(float(st["lr"]))

    def reset_state(self) -> None:
        self.exp_avg.zero_(); self.exp_avg_sq.zero_(); self.step_count.zero_()
