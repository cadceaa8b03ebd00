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

    def __init__(self, params: List[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(n, dtype=ref.dtype, device=ref.device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_(self) -> None:
        self.flat.zero_()
        # re-attach in case an optimizer dropped the views (set_to_none=True)
        off = 0
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != self.flat[off:off + p.numel()].data_ptr():
                p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def all_reduce_mean(self) -> None:
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(dist.get_world_size())
