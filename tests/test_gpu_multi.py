"""Multi-process GPU tests (one process per rank): frame-sharded inference is
byte-identical to a single-GPU run; data-parallel finetune gradients after the one
flat all-reduce equal the single-process gradients of the concatenated batch.
NCCL over distinct GPUs when the box has >= 2, otherwise gloo with both ranks on
cuda:0 (no kernel of one rank ever waits on the other)."""
import os
import socket

import numpy as np
import pytest

from conftest import norm_max_err
from oracle import mulut_oracle as O

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, backend, q):
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    from mulut_b200.dist import FlatGradBucket, shard_range
    from mulut_b200.infer import LutEngine
    from mulut_b200.model import MuLUT
    dev_index = rank if backend == "nccl" else 0
    torch.cuda.set_device(dev_index)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group(backend, rank=rank, world_size=world)
    dev = torch.device("cuda", dev_index)
    # ---- inference: contiguous frame shards, no collective ----
    luts = O.random_luts(1, 2, "sdy", 2)
    frames = np.random.default_rng(0).integers(0, 256, (5, 40, 56, 3), dtype=np.uint8)
    b, e = shard_range(len(frames), rank, world)
    with LutEngine(luts, 2, "sdy", 2, 4, device=dev_index) as eng:
        out = eng(torch.from_numpy(frames[b:e]).to(dev)).cpu().numpy()
    # ---- finetune: data parallel over patches ----
    luts4 = O.random_luts(2, 2, "sdy", 4)
    net = MuLUT(None, 2, ["s", "d", "y"], upscale=4, interval=4, luts=luts4).to(dev)
    bucket = FlatGradBucket(list(net.parameters()))
    rng = np.random.default_rng(3)
    im = (rng.integers(0, 256, (4, 1, 12, 12)) / 255.0).astype(np.float32)
    lb = (rng.integers(0, 256, (4, 1, 48, 48)) / 255.0).astype(np.float32)
    pb, pe = shard_range(4, rank, world)
    bucket.zero_()
    loss = F.mse_loss(net(torch.tensor(im[pb:pe], device=dev)), torch.tensor(lb[pb:pe], device=dev))
    loss.backward()
    bucket.all_reduce_mean()
    torch.cuda.synchronize()
    plain = bucket.packed().cpu().numpy()
    # ---- the same step the way GraphedStep runs it: fused loss head, K4 accumulating into the bucket, the last
    # stage's tables reduced from the backward hook while stage 1 still runs ----
    begun = []
    ranges = {st: bucket.range_of(net.stage_parameters(st)) for st in (1, 2)}
    net.accumulate_grads_into(True)
    net.on_stage_grads(lambda st: (begun.append(st), bucket.begin_range(*ranges[st])))
    bucket.zero_()
    net.forward_loss(torch.tensor(im[pb:pe], device=dev), torch.tensor(lb[pb:pe], device=dev)).backward()
    n_pending = len(bucket._pending)
    bucket.all_reduce_mean()
    torch.cuda.synchronize()
    assert begun == [2] and n_pending == 1
    assert norm_max_err(bucket.packed().cpu().numpy(), plain) < 1e-5
    net.on_stage_grads(None)
    net.accumulate_grads_into(False)
    if backend == "nccl":
        # and captured: two replays of the graph on the same batch equal two eager steps
        from mulut_b200.cli.finetune_lut import GraphedStep
        import copy
        net2 = copy.deepcopy(net)
        os.environ["MULUT_AR_OVERLAP"] = "1"
        gs = GraphedStep(net2, (pe - pb, 1, 12, 12), (pe - pb, 1, 48, 48), 1e-3)
        assert net2._stage_grads_cb is not None
        ims, lbs = torch.tensor(im[pb:pe], device=dev), torch.tensor(lb[pb:pe], device=dev)
        l0 = float(gs(ims, lbs, 1e-3).item())
        gs(ims, lbs, 1e-3)
        torch.cuda.synchronize()
        os.environ["MULUT_AR_OVERLAP"] = "0"
        net3 = copy.deepcopy(net)
        gs3 = GraphedStep(net3, (pe - pb, 1, 12, 12), (pe - pb, 1, 48, 48), 1e-3)
        assert net3._stage_grads_cb is None
        l3 = float(gs3(ims, lbs, 1e-3).item())
        gs3(ims, lbs, 1e-3)
        torch.cuda.synchronize()
        assert abs(l0 - l3) < 1e-6 * max(1.0, abs(l3))
        for pa, pb_ in zip(net2.parameters(), net3.parameters()):
            assert norm_max_err(pa.detach().cpu().numpy(), pb_.detach().cpu().numpy()) < 1e-5
        gs.close(); gs3.close()
        del gs, gs3, net2, net3
    q.put((rank, (b, e), out, plain))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_inference_and_finetune_match_single_process():
    import torch
    import torch.multiprocessing as mp
    import torch.nn.functional as F
    from mulut_b200.dist import FlatGradBucket
    from mulut_b200.infer import LutEngine
    from mulut_b200.model import MuLUT
    world = 2
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, backend, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)

    luts = O.random_luts(1, 2, "sdy", 2)
    frames = np.random.default_rng(0).integers(0, 256, (5, 40, 56, 3), dtype=np.uint8)
    with LutEngine(luts, 2, "sdy", 2, 4, device=0) as eng:
        single = eng(torch.from_numpy(frames).cuda()).cpu().numpy()
    sharded = np.concatenate([r[2] for r in res])
    assert [r[1] for r in res] == [(0, 3), (3, 5)]
    assert sharded.shape == single.shape and (sharded == single).all()

    luts4 = O.random_luts(2, 2, "sdy", 4)
    net = MuLUT(None, 2, ["s", "d", "y"], upscale=4, interval=4, luts=luts4).cuda()
    bucket = FlatGradBucket(list(net.parameters()))
    rng = np.random.default_rng(3)
    im = (rng.integers(0, 256, (4, 1, 12, 12)) / 255.0).astype(np.float32)
    lb = (rng.integers(0, 256, (4, 1, 48, 48)) / 255.0).astype(np.float32)
    F.mse_loss(net(torch.tensor(im).cuda()), torch.tensor(lb).cuda()).backward()
    full = bucket.packed().cpu().numpy()
    assert np.allclose(res[0][3], res[1][3])
    # equal shards: mean of per-rank mean losses == global mean loss
    assert norm_max_err(res[0][3], full) < 1e-5
