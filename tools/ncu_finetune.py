#!/usr/bin/env python
"""Three eager cfg4 finetune steps for `ncu` captures, launched the way `GraphedStep` launches them (zero the gradient
bucket, K4 forward x1 / x4, fused loss head, K4 backward x4 / x1 accumulating into the bucket, fused Adam):

    ncu --set full --clock-control none --import-source on -k regex:'stage_(fwd|bwd)' -s 8 -c 4 \\
        -o gpurun_out/prof_k4 python tools/ncu_finetune.py [--smooth]
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ft.csv \\
        python tools/ncu_finetune.py                       # launch list: every kernel of the three steps

`--unfused-head` keeps `F.mse_loss(net(im), lb)` and autograd's own accumulation (the step before the loss head).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.nn.functional as F
    from mulut_b200.cli.finetune_lut import synthetic_batch
    from mulut_b200.model import MuLUT
    from tools.finetune_bench import shipped_luts
    dev = torch.device("cuda", 0)
    net = MuLUT(None, 2, ["s", "d", "y"], upscale=4, interval=4, luts=shipped_luts()).to(dev)
    im, lb = synthetic_batch(256, 48, 4, 1000, dev)
    if "--smooth" in sys.argv:
        g = torch.Generator(device="cpu").manual_seed(7)
        coarse = torch.randint(0, 256, (256, 1, 8, 8), generator=g).float()
        im = torch.round(F.interpolate(coarse, scale_factor=8, mode="bilinear", align_corners=False)[..., :48, :48]
                         .clamp(0, 255)).div(255.0).to(dev).contiguous()
    if "--unfused-head" in sys.argv:
        for _ in range(3):
            net.zero_grad(set_to_none=False)
            loss = F.mse_loss(net(im), lb)
            loss.backward()
    else:
        from mulut_b200 import dist as mdist
        bucket = mdist.FlatGradBucket(list(net.parameters()))
        opt = mdist.FusedAdam(bucket, 1e-4)
        net.accumulate_grads_into(True)
        for _ in range(3):
            bucket.zero_()
            loss = net.forward_loss(im, lb)
            loss.backward()
            opt.step()
    torch.cuda.synchronize()
    print("ok", float(loss))


if __name__ == "__main__":
    main()
