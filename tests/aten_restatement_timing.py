"""Not a test: times the plain-ATen restatement of the reference's torch path (the fp32 oracle,
oracle/interp_torch_oracle.py: MuLUT.forward + MSE + backward) on the GPU - the bar K4 replaces.

    python tests/aten_restatement_timing.py [--batch 32]

Round 1 on one B200: 37 ms for a batch of 32 patches of 48x48 (x4 sdy 2-stage), against 2.2 ms for a
batch of 256 through the fused kernels (tools/finetune_bench.py)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--crop", type=int, default=48)
    args = ap.parse_args()
    from oracle import interp_torch_oracle as TO
    d = os.path.join(ROOT, "tests", "golden", "luts_x4")
    luts = {"s{}_{}".format(s, m): np.load(os.path.join(d, "LUT_ft_x4_4bit_int8_s{}_{}.npy".format(s, m))).reshape(
        -1, 1 if s == 1 else 16) for s in (1, 2) for m in "sdy"}
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(0)
    im = (torch.randint(0, 256, (args.batch, 1, args.crop, args.crop), generator=g).float() / 255.0).to(dev)
    lb = (torch.randint(0, 256, (args.batch, 1, args.crop * 4, args.crop * 4), generator=g).float() / 255.0).to(dev)
    ws = {k: torch.tensor(v.astype(np.float32) / 127.0, device=dev, requires_grad=True) for k, v in luts.items()}
    for _ in range(2):
        F.mse_loss(TO.mulut_forward(ws, im, 2, "sdy", 4), lb).backward()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 3
    for _ in range(n):
        F.mse_loss(TO.mulut_forward(ws, im, 2, "sdy", 4), lb).backward()
    torch.cuda.synchronize()
    print(json.dumps({"batch": args.batch, "ms_fwd_bwd": (time.perf_counter() - t0) / n * 1e3,
                      "what": "sorted-simplex ATen restatement of sr/model.py (oracle), same GPU"}))


if __name__ == "__main__":
    main()
