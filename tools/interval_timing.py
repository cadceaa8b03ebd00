#!/usr/bin/env python
"""Device-resident throughput of configurations no shipped model uses (other --interval values, scale 3):
8 x 1080p frames per launch, per-kernel CUDA-event times.  The numbers of DESIGN.md section 6 ("Other intervals and
scale 3") come from here.

    python tools/interval_timing.py
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mulut_b200.infer import LutEngine
rng = np.random.default_rng(0)
frames = rng.integers(0, 256, (8, 1080, 1920, 3), dtype=np.uint8)
d = torch.from_numpy(frames).cuda()
for interval, scale in [(4, 2), (5, 2), (6, 2), (5, 3), (4, 3), (5, 4), (6, 4), (3, 2)]:
    L = (1 << (8 - interval)) + 1
    lrng = np.random.default_rng(1)
    luts = {"s{}_{}".format(st + 1, m): lrng.integers(-127, 128, (L ** 4, scale * scale if st == 1 else 1), dtype=np.int8)
            for st in range(2) for m in "sdy"}
    with LutEngine(luts, 2, "sdy", scale, interval, device=0) as eng:
        out = eng(d)
        torch.cuda.synchronize()
        eng.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            eng.infer_device(d, out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        p = eng.profile_read()
        print("interval", interval, "scale", scale, "ms/8 frames %.3f" % ms, "Gpix/s %.2f" % (8 * 1080 * 1920 * scale * scale / ms / 1e6),
              {k: round(v[0] / v[1], 3) for k, v in p.items()}, flush=True)
