#!/usr/bin/env python
"""Per-kernel CUDA-event times of one config's step (device-resident), for A/B runs of side builds:

    [MULUT_B200_LIB=mulut_b200/libmulut_b200_exp.so] python tests/k1_timing.py [--config cfg2] [--steps 10] [--check]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--data", default="uniform")
    ap.add_argument("--check", action="store_true", help="compare frame 0 with the C oracle")
    args = ap.parse_args()
    import numpy as np
    import torch
    import bench
    from mulut_b200.infer import LutEngine
    bench.DATA = args.data
    frames = bench.select_config(args.config)
    luts = bench.make_luts()
    eng = LutEngine(luts, bench.STAGES, bench.MODES, bench.SCALE, bench.INTERVAL, device=0)
    host = bench.make_frames(frames, 1000)
    d = torch.from_numpy(host).cuda()
    out = torch.empty((frames, bench.H * bench.SCALE, bench.W * bench.SCALE, bench.C), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        eng.infer_device(d, out)
    torch.cuda.synchronize()
    eng.profile(True)
    for _ in range(args.steps):
        eng.infer_device(d, out)
    p = eng.profile_read()
    res = {k: round(v[0] / v[1], 4) for k, v in p.items()}
    res["sum_ms"] = round(sum(v[0] / v[1] for v in p.values()), 4)
    if args.check:
        from oracle import c_oracle as CO
        ref = CO.sr_u8(host[:1], luts, bench.STAGES, bench.MODES, bench.SCALE, bench.INTERVAL)
        res["bit_exact"] = bool(np.array_equal(out[:1].cpu().numpy(), ref))
    print(os.environ.get("MULUT_B200_LIB", "default"), args.config, args.data, res)


if __name__ == "__main__":
    main()
