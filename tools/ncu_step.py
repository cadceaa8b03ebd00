#!/usr/bin/env python
"""One step of a bench config for `ncu` captures: 2 warm-up steps, then ONE timed-shape step.

    python tools/ncu_step.py [--config cfg2] [--frames 16]
    ncu --set full --clock-control none --import-source on -k regex:'stage_|combine_kernel' -s <2 x kernels per step> \\
        -c <kernels per step> -o gpurun_out/prof python tools/ncu_step.py

Kernels per step: cfg2 default 4 (K1h, K1b, K1f, orphan list); with MULUT_K1_FUSED=1 3 (K1i, K1f, orphan list);
cfg3 3 (K1h, K1b, K1e).
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--frames", type=int, default=0)
    args = ap.parse_args()
    import torch
    import bench
    from mulut_b200.infer import LutEngine
    frames = bench.select_config(args.config)
    if args.frames > 0:
        frames = args.frames
    luts = bench.make_luts()
    eng = LutEngine(luts, bench.STAGES, bench.MODES, bench.SCALE, bench.INTERVAL, device=0)
    d = torch.from_numpy(bench.make_frames(frames, 1000)).cuda()
    out = torch.empty((frames, bench.H * bench.SCALE, bench.W * bench.SCALE, bench.C), dtype=torch.uint8, device="cuda")
    eng.reserve(frames, bench.H, bench.W, bench.C)
    for _ in range(3):
        eng.infer_device(d, out)
    torch.cuda.synchronize()
    print("ok", int(out[0, :4, :4].sum()))


if __name__ == "__main__":
    main()
