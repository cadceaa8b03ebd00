#!/usr/bin/env python
"""Phase timing of the binned kernel K1f (instrumented side build, -DMULUT_BN_TIMING).

    python tools/bn_timing.py --build            # here (no GPU): builds mulut_b200/libmulut_b200_timing.so
    python tools/bn_timing.py [--config cfg2|x2s1] [--frames 16]     # on the GPU box

Prints, averaged over CTAs, the SM cycles thread 0 spent per phase and per tile visit:
calibrates the plan kernel's cost model (BN_CV / BN_CS in csrc/infer_binned.cu).
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "mulut_b200", "libmulut_b200_timing.so")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--build", action="store_true")
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--data", default="uniform", choices=["uniform", "natural"])
    args = ap.parse_args()
    if args.build:
        from mulut_b200 import build
        print(build.build(defines=["MULUT_BN_TIMING"], lib_path=LIB))
        return
    os.environ["MULUT_B200_LIB"] = LIB
    import torch
    import bench
    from mulut_b200 import _lib
    from mulut_b200.infer import LutEngine
    bench.DATA = args.data
    bench.select_config(args.config)
    luts = bench.make_luts()
    eng = LutEngine(luts, bench.STAGES, bench.MODES, bench.SCALE, bench.INTERVAL, device=0, kernel=_lib.KERNEL_TILED_BINNED)
    d_in = torch.from_numpy(bench.make_frames(args.frames, 1000)).cuda()
    for _ in range(3):
        out = eng.infer_device(d_in)
    torch.cuda.synchronize()
    res = (ctypes.c_double * 10)()
    fn = _lib.lib().mulut_debug_bn_timing
    fn.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.c_int]
    assert fn(res, 148) == 0
    wait, fixup, scan, barrier, rounds, n_rounds, entries, visits, t_max, t_mean = list(res)
    total = wait + fixup + scan + barrier + rounds
    raw = (ctypes.c_ulonglong * 2048)()
    fr = _lib.lib().mulut_debug_bn_timing_raw
    fr.argtypes = [ctypes.POINTER(ctypes.c_ulonglong)]
    assert fr(raw) == 0
    bins = {}
    for c in range(148):
        r = list(raw[c * 8:c * 8 + 8])
        b = r[7] >> 48
        d = bins.setdefault(b, {"ctas": 0, "cycles": 0, "max_cycles": 0, "entries": 0, "visits": 0, "rounds": 0})
        t = sum(r[:5])
        d["ctas"] += 1; d["cycles"] += t; d["max_cycles"] = max(d["max_cycles"], t)
        d["entries"] += r[6]; d["visits"] += r[7] & ((1 << 48) - 1); d["rounds"] += r[5]
    for b in sorted(bins):
        d = bins[b]
        print("bin", b, "ctas", d["ctas"], "mean_cycles %.0f" % (d["cycles"] / d["ctas"]), "max", d["max_cycles"],
              "entries", d["entries"], "visits", d["visits"], "rounds", d["rounds"], file=sys.stderr)
    print(json.dumps({
        "config": args.config, "data": args.data, "frames": args.frames, "visits_per_cta": visits, "rounds_per_cta": n_rounds,
        "entries_per_cta": entries, "cycles_per_cta": total, "slowest_cta_cycles": t_max,
        "load_balance": t_mean / t_max if t_max else None,
        "per_visit": {"tma_wait": wait / visits, "border_patch": fixup / visits, "scan": scan / visits,
                      "barrier": barrier / visits, "rounds": rounds / visits},
        "cycles_per_entry_in_rounds": rounds / max(entries, 1),
        "overhead_cycles_per_visit": (total - rounds) / visits}))


if __name__ == "__main__":
    main()
