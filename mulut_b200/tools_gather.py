"""`python -m mulut_b200.tools_gather` — prints the gather micro-benchmark table
(the measured L1/L2/shared-memory gather rates behind DESIGN.md's layout
choices and the gather-roofline denominator)."""
from __future__ import annotations

import ctypes
import json
import sys

from . import _lib

CASES = [
    # (variant, table bytes, note)
    ("lds_u8", 83584, "stage-1 LUT in shared memory"),
    ("lds_u32", 83584, ""),
    ("ldg_u8", 83584, "stage-1 LUT, one mode, via L1/L2"),
    ("ldg_u8", 250752, "stage-1 LUTs, three modes"),
    ("ldg_u32", 334084, "x2 last-stage LUT, one mode, vertex-major"),
    ("ldg_u32", 1002252, "x2 last-stage LUTs, three modes, vertex-major"),
    ("ldg_u128", 4009008, "x4 last-stage LUTs, three modes, vertex-major (16 B rows)"),
    ("ldg_u128", 3 << 20, "cell-major stage-1 (16 B cells), three modes"),
    ("quad_cell64", 4 << 20, "x2 cell-major, one mode"),
    ("quad_cell64", 12 << 20, "x2 cell-major, three modes"),
    ("pair_cell64", 12 << 20, "x2 cell-major, 2 lanes x 256-bit loads"),
    ("oct_cell128", 24 << 20, "128-B lines, 8 lanes"),
    ("oct_cell128", 50 << 20, "x4 half-cells (128 B), three modes"),
    ("cpasync_cell64", 12 << 20, "x2 cell-major staged through smem with cp.async + 5 LDS"),
    ("quad_cell256_3rows", 50 << 20, "x4 cell-major, three modes: K1e's three 64-B row-blocks by 4 lanes x LDG.128"),
    ("quad_cell256_4sect", 50 << 20, "x4 sector layout, one LDG.256 per lane"),
    ("bulk_cell256", 50 << 20, "x4 cell-major: one cp.async.bulk (TMA unit, UBLKCP) of the 256-B cell per lane -> smem ring -> LDS"),
    ("bulk_rows64x3", 50 << 20, "x4 cell-major: three 64-B cp.async.bulk per lane -> smem ring -> LDS"),
    ("quad_cell64", 200 << 20, "beyond L2 (HBM gathers)"),
]


def run(device=0, iters=256, configs=((2, 512), (4, 256), (1, 1024))):
    L = _lib.lib()
    out = (ctypes.c_double * 3)()
    rows = []
    for name, table, note in CASES:
        best = None
        for bps, tpb in configs:
            if name == "cpasync_cell64" and tpb > 512:
                continue
            if name.startswith("bulk_") and tpb > 256:           # 16 KB of ring per warp
                continue
            rc = L.mulut_gather_bench(device, _lib.GB_VARIANTS[name], table, iters, bps, tpb, 3, out)
            if rc != 0:
                rows.append({"variant": name, "table_bytes": table, "error": _lib.last_error()})
                continue
            r = {"variant": name, "table_bytes": table, "blocks_per_sm": bps, "threads": tpb,
                 "Ggathers_per_s": out[0] / 1e9, "useful_GBps": out[1] / 1e9, "ms": out[2] * 1e3, "note": note}
            if best is None or r["Ggathers_per_s"] > best["Ggathers_per_s"]:
                best = r
        if best:
            rows.append(best)
    return rows


if __name__ == "__main__":
    rows = run()
    for r in rows:
        print(json.dumps(r))
