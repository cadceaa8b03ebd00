#!/usr/bin/env python
"""Aggregate an ncu SASS source page (ncu -i X.ncu-rep --page source --csv --print-source sass)
into address regions: executed instructions, stall samples and shared-memory wavefronts.

    python tools/ncu_sass_regions.py sass.csv [n_regions | addr,addr,...]
"""
import csv
import sys


def main(path, spec="24"):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ix = {k: hdr.index(k) for k in ("Address", "Source", "# Samples", "Instructions Executed", "L1 Wavefronts Shared",
                                    "L1 Wavefronts Shared Ideal", "stall_barrier", "stall_mio", "stall_short_sb",
                                    "stall_math", "stall_long_sb", "stall_wait")}
    data = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        def f(k):
            try:
                return float(r[ix[k]])
            except ValueError:
                return 0.0
        data.append((int(r[ix["Address"]], 16) if r[ix["Address"]].startswith("0x") else int(r[ix["Address"]]),
                     r[ix["Source"]], f("# Samples"), f("Instructions Executed"), f("L1 Wavefronts Shared"),
                     f("L1 Wavefronts Shared Ideal"), f("stall_barrier"), f("stall_mio"), f("stall_short_sb"),
                     f("stall_math"), f("stall_long_sb"), f("stall_wait")))
    base = data[0][0]
    if "," in spec:
        cuts = [int(x, 16) for x in spec.split(",")]
    else:
        n = int(spec)
        cuts = [i * len(data) // n * 16 for i in range(1, n)]
    cuts = [0] + cuts + [1 << 40]
    tot_s = sum(d[2] for d in data) or 1
    tot_i = sum(d[3] for d in data) or 1
    print("{:>7s} {:>7s} {:>7s} {:>7s} {:>9s} {:>9s} {:>6s} {:>6s} {:>6s} {:>6s} {:>6s}  first instruction".format(
        "from", "to", "smp%", "inst%", "wavefr M", "ideal M", "barr%", "mio%", "ssb%", "math%", "lsb%"))
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        seg = [d for d in data if lo <= d[0] - base < hi]
        if not seg:
            continue
        s = sum(d[2] for d in seg)
        print("{:7x} {:7x} {:7.2f} {:7.2f} {:9.2f} {:9.2f} {:6.1f} {:6.1f} {:6.1f} {:6.1f} {:6.1f}  {}".format(
            seg[0][0] - base, seg[-1][0] - base, 100 * s / tot_s, 100 * sum(d[3] for d in seg) / tot_i,
            sum(d[4] for d in seg) / 1e6, sum(d[5] for d in seg) / 1e6,
            100 * sum(d[6] for d in seg) / tot_s, 100 * sum(d[7] for d in seg) / tot_s,
            100 * sum(d[8] for d in seg) / tot_s, 100 * sum(d[9] for d in seg) / tot_s,
            100 * sum(d[10] for d in seg) / tot_s, seg[0][1][:60]))


if __name__ == "__main__":
    main(*sys.argv[1:])
