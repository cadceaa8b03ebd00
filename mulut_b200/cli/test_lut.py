"""`python -m mulut_b200.cli.test_lut --stages 2 --modes sdy -e ../models/sr_x2sdy`

Drop-in for the reference's step 4 (sr/4_test_lut.py:319-340): same flags, same
LUT file-name rule, same dataset layout, same result file names and summary
line; the per-image work runs on the GPU through LutEngine."""
from __future__ import annotations

import sys

from ..infer import eltr, load_luts
from ..options import TestOptions


def main(argv=None):
    opt = TestOptions().parse(argv)
    lutDict = load_luts(opt.expDir, opt.stages, opt.modes, opt.scale, opt.interval, opt.lutName)
    results = {}
    for dataset in [d for d in opt.datasets.split(",") if d]:
        results[dataset] = eltr(dataset, opt, lutDict, device=opt.device).run()
    return results


if __name__ == "__main__":
    main(sys.argv[1:])
