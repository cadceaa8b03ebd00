"""Command-line options: the same flags, short names and defaults as the
reference's common/option.py (Base :13-32, Train :160-186, Test :189-199), so
`--stages/--modes/-e/--scale/-r/--interval/--lutName/--testDir/--resultRoot`
select the same LUT files and loop bounds.

Differences, on purpose: parsing has no side effects unless asked (the
reference copies every *.py under cwd into expDir/code on each parse,
option.py:104-110,155-156); `--load_from_opt_file` is accepted and ignored
(broken upstream, option.py:79-90).
"""
from __future__ import annotations

import argparse
import os


def _base(p: argparse.ArgumentParser) -> None:
    p.add_argument("--model", type=str, default="SRNets")
    p.add_argument("--task", "-t", type=str, default="sr")
    p.add_argument("--scale", "-r", type=int, default=4, help="up scale factor")
    p.add_argument("--sigma", "-s", type=int, default=25, help="noise level")
    p.add_argument("--qf", "-q", type=int, default=20, help="deblocking quality factor")
    p.add_argument("--nf", type=int, default=64, help="number of filters of convolutional layers")
    p.add_argument("--stages", type=int, default=2, help="stages of MuLUT")
    p.add_argument("--modes", type=str, default="sdy", help="sampling modes to use in every stage")
    p.add_argument("--interval", type=int, default=4, help="N bit uniform sampling")
    p.add_argument("--modelRoot", type=str, default="../models")
    p.add_argument("--expDir", "-e", type=str, default="", help="experiment folder")
    p.add_argument("--load_from_opt_file", action="store_true", default=False)
    p.add_argument("--debug", default=False, action="store_true")
    # additions of this implementation
    p.add_argument("--device", type=int, default=0, help="CUDA device index (one process per GPU)")


def _finish(opt, is_train: bool):
    opt.isTrain = is_train
    # option.py:92-102
    opt.flag = opt.sigma if "dn" in opt.task else opt.qf if "db" in opt.task else opt.scale if "sr" in opt.task else "0"
    if opt.expDir == "":
        # option.py:119-134 picks modelRoot/debug/expr_<first free N> (and creates it);
        # the same path is chosen here, creation is left to the caller.
        model_dir = os.path.join(opt.modelRoot, "debug")
        count = 1
        while os.path.isdir(os.path.join(model_dir, "expr_{}".format(count))):
            count += 1
        opt.expDir = os.path.join(model_dir, "expr_{}".format(count))
    opt.modelPath = os.path.join(opt.expDir, "Model.pth")
    return opt


class TestOptions:
    def __init__(self, debug: bool = False):
        self.debug = debug

    def build_parser(self) -> argparse.ArgumentParser:
        p = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
        _base(p)
        p.add_argument("--loadIter", "-i", type=int, default=200000)
        p.add_argument("--testDir", type=str, default="../data/SRBenchmark")
        p.add_argument("--resultRoot", type=str, default="../results")
        p.add_argument("--lutName", type=str, default="LUT_ft")
        p.add_argument("--datasets", type=str, default="Set5",
                       help="comma separated benchmark folders under --testDir")
        return p

    def parse(self, argv=None):
        p = self.build_parser()
        opt = p.parse_args([] if self.debug else argv)
        return _finish(opt, False)


class TrainOptions:
    def __init__(self, debug: bool = False):
        self.debug = debug

    def build_parser(self) -> argparse.ArgumentParser:
        p = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
        _base(p)
        p.add_argument("--batchSize", type=int, default=32)
        p.add_argument("--cropSize", type=int, default=48, help="input LR training patch size")
        p.add_argument("--trainDir", type=str, default="../data/DIV2K")
        p.add_argument("--valDir", type=str, default="../data/SRBenchmark")
        p.add_argument("--startIter", type=int, default=0)
        p.add_argument("--totalIter", type=int, default=200000, help="Total number of training iterations")
        p.add_argument("--displayStep", type=int, default=100, help="display info every N iteration")
        p.add_argument("--valStep", type=int, default=2000, help="validate every N iteration")
        p.add_argument("--saveStep", type=int, default=2000, help="save models every N iteration")
        p.add_argument("--lr0", type=float, default=1e-3)
        p.add_argument("--lr1", type=float, default=1e-4)
        p.add_argument("--weightDecay", type=float, default=0)
        p.add_argument("--gpuNum", "-g", type=int, default=1)
        p.add_argument("--workerNum", "-n", type=int, default=8)
        p.add_argument("--synthetic", action="store_true", default=False,
                       help="train on seeded synthetic patches (no dataset on disk)")
        p.add_argument("--eager", action="store_true", default=False,
                       help="launch every kernel of a step from Python instead of replaying one CUDA graph")
        return p

    def parse(self, argv=None):
        p = self.build_parser()
        opt = p.parse_args([] if self.debug else argv)
        return _finish(opt, True)
