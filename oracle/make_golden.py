"""Generate the committed fixtures under tests/golden/ FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden

Everything written here is either a byte-for-byte artefact of the reference
repository (its shipped LUT tables, its five golden Set5 result PNGs and the
five LR inputs, stored decoded) or an output of the reference's own Python code
(imported unmodified by oracle/ref_import.py) on seeded synthetic inputs.
The GPU box has no /root/reference, so the parity tests read these files.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mulut_oracle as O          # noqa: E402
from oracle import ref_import as R            # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def set5():
    from PIL import Image
    names = ["baby", "bird", "butterfly", "head", "woman"]
    d = {}
    for n in names:
        d["lr_" + n] = np.array(Image.open(os.path.join(R.REF_ROOT, "data/SRBenchmark/Set5/LR_bicubic/X4", n + ".png")))
        d["sr_" + n] = np.array(Image.open(os.path.join(R.REF_ROOT, "results/sr_x2sdy/Set5/X4", n + "_LUT_ft_4bit.png")))
    np.savez_compressed(os.path.join(GOLD, "set5_x4.npz"), **d)
    print("set5_x4.npz", {k: v.shape for k, v in d.items()})


def set5_hr_metrics():
    """The five Set5 HR images (decoded) and the report numbers the reference's OWN evaluation code
    (common/utils.py:28-101 modcrop, _rgb2ycbcr, PSNR, cal_ssim, called as eltr._worker does,
    sr/4_test_lut.py:273-314) gives for its golden results: the known answer of sr/4_test_lut.py:343."""
    from PIL import Image
    utils = R.utils_module()
    names = ["baby", "bird", "butterfly", "head", "woman"]
    d, rows = {}, {}
    for n in names:
        hr = np.array(Image.open(os.path.join(R.REF_ROOT, "data/SRBenchmark/Set5/HR", n + ".png")))
        sr = np.array(Image.open(os.path.join(R.REF_ROOT, "results/sr_x2sdy/Set5/X4", n + "_LUT_ft_4bit.png")))
        d["hr_" + n] = hr
        gt = utils.modcrop(hr, 4)
        y_gt, y_out = utils._rgb2ycbcr(gt)[:, :, 0], utils._rgb2ycbcr(sr)[:, :, 0]
        rows[n] = {"psnr": float(utils.PSNR(y_gt, y_out, 4)), "ssim": float(utils.cal_ssim(y_gt, y_out)),
                   "y_gt_sum": float(y_gt.sum()), "y_out_sum": float(y_out.sum())}
    rows["_mean"] = {"psnr": float(np.mean([rows[n]["psnr"] for n in names])),
                     "ssim": float(np.mean([rows[n]["ssim"] for n in names]))}
    rows["_printed"] = "Dataset Set5 | AVG LUT PSNR: {:.2f} SSIM: {:.4f}".format(rows["_mean"]["psnr"], rows["_mean"]["ssim"])
    np.savez_compressed(os.path.join(GOLD, "set5_hr.npz"), **d)
    with open(os.path.join(GOLD, "set5_metrics.json"), "w") as f:
        json.dump(rows, f, indent=1)
    print("set5_hr.npz / set5_metrics.json:", rows["_printed"])


def shipped_luts():
    out = os.path.join(GOLD, "luts_x4")
    os.makedirs(out, exist_ok=True)
    for s in (1, 2):
        for m in "sdy":
            name = "LUT_ft_x4_4bit_int8_s{}_{}.npy".format(s, m)
            arr = np.load(os.path.join(R.REF_ROOT, "models", "sr_x2sdy", name))
            np.save(os.path.join(out, name), arr)
    print("luts_x4/ written")


PIPE_CASES = [
    # (name, H, W, C, stages, modes, scale, img_seed, lut_seed)
    ("x2_sdy_2st", 37, 53, 3, 2, "sdy", 2, 100, 1),
    ("x4_sdy_2st", 21, 30, 3, 2, "sdy", 4, 101, 2),
    ("x3_sdy_2st", 18, 13, 3, 2, "sdy", 3, 102, 3),
    ("x1_sdy_2st", 25, 25, 3, 2, "sdy", 1, 103, 4),
    ("x2_sdy_3st", 19, 23, 3, 3, "sdy", 2, 104, 5),
    ("x2_sdy_1st", 16, 40, 3, 1, "sdy", 2, 105, 6),
    ("x2_ys_2st", 20, 17, 3, 2, "ys", 2, 106, 7),
    ("x2_d_2st", 9, 31, 3, 2, "d", 2, 107, 8),
    ("x4_tiny_1x1", 1, 1, 3, 2, "sdy", 4, 108, 9),
    ("x2_tiny_2x3", 2, 3, 3, 2, "sdy", 2, 109, 10),
    ("x2_row_1x97", 1, 97, 3, 2, "sdy", 2, 110, 11),
    ("x2_extremes", 12, 12, 3, 2, "sdy", 2, -1, 12),
    ("x2_wide_tiles", 40, 101, 3, 2, "sdy", 2, 111, 13),
]


def case_image(H, W, C, seed):
    if seed < 0:                                   # 0/255 extremes + constant blocks
        img = np.zeros((H, W, C), np.uint8)
        img[::2, 1::2] = 255
        img[H // 2:, : W // 2] = 240
        img[: H // 2, W // 2:] = 15
        return img
    return np.random.default_rng(seed).integers(0, 256, (H, W, C), dtype=np.uint8)


def pipeline_cases():
    d = {}
    meta = []
    for name, H, W, C, stages, modes, scale, iseed, lseed in PIPE_CASES:
        img = case_image(H, W, C, iseed)
        luts = O.random_luts(lseed, stages, modes, scale)
        ref = R.ref_pipeline(img, luts, stages, list(modes), scale)
        mine = O.sr_pipeline(img, luts, stages, modes, scale)
        assert ref.dtype == np.uint8 and (ref == mine).all(), name
        d["in_" + name] = img
        d["out_" + name] = ref
        meta.append(dict(name=name, H=H, W=W, C=C, stages=stages, modes=modes, scale=scale, img_seed=iseed,
                         lut_seed=lseed))
    np.savez_compressed(os.path.join(GOLD, "ref_pipeline_cases.npz"), **d)
    with open(os.path.join(GOLD, "ref_pipeline_cases.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("ref_pipeline_cases:", len(meta))


INTERVAL_CASES = [
    # (name, H, W, C, stages, modes, scale, interval, img_seed, lut_seed): --interval other than the shipped 4
    # (common/option.py `--interval`: q = 2^interval, L = 2^(8-interval) + 1) and scale 3 at several intervals
    ("i5_x2_sdy_2st", 23, 31, 3, 2, "sdy", 2, 5, 120, 31),
    ("i5_x3_sdy_2st", 14, 19, 3, 2, "sdy", 3, 5, 121, 32),
    ("i5_x4_sdy_1st", 11, 13, 3, 1, "sdy", 4, 5, 122, 33),
    ("i6_x2_sdy_2st", 17, 29, 3, 2, "sdy", 2, 6, 123, 34),
    ("i7_x4_yd_2st", 9, 12, 3, 2, "yd", 4, 7, 124, 35),
    ("i3_x2_sdy_2st", 13, 21, 3, 2, "sdy", 2, 3, 125, 36),
    ("i4_x3_sdy_3st", 16, 11, 3, 3, "sdy", 3, 4, 126, 37),
    ("i6_x1_s_2st", 8, 8, 1, 2, "s", 1, 6, 127, 38),
]


def interval_cases():
    """Whole-pipeline cases at other intervals / scale 3, generated by the reference's own code."""
    d = {}
    meta = []
    for name, H, W, C, stages, modes, scale, interval, iseed, lseed in INTERVAL_CASES:
        img = case_image(H, W, C, iseed)
        if C == 1:
            img = img[:, :, 0]                       # grey input: the reference replicates it to 3 channels
        luts = O.random_luts(lseed, stages, modes, scale, interval)
        ref = R.ref_pipeline(img, luts, stages, list(modes), scale, interval)
        mine = O.sr_pipeline(img, luts, stages, modes, scale, interval)
        assert ref.dtype == np.uint8 and (ref == mine).all(), name
        d["in_" + name] = img
        d["out_" + name] = ref
        meta.append(dict(name=name, H=H, W=W, C=C, stages=stages, modes=modes, scale=scale, interval=interval,
                         img_seed=iseed, lut_seed=lseed))
    np.savez_compressed(os.path.join(GOLD, "ref_interval_cases.npz"), **d)
    with open(os.path.join(GOLD, "ref_interval_cases.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("ref_interval_cases:", len(meta))


def pass_cases():
    fn = R.test_lut_module().FourSimplexInterpFaster
    rng = np.random.default_rng(200)
    d = {}
    meta = []
    i = 0
    for mode in "sdy":
        for up in (1, 2, 4):
            for rot in (1, 2, 3, 4):
                p = O.MODE_PAD[mode]
                h, w = 7, 5
                x = rng.integers(0, 256, (3, h + p, w + p)).astype(np.float32)
                # the LUT is regenerated from its own seed by the tests (keeps the fixture small)
                lut = np.random.default_rng(300 + i).integers(-127, 128, (83521, up * up), dtype=np.int8)
                ref = fn(lut.astype(np.float32), x, h, w, 4, rot, upscale=up, mode=mode)
                assert (ref == O.four_simplex_interp(lut, x, h, w, 4, rot, up, mode)).all()
                d["x_%d" % i] = x
                d["out_%d" % i] = ref
                meta.append(dict(i=i, mode=mode, up=up, rot=rot, h=h, w=w, lut_seed=300 + i))
                i += 1
    np.savez_compressed(os.path.join(GOLD, "ref_pass_cases.npz"), **d)
    with open(os.path.join(GOLD, "ref_pass_cases.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("ref_pass_cases:", len(meta))


def sparse(a):
    a = np.asarray(a)
    idx = np.flatnonzero(a)
    return idx.astype(np.int64), a.reshape(-1)[idx]


def finetune_cases():
    import torch
    torch.manual_seed(0)
    torch.set_num_threads(4)
    model = R.model_module()
    d = {}
    meta = []
    # ---- (1) single InterpTorchBatch calls: out, d/dweight, d/dimg ----
    dummy = object.__new__(model.MuLUT)
    torch.nn.Module.__init__(dummy)
    dummy.interval = 4
    i = 0
    for mode in "sdy":
        for up in (1, 2, 4):
            rng = np.random.default_rng(400 + i)
            B, C, h, w = 2, 2, 6, 5
            bd = O.MODE_PAD[mode]
            # weights: int8/127 plus noise, some beyond the +-1 clamp
            wnp = (rng.integers(-140, 141, (83521, up * up)) / 127.0 + rng.normal(0, 1e-3, (83521, up * up))).astype(np.float32)
            x = rng.integers(0, 256, (B, C, h + bd, w + bd)).astype(np.float32)
            x[0, 0, :3, :3] = 37.0                    # exact ties between fractions
            x[1, 1, 2:, 2:] = 255.0
            g = rng.normal(0, 1, (B, C, h * up, w * up)).astype(np.float32)
            wt = torch.tensor(wnp, requires_grad=True)
            xt = torch.tensor(x, requires_grad=True)
            out = dummy.InterpTorchBatch(wt, up, mode, xt, bd)
            out.backward(torch.tensor(g))
            gi, gv = sparse(wt.grad.numpy())
            d["out_%d" % i] = out.detach().numpy()
            d["gw_idx_%d" % i] = gi
            d["gw_val_%d" % i] = gv
            d["gx_%d" % i] = xt.grad.numpy()
            meta.append(dict(i=i, mode=mode, up=up, B=B, C=C, h=h, w=w, bd=bd, seed=400 + i))
            i += 1
    # ---- (2) whole MuLUT.forward + mse_loss backward, all six tables ----
    for ci, (scale, B, C, h, w, lseed) in enumerate([(4, 2, 1, 12, 10, -1), (2, 2, 1, 9, 11, 21), (4, 1, 3, 7, 8, 22)]):
        with tempfile.TemporaryDirectory() as tmp:
            if lseed < 0:
                luts = R.shipped_luts()
            else:
                luts = O.random_luts(lseed, 2, "sdy", scale)
            for k, v in luts.items():
                np.save(os.path.join(tmp, "LUT_x{}_4bit_int8_{}.npy".format(scale, k)), v)
            net = model.MuLUT(lut_folder=tmp, stages=2, modes=["s", "d", "y"], upscale=scale, interval=4)
        rng = np.random.default_rng(500 + ci)
        im = (rng.integers(0, 256, (B, C, h, w)) / 255.0).astype(np.float32)
        lb = (rng.integers(0, 256, (B, C, h * scale, w * scale)) / 255.0).astype(np.float32)
        pred = net(torch.tensor(im))
        loss = torch.nn.functional.mse_loss(pred, torch.tensor(lb))
        loss.backward()
        d["fw_pred_%d" % ci] = pred.detach().numpy()
        d["fw_loss_%d" % ci] = np.float32(loss.item())
        for k in luts:
            gi, gv = sparse(getattr(net, "weight_" + k).grad.numpy())
            d["fw_g_idx_%d_%s" % (ci, k)] = gi
            d["fw_g_val_%d_%s" % (ci, k)] = gv
        meta.append(dict(forward_case=ci, scale=scale, B=B, C=C, h=h, w=w, lut_seed=lseed, seed=500 + ci))
    np.savez_compressed(os.path.join(GOLD, "ref_finetune_cases.npz"), **d)
    with open(os.path.join(GOLD, "ref_finetune_cases.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("ref_finetune_cases:", len(meta))


def main():
    if not R.available():
        raise SystemExit("reference tree not found at " + R.REF_ROOT)
    os.makedirs(GOLD, exist_ok=True)
    set5()
    set5_hr_metrics()
    shipped_luts()
    pipeline_cases()
    interval_cases()
    pass_cases()
    finetune_cases()


if __name__ == "__main__":
    main()
