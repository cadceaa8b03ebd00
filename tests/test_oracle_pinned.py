"""Pins the oracles (numpy, C, torch) to the reference: golden Set5 PNGs shipped by
the reference, fixtures produced by running the reference's own code
(oracle/make_golden.py), and - when /root/reference is present - the reference
imported live.  CPU only."""
import numpy as np
import pytest

from conftest import densify, norm_max_err
from oracle import c_oracle as CO
from oracle import mulut_oracle as O
from oracle import ref_import as R

NAMES = ["baby", "bird", "butterfly", "head", "woman"]


@pytest.mark.parametrize("name", NAMES)
def test_numpy_oracle_reproduces_reference_golden_pngs(name, set5, shipped_luts):
    out = O.sr_pipeline(set5["lr_" + name], shipped_luts, 2, "sdy", 4, 4)
    assert out.shape == set5["sr_" + name].shape
    assert (out == set5["sr_" + name]).all()


def test_c_oracle_reproduces_reference_golden_pngs(set5, shipped_luts):
    for name in NAMES:
        out = CO.sr_u8(set5["lr_" + name], shipped_luts, 2, "sdy", 4)
        assert (out == set5["sr_" + name]).all(), name


def test_oracles_match_reference_pipeline_fixtures(pipeline_cases):
    meta, data = pipeline_cases
    for c in meta:
        luts = O.random_luts(c["lut_seed"], c["stages"], c["modes"], c["scale"])
        img, ref = data["in_" + c["name"]], data["out_" + c["name"]]
        assert (O.sr_pipeline(img, luts, c["stages"], c["modes"], c["scale"]) == ref).all(), c["name"]
        assert (CO.sr_u8(img, luts, c["stages"], c["modes"], c["scale"]) == ref).all(), c["name"]
        assert (CO.sr_u8(img, luts, c["stages"], c["modes"], c["scale"], nthreads=1) == ref).all(), c["name"]


def test_oracles_match_reference_interval_fixtures(interval_cases):
    """--interval 3/5/6/7 and scale 3: numpy and C oracles against outputs of the reference's own code."""
    meta, data = interval_cases
    for c in meta:
        luts = O.random_luts(c["lut_seed"], c["stages"], c["modes"], c["scale"], c["interval"])
        img, ref = data["in_" + c["name"]], data["out_" + c["name"]]
        assert (O.sr_pipeline(img, luts, c["stages"], c["modes"], c["scale"], c["interval"]) == ref).all(), c["name"]
        img3 = img if img.ndim == 3 else np.stack([img] * 3, axis=2)
        assert (CO.sr_u8(img3, luts, c["stages"], c["modes"], c["scale"], c["interval"]) == ref).all(), c["name"]


def test_numpy_oracle_matches_reference_pass_fixtures(pass_cases):
    meta, data = pass_cases
    for c in meta:
        lut = np.random.default_rng(c["lut_seed"]).integers(-127, 128, (83521, c["up"] ** 2), dtype=np.int8)
        out = O.four_simplex_interp(lut, data["x_%d" % c["i"]], c["h"], c["w"], 4, c["rot"], c["up"], c["mode"])
        ref = data["out_%d" % c["i"]]
        assert out.shape == ref.shape and out.dtype == ref.dtype == np.float64
        assert (out == ref).all(), c


def test_unknown_mode_raises_like_reference():
    with pytest.raises(ValueError, match="Mode e not implemented."):
        O.sr_pipeline(np.zeros((4, 4, 3), np.uint8), {}, 1, "e", 2)
    with pytest.raises(ValueError, match="Mode h not implemented."):
        CO.sr_u8(np.zeros((4, 4, 3), np.uint8), {"s1_h": np.zeros((83521, 4), np.int8)}, 1, "h", 2)


def test_round_half_even_integer_form():
    num = np.arange(-400, 4000)
    for den in (48, 192, 16, 64, 3):
        ref = np.round(num / den)          # float64 is exact enough for these magnitudes
        assert (O.round_half_even_div(num, den) == ref).all()


def test_sorted_simplex_weights_sum_and_order():
    rng = np.random.default_rng(0)
    t = rng.integers(0, 256, (4, 5000))
    verts, w, order = O.simplex_vertices(t, 4)
    assert (w.sum(axis=0) == 16).all() and (w >= 0).all()
    assert (verts[4] - verts[0] == 17 ** 3 + 17 ** 2 + 17 + 1).all()
    assert verts.max() <= 83520
    # ties: higher tap index first
    _, _, o = O.simplex_vertices(np.array([[5], [5], [5], [5]]), 4)
    assert o[:, 0].tolist() == [3, 2, 1, 0]


def _torch_case(c, data):
    import torch
    from oracle import interp_torch_oracle as TO
    rng = np.random.default_rng(c["seed"])
    up, mode, B, C, h, w, bd = c["up"], c["mode"], c["B"], c["C"], c["h"], c["w"], c["bd"]
    wnp = (rng.integers(-140, 141, (83521, up * up)) / 127.0 + rng.normal(0, 1e-3, (83521, up * up))).astype(np.float32)
    x = rng.integers(0, 256, (B, C, h + bd, w + bd)).astype(np.float32)
    x[0, 0, :3, :3] = 37.0
    x[1, 1, 2:, 2:] = 255.0
    g = rng.normal(0, 1, (B, C, h * up, w * up)).astype(np.float32)
    return wnp, x, g


def test_torch_oracle_matches_reference_interp_fixtures(finetune_cases):
    import torch
    from oracle import interp_torch_oracle as TO
    meta, data = finetune_cases
    for c in [m for m in meta if "i" in m]:
        wnp, x, g = _torch_case(c, data)
        wt = torch.tensor(wnp, requires_grad=True)
        xt = torch.tensor(x, requires_grad=True)
        out = TO.interp_torch_batch(wt, c["up"], c["mode"], xt, c["bd"])
        out.backward(torch.tensor(g))
        i = c["i"]
        assert np.array_equal(out.detach().numpy(), data["out_%d" % i]), c
        gw_ref = densify(data["gw_idx_%d" % i], data["gw_val_%d" % i], wnp.shape)
        assert norm_max_err(wt.grad.numpy(), gw_ref) < 1e-6, c
        assert norm_max_err(xt.grad.numpy(), data["gx_%d" % i]) < 1e-6, c


def test_torch_oracle_matches_reference_forward_fixtures(finetune_cases, shipped_luts):
    import torch
    from oracle import interp_torch_oracle as TO
    meta, data = finetune_cases
    for c in [m for m in meta if "forward_case" in m]:
        ci, scale = c["forward_case"], c["scale"]
        luts = shipped_luts if c["lut_seed"] < 0 else O.random_luts(c["lut_seed"], 2, "sdy", scale)
        rng = np.random.default_rng(c["seed"])
        im = (rng.integers(0, 256, (c["B"], c["C"], c["h"], c["w"])) / 255.0).astype(np.float32)
        lb = (rng.integers(0, 256, (c["B"], c["C"], c["h"] * scale, c["w"] * scale)) / 255.0).astype(np.float32)
        ws = {k: torch.tensor(v.reshape(v.shape[0], -1).astype(np.float32) / 127.0, requires_grad=True)
              for k, v in luts.items()}
        pred = TO.mulut_forward(ws, torch.tensor(im), 2, "sdy", scale)
        loss = torch.nn.functional.mse_loss(pred, torch.tensor(lb))
        loss.backward()
        assert np.array_equal(pred.detach().numpy(), data["fw_pred_%d" % ci])
        for k in luts:
            ref = densify(data["fw_g_idx_%d_%s" % (ci, k)], data["fw_g_val_%d_%s" % (ci, k)], ws[k].shape)
            assert norm_max_err(ws[k].grad.numpy(), ref) < 1e-5, (ci, k)


# ---- live reference (build container only) ---------------------------------
needs_ref = pytest.mark.skipif(not R.available(), reason="/root/reference not present")


@needs_ref
def test_live_reference_random_images():
    rng = np.random.default_rng(77)
    for scale, stages, modes, hw in [(2, 2, "sdy", (11, 17)), (4, 2, "sdy", (6, 9)), (2, 1, "y", (5, 5)),
                                     (3, 2, "ds", (8, 4))]:
        img = rng.integers(0, 256, hw + (3,), dtype=np.uint8)
        luts = O.random_luts(int(rng.integers(1 << 30)), stages, modes, scale)
        ref = R.ref_pipeline(img, luts, stages, list(modes), scale)
        assert (O.sr_pipeline(img, luts, stages, modes, scale) == ref).all()
        assert (CO.sr_u8(img, luts, stages, modes, scale) == ref).all()


@needs_ref
def test_live_reference_all_fraction_tuples():
    """All 16^4 LSB tuples x random MSBs as a (65536,2,2) 'image' in mode s."""
    fn = R.test_lut_module().FourSimplexInterpFaster
    rng = np.random.default_rng(3)
    f = np.stack(np.meshgrid(*[np.arange(16)] * 4, indexing="ij"), -1).reshape(-1, 4)
    m = rng.integers(0, 16, f.shape)
    x = (m * 16 + f).reshape(-1, 2, 2).astype(np.float32)
    for up in (1, 2):
        lut = rng.integers(-127, 128, (83521, up * up), dtype=np.int8)
        ref = fn(lut.astype(np.float32), x, 1, 1, 4, 4, upscale=up, mode="s")
        out = O.four_simplex_interp(lut, x, 1, 1, 4, 4, up, "s")
        assert (out == ref).all()


@pytest.mark.skipif(not R.full_tree(), reason="/root/reference not present (the staged copy holds the hot-path modules only)")
def test_live_reference_transfer_grid(monkeypatch):
    """mulut_b200.transfer enumerates the LUT grid exactly like the reference's own
    get_input_tensor / get_mode_input_tensor (sr/2_transfer_to_lut.py:12-66).  The reference calls
    .cuda() on its index vectors: that single method is neutralised for the call."""
    import types
    import torch
    from mulut_b200 import transfer as T
    ref = R.transfer_module()
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    for interval in (4, 5):
        want = ref.get_input_tensor(types.SimpleNamespace(interval=interval))
        got = T.get_input_tensor(interval)
        assert got.shape == want.shape and torch.equal(got, want), interval
    x = T.get_input_tensor(5)
    for mode in "dy":
        assert torch.equal(T.get_mode_input_tensor(x, mode), ref.get_mode_input_tensor(x, mode)), mode
