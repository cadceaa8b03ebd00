"""K1f (binned shared-memory last stage, TMA tile ring) against the C oracle.

The kernel reads its tiles through TMA (16-byte aligned frames, row pitch a multiple of 16; other
frames go through a pitched staging copy first); the cases assert through the per-kernel profile
counters that K1f is the kernel that actually ran."""
import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import mulut_oracle as O

pytestmark = pytest.mark.gpu


def _run(eng, frames):
    import torch
    eng.profile(True)
    out = eng(torch.from_numpy(np.ascontiguousarray(frames)).cuda()).cpu().numpy()
    prof = eng.profile_read()
    eng.profile(False)
    assert "last_binned" in prof and "last_tiled" not in prof, prof
    return out


@pytest.mark.parametrize("C,shape", [(3, (1, 64, 96)), (3, (2, 35, 80)), (3, (3, 97, 208)), (3, (1, 1, 16)),
                                     (3, (1, 2, 16)), (3, (1, 130, 32)), (1, (2, 67, 96)), (1, (1, 33, 208)),
                                     (1, (1, 5, 16)), (3, (1, 200, 400)), (2, (2, 40, 48)), (4, (1, 70, 100)), (4, (2, 33, 24))])
@pytest.mark.parametrize("orphans", ["1", "0"])
def test_binned_matches_oracle(C, shape, orphans, monkeypatch):
    """orphans=0 keeps every non-empty bin in shared memory (small frames would otherwise be
    finished entirely by the orphan list kernel)."""
    from mulut_b200.infer import LutEngine
    monkeypatch.setenv("MULUT_BN_ORPHANS", orphans)
    N, H, W = shape
    rng = np.random.default_rng(H * 1000 + W + C)
    luts = O.random_luts(77 + C, 2, "sdy", 2)
    frames = rng.integers(0, 256, (N, H, W, C), dtype=np.uint8)
    ref = CO.sr_u8(frames, luts, 2, "sdy", 2)
    with LutEngine(luts, 2, "sdy", 2, 4, device=0, kernel=3) as eng:
        out = _run(eng, frames)
    assert (out == ref).all(), (shape, C, int((out != ref).sum()))


@pytest.mark.parametrize("orphans", ["1", "0"])
@pytest.mark.parametrize("kind", ["dark", "bright", "two_level", "ramp", "sparse_bins", "single_bin"])
def test_binned_skewed_value_distributions(kind, orphans, monkeypatch):
    """The CTA allocation follows the sample histogram: exercise empty, tiny and
    dominant bins (single stage, so the engine's input IS the binned stage's input)."""
    from mulut_b200.infer import LutEngine
    monkeypatch.setenv("MULUT_BN_ORPHANS", orphans)
    rng = np.random.default_rng(5)
    H, W = 150, 240
    if kind == "dark":
        img = rng.integers(0, 20, (1, H, W, 3))
    elif kind == "bright":
        img = rng.integers(236, 256, (1, H, W, 3))
    elif kind == "two_level":
        img = np.where(rng.random((1, H, W, 3)) < 0.5, 3, 250)
    elif kind == "ramp":
        img = np.broadcast_to((np.arange(W) * 255 // (W - 1))[None, None, :, None], (1, H, W, 3))
    elif kind == "sparse_bins":
        img = rng.integers(96, 160, (1, H, W, 3))
        img[0, ::37, ::41] = rng.integers(0, 256, img[0, ::37, ::41].shape)
    else:
        img = rng.integers(128, 160, (1, H, W, 3))
    img = np.ascontiguousarray(img, dtype=np.uint8)
    luts = O.random_luts(9, 1, "sdy", 2)
    ref = CO.sr_u8(img, luts, 1, "sdy", 2)
    with LutEngine(luts, 1, "sdy", 2, 4, device=0, kernel=3) as eng:
        out = _run(eng, img)
    assert (out == ref).all(), (kind, int((out != ref).sum()))


@pytest.mark.parametrize("modes,stages", [("s", 1), ("yd", 2), ("dsy", 3)])
def test_binned_mode_subsets_and_stages(modes, stages, monkeypatch):
    from mulut_b200.infer import LutEngine
    monkeypatch.setenv("MULUT_BN_ORPHANS", "0")
    rng = np.random.default_rng(11)
    luts = O.random_luts(21, stages, modes, 2)
    frames = rng.integers(0, 256, (2, 50, 112, 3), dtype=np.uint8)
    ref = CO.sr_u8(frames, luts, stages, modes, 2)
    with LutEngine(luts, stages, modes, 2, 4, device=0, kernel=3) as eng:
        out = _run(eng, frames)
    assert (out == ref).all(), (modes, stages, int((out != ref).sum()))


def test_unmappable_frames_go_through_the_pitched_copy():
    """W*C % 16 != 0 or a misaligned device pointer: TMA cannot map the frames in place, so the stage input
    is first copied into a pitched staging buffer - the TMA-fed kernels still run, still bit-exact."""
    import torch
    from mulut_b200.infer import LutEngine
    rng = np.random.default_rng(3)
    luts = O.random_luts(4, 2, "sdy", 2)
    for shape in [(41, 37, 3), (2, 33, 50, 1), (3, 70, 21, 4), (1, 19, 3, 2), (1, 40, 64, 3)]:
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        ref = CO.sr_u8(img, luts, 2, "sdy", 2)
        with LutEngine(luts, 2, "sdy", 2, 4, device=0, kernel=3) as eng:
            eng.profile(True)
            out = eng(torch.from_numpy(img).cuda()).cpu().numpy()
            prof = eng.profile_read()
            assert "last_binned" in prof and "last_tiled" not in prof, (shape, list(prof))
            assert (out == ref).all(), shape
            # the same frames at an address that is not 16-byte aligned
            buf = torch.empty(img.size + 64, dtype=torch.uint8, device="cuda")
            off = 16 - buf.data_ptr() % 16 + 5
            view = buf[off:off + img.size].view(*img.shape)
            view.copy_(torch.from_numpy(img))
            assert view.data_ptr() % 16 == 5
            assert (eng(view).cpu().numpy() == ref).all(), shape


@pytest.mark.parametrize("kernel", [3, -1])
def test_single_stage_model_with_a_misaligned_input_pointer(kernel):
    """A ONE-stage x2 model reads the caller's pointer in its last stage: K1f's histogram and orphan scan use
    16-byte loads on the dense input, so a misaligned pointer must not reach them (it takes K1c instead of
    faulting); forced (kernel 3) and through AUTO at >= 2^20 samples."""
    import torch
    from mulut_b200.infer import LutEngine
    rng = np.random.default_rng(21)
    luts = O.random_luts(22, 1, "sdy", 2)
    img = rng.integers(0, 256, (2, 384, 480, 3), dtype=np.uint8)          # 1.1 M samples, W*C % 16 == 0
    img[0, :200] = rng.integers(100, 140, (200, 480, 3))                   # some sparse bins as well
    ref = CO.sr_u8(img, luts, 1, "sdy", 2)
    with LutEngine(luts, 1, "sdy", 2, 4, device=0, kernel=kernel) as eng:
        for misalign in (0, 1, 5, 8):
            buf = torch.empty(img.size + 64, dtype=torch.uint8, device="cuda")
            off = (16 - buf.data_ptr() % 16) % 16 + misalign
            view = buf[off:off + img.size].view(*img.shape)
            view.copy_(torch.from_numpy(img))
            assert view.data_ptr() % 16 == misalign
            eng.profile(True)
            out = eng(view).cpu().numpy()
            prof = eng.profile_read()
            eng.profile(False)
            torch.cuda.synchronize()
            assert (out == ref).all(), (kernel, misalign, int((out != ref).sum()))
            assert ("last_binned" in prof) == (misalign == 0), (misalign, list(prof))


def test_orphan_list_gets_the_sparse_bins():
    """A frame whose values sit in two bins plus a sprinkle elsewhere: the sprinkle must go
    through the orphan list kernel, the two dense bins through shared memory."""
    import torch
    from mulut_b200.infer import LutEngine
    rng = np.random.default_rng(12)
    img = rng.integers(96, 160, (2, 256, 384, 3))
    sel = rng.random(img.shape) < 0.004
    img[sel] = rng.integers(0, 256, int(sel.sum()))
    img = np.ascontiguousarray(img, dtype=np.uint8)
    luts = O.random_luts(13, 1, "sdy", 2)
    ref = CO.sr_u8(img, luts, 1, "sdy", 2)
    with LutEngine(luts, 1, "sdy", 2, 4, device=0, kernel=3) as eng:
        eng.profile(True)
        out = eng(torch.from_numpy(img).cuda()).cpu().numpy()
        prof = eng.profile_read()
    assert "last_binned" in prof and "bin_orphans" in prof, prof
    assert (out == ref).all(), int((out != ref).sum())


def test_auto_picks_binned_for_large_launches_and_host_path_agrees():
    import torch
    from mulut_b200.infer import LutEngine
    rng = np.random.default_rng(8)
    luts = O.random_luts(2, 2, "sdy", 2)
    frames = rng.integers(0, 256, (2, 540, 960, 3), dtype=np.uint8)
    ref = CO.sr_u8(frames, luts, 2, "sdy", 2)
    with LutEngine(luts, 2, "sdy", 2, 4, device=0) as eng:
        eng.profile(True)
        out = eng(torch.from_numpy(frames).cuda()).cpu().numpy()
        prof = eng.profile_read()
        assert "last_binned" in prof, prof
        assert (out == ref).all(), int((out != ref).sum())
        assert (eng(frames) == ref).all()          # host path: one frame per launch


@pytest.mark.parametrize("scale,stages,modes,shape", [(1, 2, "sdy", (2, 70, 112, 3)), (1, 3, "ds", (1, 33, 96, 1)),
                                                      (4, 2, "sdy", (2, 45, 80, 3)), (3, 2, "ys", (1, 64, 160, 3)),
                                                      (1, 1, "y", (1, 2, 16, 3))])
@pytest.mark.parametrize("fused", ["0", "1"])
def test_tma_stage_kernel_k1g(scale, stages, modes, shape, fused, monkeypatch):
    """The TMA-fed shared-memory stage kernels serve every up = 1 stage of TMA-mappable frames, whatever the
    last stage is (scales 1/3/4 pair them with an up = 1 last stage, K0 or K1e): K1h + K1b (the default) and,
    with MULUT_K1_FUSED=1, K1i (stage + mode combine + epilogue in one cooperative kernel)."""
    import torch
    monkeypatch.setenv("MULUT_K1_FUSED", fused)
    from mulut_b200.infer import LutEngine
    rng = np.random.default_rng(sum(shape) + scale)
    luts = O.random_luts(50 + scale, stages, modes, scale)
    frames = rng.integers(0, 256, shape, dtype=np.uint8)
    ref = CO.sr_u8(frames, luts, stages, modes, scale)
    with LutEngine(luts, stages, modes, scale, 4, device=0) as eng:          # AUTO
        eng.profile(True)
        out = eng(torch.from_numpy(frames).cuda()).cpu().numpy()
        prof = eng.profile_read()
    if stages > 1 or scale == 1:
        if fused == "1":
            assert "fused_stage" in prof and "smem_stage" not in prof and "combine" not in prof, prof
        else:
            assert "smem_stage" in prof and "combine" in prof and "fused_stage" not in prof, prof
    assert (out == ref).all(), (scale, stages, modes, int((out != ref).sum()))
    with LutEngine(luts, stages, modes, scale, 4, device=0, kernel=1) as eng:  # K1a (no TMA) gives the same bytes
        assert (eng(torch.from_numpy(frames).cuda()).cpu().numpy() == ref).all()


@pytest.mark.parametrize("modes,stages,scale", [("sdy", 2, 2), ("sd", 3, 2), ("y", 2, 2), ("sdys", 2, 1)])
def test_fused_stage_kernel_many_tiles_and_back_to_back_launches(modes, stages, scale, monkeypatch):
    """K1i (opt-in: MULUT_K1_FUSED=1) over enough tiles that every stream reuses its exchange slots many times,
    with 1 to 4 modes, launched back to back (the kernel must leave its hand-off counters zeroed) and from two
    streams of one handle."""
    import torch
    from mulut_b200.infer import LutEngine
    monkeypatch.setenv("MULUT_K1_FUSED", "1")
    rng = np.random.default_rng(len(modes) * 10 + stages)
    luts = O.random_luts(70 + stages, stages, modes, scale)
    frames = rng.integers(0, 256, (3, 600, 1120, 3), dtype=np.uint8)       # 1197 stage tiles: ~12 per stream
    ref = CO.sr_u8(frames, luts, stages, modes, scale)
    with LutEngine(luts, stages, modes, scale, 4, device=0) as eng:
        d = torch.from_numpy(frames).cuda()
        eng.profile(True)
        outs = [eng(d) for _ in range(3)]
        prof = eng.profile_read()
        eng.profile(False)
        assert "fused_stage" in prof and "combine" not in prof, prof
        for o in outs:
            assert (o.cpu().numpy() == ref).all()
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            o2 = eng(d)
        o1 = eng(d)
        torch.cuda.synchronize()
        assert (o1.cpu().numpy() == ref).all() and (o2.cpu().numpy() == ref).all()
        small = rng.integers(0, 256, (1, 5, 16, 3), dtype=np.uint8)          # fewer tiles than streams
        assert (eng(torch.from_numpy(small).cuda()).cpu().numpy() == CO.sr_u8(small, luts, stages, modes, scale)).all()


def test_binned_sparse_resident_bins_walk_many_tiles(monkeypatch):
    """With orphaning off, a bin holding 0.5 % of the samples still gets CTAs of its own: each walks
    hundreds of tiles, wraps the TMA ring many times and lives on carried queue entries and forced
    partial rounds - the paths a dense bin never takes."""
    import torch
    from mulut_b200.infer import LutEngine
    monkeypatch.setenv("MULUT_BN_ORPHANS", "0")
    rng = np.random.default_rng(21)
    img = rng.integers(100, 156, (2, 540, 960, 3))
    sel = rng.random(img.shape) < 0.02
    img[sel] = rng.integers(0, 256, int(sel.sum()))
    img = np.ascontiguousarray(img, dtype=np.uint8)
    luts = O.random_luts(17, 1, "sdy", 2)
    ref = CO.sr_u8(img, luts, 1, "sdy", 2)
    with LutEngine(luts, 1, "sdy", 2, 4, device=0, kernel=3) as eng:
        out = _run(eng, img)
    assert (out == ref).all(), int((out != ref).sum())


@pytest.mark.parametrize("seed", range(12))
def test_binned_randomised_shapes_and_distributions(seed, monkeypatch):
    """Seeded fuzz: random TMA-mappable shapes, batch sizes, channel counts, stage counts and value
    distributions (uniform, narrow band, bimodal, heavy tails), with and without orphaning."""
    from mulut_b200.infer import LutEngine
    rng = np.random.default_rng(1000 + seed)
    C = int(rng.choice([1, 3]))
    W = int(rng.integers(1, 20)) * 16
    H = int(rng.integers(1, 150))
    N = int(rng.integers(1, 4))
    stages = int(rng.integers(1, 3))
    monkeypatch.setenv("MULUT_BN_ORPHANS", str(seed % 2))
    kind = seed % 4
    if kind == 0:
        img = rng.integers(0, 256, (N, H, W, C))
    elif kind == 1:
        lo = int(rng.integers(0, 200))
        img = rng.integers(lo, lo + int(rng.integers(2, 56)), (N, H, W, C))
    elif kind == 2:
        img = np.where(rng.random((N, H, W, C)) < 0.5, rng.integers(0, 64, (N, H, W, C)), rng.integers(192, 256, (N, H, W, C)))
    else:
        img = np.clip(rng.normal(128, 20, (N, H, W, C)) + (rng.random((N, H, W, C)) < 0.01) * rng.normal(0, 120, (N, H, W, C)), 0, 255)
    img = np.ascontiguousarray(img, dtype=np.uint8)
    luts = O.random_luts(300 + seed, stages, "sdy", 2)
    ref = CO.sr_u8(img, luts, stages, "sdy", 2)
    with LutEngine(luts, stages, "sdy", 2, 4, device=0, kernel=3) as eng:
        out = _run(eng, img)
    assert (out == ref).all(), (seed, img.shape, stages, int((out != ref).sum()))


@pytest.mark.parametrize("shape", [(2, 64, 96, 3), (1, 41, 37, 3)])
def test_device_path_is_cuda_graph_capturable_after_reserve(shape):
    """mulut_reserve pre-sizes every buffer the hot call needs (including the pitched staging copy of frames
    TMA cannot map in place), so mulut_sr_infer_u8 can be captured into a CUDA graph and replayed on new frames."""
    import torch
    from mulut_b200.infer import LutEngine
    rng = np.random.default_rng(21)
    luts = O.random_luts(31, 2, "sdy", 2)
    N, H, W, C = shape
    with LutEngine(luts, 2, "sdy", 2, 4, device=0, kernel=3) as eng:
        eng.reserve(N, H, W, C)
        d_in = torch.zeros(shape, dtype=torch.uint8, device="cuda")
        d_out = torch.empty((N, 2 * H, 2 * W, C), dtype=torch.uint8, device="cuda")
        side = torch.cuda.Stream()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            eng.infer_device(d_in, d_out)                  # warm-up outside the capture
            side.synchronize()
            with torch.cuda.graph(graph, stream=side):
                eng.infer_device(d_in, d_out)
        for seed in range(3):
            frames = rng.integers(0, 256, shape, dtype=np.uint8)
            d_in.copy_(torch.from_numpy(frames))
            graph.replay()
            torch.cuda.synchronize()
            assert (d_out.cpu().numpy() == CO.sr_u8(frames, luts, 2, "sdy", 2)).all(), (shape, seed)
