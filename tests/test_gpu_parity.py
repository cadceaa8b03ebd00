"""GPU parity tests proper: the CUDA path, called through the C ABI
(libmulut_b200.so via mulut_b200), against the oracle and the committed golden
fixtures.  Bit-exact for all uint8 work."""
import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import mulut_oracle as O

pytestmark = pytest.mark.gpu

NAMES = ["baby", "bird", "butterfly", "head", "woman"]
KERNELS = {"generic": 0, "tiled": 1, "cell": 2, "binned": 3}   # 1 = quad K1c, 2 = owner-only K1d, 3 = binned smem K1f


def _engine(luts, stages, modes, scale, kernel="tiled"):
    from mulut_b200.infer import LutEngine
    return LutEngine(luts, stages, modes, scale, 4, device=0, kernel=KERNELS[kernel])


@pytest.mark.parametrize("kernel", list(KERNELS))
def test_reference_golden_set5(kernel, set5, shipped_luts):
    """The reference's only golden vectors: results/sr_x2sdy/Set5/X4/*.png."""
    with _engine(shipped_luts, 2, "sdy", 4, kernel) as eng:
        for name in NAMES:
            out = eng(set5["lr_" + name])
            assert out.dtype == np.uint8 and out.shape == set5["sr_" + name].shape
            assert (out == set5["sr_" + name]).all(), name


@pytest.mark.parametrize("kernel", list(KERNELS))
def test_reference_generated_fixtures(kernel, pipeline_cases):
    meta, data = pipeline_cases
    for c in meta:
        luts = O.random_luts(c["lut_seed"], c["stages"], c["modes"], c["scale"])
        with _engine(luts, c["stages"], c["modes"], c["scale"], kernel) as eng:
            out = eng(data["in_" + c["name"]])
        ref = data["out_" + c["name"]]
        assert out.shape == ref.shape, c["name"]
        assert (out == ref).all(), (c["name"], int((out != ref).sum()))


@pytest.mark.parametrize("kernel", [-1, 0, 1])
def test_reference_generated_interval_fixtures(kernel, interval_cases):
    """--interval 3/5/6/7 and scale 3 (no shipped model uses them): outputs of the reference's own code."""
    from mulut_b200.infer import LutEngine
    meta, data = interval_cases
    for c in meta:
        luts = O.random_luts(c["lut_seed"], c["stages"], c["modes"], c["scale"], c["interval"])
        with LutEngine(luts, c["stages"], c["modes"], c["scale"], c["interval"], device=0, kernel=kernel) as eng:
            out = eng(data["in_" + c["name"]])
        ref = data["out_" + c["name"]]
        assert out.shape == ref.shape, c["name"]
        assert (out == ref).all(), (c["name"], int((out != ref).sum()))


@pytest.mark.parametrize("interval,scale,stages,modes", [(5, 2, 2, "sdy"), (6, 4, 2, "sdy"), (7, 3, 2, "yd"), (5, 1, 3, "sdy"),
                                                         (6, 3, 1, "s"), (5, 4, 2, "sdy")])
def test_small_table_intervals_at_size(interval, scale, stages, modes):
    """Intervals 5-7 at a few hundred thousand samples per launch (K0 serves them): bit-exact against the C oracle."""
    import torch
    from mulut_b200.infer import LutEngine
    rng = np.random.default_rng(interval * 10 + scale)
    luts = O.random_luts(80 + interval, stages, modes, scale, interval)
    frames = rng.integers(0, 256, (2, 211, 317, 3), dtype=np.uint8)           # 401 k samples
    ref = CO.sr_u8(frames, luts, stages, modes, scale, interval)
    with LutEngine(luts, stages, modes, scale, interval, device=0) as eng:
        out = eng(torch.from_numpy(frames).cuda()).cpu().numpy()
    assert (out == ref).all(), (interval, scale, int((out != ref).sum()))


def test_full_size_x3_matches_c_oracle():
    """Scale 3 at the shipped interval: K1h + K1b + K1e3 (the x4 cell kernel on 9-of-16-column cells), one
    640x360 frame -> 1920x1080, bit-exact against the C oracle for AUTO, the tiled selection and K0."""
    import torch
    from mulut_b200.infer import LutEngine
    rng = np.random.default_rng(303)
    luts = O.random_luts(33, 2, "sdy", 3)
    frame = rng.integers(0, 256, (2, 360, 640, 3), dtype=np.uint8)
    ref = CO.sr_u8(frame, luts, 2, "sdy", 3)
    for kernel in (-1, 1, 0):
        with LutEngine(luts, 2, "sdy", 3, 4, device=0, kernel=kernel) as eng:
            eng.profile(True)
            out = eng(torch.from_numpy(frame).cuda()).cpu().numpy()
            prof = eng.profile_read()
        assert ("last_tiled" in prof) == (kernel != 0), (kernel, list(prof))
        assert out.shape == (2, 1080, 1920, 3) and (out == ref).all(), (kernel, int((out != ref).sum()))


@pytest.mark.parametrize("kernel", list(KERNELS))
@pytest.mark.parametrize("C", [1, 2, 3, 4, 5])
def test_channel_counts_and_ragged_sizes(kernel, C):
    rng = np.random.default_rng(10 + C)
    luts = O.random_luts(40 + C, 2, "sdy", 2)
    with _engine(luts, 2, "sdy", 2, kernel) as eng:
        for (H, W) in [(1, 1), (3, 2), (17, 33), (35, 97), (64, 96), (33, 129)]:
            img = rng.integers(0, 256, (H, W, C), dtype=np.uint8)
            out = eng(img)
            ref = CO.sr_u8(img, luts, 2, "sdy", 2)
            assert (out == ref).all(), (C, H, W, int((out != ref).sum()))


@pytest.mark.parametrize("kernel", list(KERNELS))
def test_batch_device_path_and_determinism(kernel):
    import torch
    rng = np.random.default_rng(5)
    luts = O.random_luts(6, 2, "sdy", 2)
    frames = rng.integers(0, 256, (5, 70, 113, 3), dtype=np.uint8)
    ref = CO.sr_u8(frames, luts, 2, "sdy", 2)
    with _engine(luts, 2, "sdy", 2, kernel) as eng:
        d = torch.from_numpy(frames).cuda()
        o1 = eng(d)
        o2 = eng(d)
        torch.cuda.synchronize()
        assert o1.is_cuda and o1.dtype == torch.uint8
        assert (o1.cpu().numpy() == ref).all()
        assert torch.equal(o1, o2)
        # host path (pipelined H2D / kernels / D2H), pinned and pageable
        from mulut_b200.infer import pinned_empty
        pin = pinned_empty(frames.shape)
        pin[...] = frames
        pout = pinned_empty(ref.shape)
        eng.infer_host(pin, pout)
        assert (pout == ref).all()
        assert (eng(frames) == ref).all()
        assert eng.launch_count > 0


@pytest.mark.parametrize("kernel", ["tiled", "binned"])
def test_streaming_host_path_overlapping_calls(kernel):
    """mulut_sr_infer_u8_host_async / mulut_sr_host_sync: several batches of different frames and sizes in
    flight at once come back exactly as the blocking path computes them."""
    from mulut_b200.infer import pinned_empty
    rng = np.random.default_rng(11)
    luts = O.random_luts(8, 2, "sdy", 2)
    with _engine(luts, 2, "sdy", 2, kernel) as eng:
        jobs = []
        for n, hh, ww in [(4, 70, 112), (1, 70, 112), (7, 70, 112), (3, 96, 160), (5, 70, 112)]:
            fr = pinned_empty((n, hh, ww, 3))
            fr[...] = rng.integers(0, 256, fr.shape, dtype=np.uint8)
            out = pinned_empty((n, 2 * hh, 2 * ww, 3))
            out[...] = 0
            jobs.append((fr, out))
        for fr, out in jobs:
            eng.infer_host_async(fr, out)
        eng.host_sync()
        for fr, out in jobs:
            assert (out == CO.sr_u8(np.asarray(fr), luts, 2, "sdy", 2)).all()
        eng.host_sync()                                   # idempotent
        # pageable (plain numpy) buffers work too, only slower: results are complete after host_sync
        fr = rng.integers(0, 256, (3, 70, 112, 3), dtype=np.uint8)
        out = np.zeros((3, 140, 224, 3), dtype=np.uint8)
        eng.infer_host_async(fr, out)
        eng.infer_host_async(jobs[0][0], jobs[0][1])
        eng.host_sync()
        assert (out == CO.sr_u8(fr, luts, 2, "sdy", 2)).all()
        with pytest.raises(ValueError):
            eng.infer_host_async(jobs[0][0], jobs[1][1])  # wrong out shape


@pytest.mark.parametrize("scale,stages,modes", [(1, 2, "sdy"), (2, 1, "sdy"), (2, 3, "sdy"), (3, 2, "sdy"),
                                                (4, 2, "sdy"), (2, 2, "y"), (2, 2, "ds"), (4, 2, "yds")])
def test_scales_stages_modes(scale, stages, modes):
    rng = np.random.default_rng(scale * 100 + stages)
    luts = O.random_luts(scale * 7 + stages, stages, modes, scale)
    img = rng.integers(0, 256, (45, 71, 3), dtype=np.uint8)
    ref = CO.sr_u8(img, luts, stages, modes, scale)
    for kernel in KERNELS:
        with _engine(luts, stages, modes, scale, kernel) as eng:
            out = eng(img)
        assert (out == ref).all(), (kernel, scale, stages, modes, int((out != ref).sum()))


def test_other_intervals_generic():
    from mulut_b200.infer import LutEngine
    rng = np.random.default_rng(9)
    for interval in (5, 6):
        luts = O.random_luts(interval, 2, "sdy", 2, interval)
        img = rng.integers(0, 256, (20, 31, 3), dtype=np.uint8)
        ref = O.sr_pipeline(img, luts, 2, "sdy", 2, interval)
        with LutEngine(luts, 2, "sdy", 2, interval) as eng:
            assert (eng(img) == ref).all(), interval


def test_extremes_and_constant_images():
    luts = O.random_luts(3, 2, "sdy", 2)
    with _engine(luts, 2, "sdy", 2, "tiled") as eng:
        for val in (0, 15, 16, 127, 240, 255):
            img = np.full((19, 23, 3), val, np.uint8)
            assert (eng(img) == CO.sr_u8(img, luts, 2, "sdy", 2)).all(), val
        img = np.zeros((32, 32, 3), np.uint8)
        img[::2, ::2] = 255
        assert (eng(img) == CO.sr_u8(img, luts, 2, "sdy", 2)).all()
    # saturating LUTs exercise both clamps of the epilogue
    for fill in (-127, 127):
        l2 = {k: np.full_like(v, fill) for k, v in luts.items()}
        img = np.random.default_rng(1).integers(0, 256, (9, 9, 3), dtype=np.uint8)
        with _engine(l2, 2, "sdy", 2, "tiled") as eng:
            assert (eng(img) == CO.sr_u8(img, l2, 2, "sdy", 2)).all()


def test_full_size_1080p_x2_matches_c_oracle():
    """BASELINE config 2 at full size: one 1920x1080 RGB frame -> 4K."""
    import torch
    rng = np.random.default_rng(0)
    luts = O.random_luts(1, 2, "sdy", 2)
    frame = rng.integers(0, 256, (1, 1080, 1920, 3), dtype=np.uint8)
    ref = CO.sr_u8(frame, luts, 2, "sdy", 2)
    for kernel in KERNELS:
        with _engine(luts, 2, "sdy", 2, kernel) as eng:
            out = eng(torch.from_numpy(frame).cuda()).cpu().numpy()
        assert out.shape == (1, 2160, 3840, 3)
        assert (out == ref).all(), (kernel, int((out != ref).sum()))


def test_full_size_properties_x4_960x540(shipped_luts):
    """BASELINE config 3 at full size through size-independent properties:
    (a) tiled == generic kernel, (b) translation covariance of interior crops,
    (c) frames of a batch are independent."""
    import torch
    rng = np.random.default_rng(2)
    frames = rng.integers(0, 256, (2, 540, 960, 3), dtype=np.uint8)
    d = torch.from_numpy(frames).cuda()
    with _engine(shipped_luts, 2, "sdy", 4, "tiled") as et, _engine(shipped_luts, 2, "sdy", 4, "generic") as eg:
        ot = et(d).cpu().numpy()
        og = eg(d).cpu().numpy()
        assert (ot == og).all()
        single = et(d[1]).cpu().numpy()
        assert (single == ot[1]).all()
        crop = np.ascontiguousarray(frames[0, 100:200, 300:420])
        oc = et(torch.from_numpy(crop).cuda()).cpu().numpy()
        # receptive field is 9x9 (2 stages): interior of the crop (4 LR px in) must agree
        assert (oc[16:-16, 16:-16] == ot[0, 400 + 16:800 - 16, 1200 + 16:1680 - 16]).all()
    spot = CO.sr_u8(frames[0, :64, :64], shipped_luts, 2, "sdy", 4)
    assert (spot[:-16, :-16] == ot[0, :256 - 16, :256 - 16]).all()


def test_full_size_x4_960x540_matches_c_oracle(shipped_luts):
    """BASELINE config 3 at full size: one 960x540 RGB frame -> 4K with the reference's shipped x4 LUTs,
    bit-exact against the C oracle for AUTO and for every kernel family."""
    import torch
    from mulut_b200.infer import LutEngine
    rng = np.random.default_rng(33)
    frame = rng.integers(0, 256, (1, 540, 960, 3), dtype=np.uint8)
    ref = CO.sr_u8(frame, shipped_luts, 2, "sdy", 4)
    assert ref.shape == (1, 2160, 3840, 3)
    d = torch.from_numpy(frame).cuda()
    for kernel in [-1] + sorted(set(KERNELS.values())):
        with LutEngine(shipped_luts, 2, "sdy", 4, 4, device=0, kernel=kernel) as eng:
            out = eng(d).cpu().numpy()
        assert (out == ref).all(), (kernel, int((out != ref).sum()))


def test_full_size_8k_x2_matches_c_oracle():
    """BASELINE config 5 at full size: one 7680x4320 RGB frame -> 15360x8640 (398 MB), AUTO policy,
    bit-exact against the C oracle; also through the host path."""
    import torch
    from mulut_b200.infer import LutEngine, pinned_empty
    rng = np.random.default_rng(55)
    luts = O.random_luts(1, 2, "sdy", 2)
    frame = rng.integers(0, 256, (1, 4320, 7680, 3), dtype=np.uint8)
    ref = CO.sr_u8(frame, luts, 2, "sdy", 2)
    with LutEngine(luts, 2, "sdy", 2, 4, device=0) as eng:
        eng.profile(True)
        out = eng(torch.from_numpy(frame).cuda()).cpu().numpy()
        prof = eng.profile_read()
        eng.profile(False)
        assert "last_binned" in prof and ("fused_stage" in prof or "smem_stage" in prof), prof      # AUTO took the TMA-fed kernels
        assert out.shape == (1, 8640, 15360, 3)
        assert (out == ref).all(), int((out != ref).sum())
        del out
        hout = pinned_empty(ref.shape)
        eng.infer_host(frame, hout)
        assert (hout == ref).all()
        del hout
        # strip sharding of the single frame: 2 and 3 row strips with a 4-row halo, computed
        # independently (as 2 / 3 GPUs would), concatenate to the whole-frame bytes
        d = torch.from_numpy(frame[0]).cuda()
        for world in (2, 3):
            rows = []
            for rank in range(world):
                b, e, strip = eng.infer_strip(d, rank, world)
                assert strip.shape == ((e - b) * 2, 15360, 3)
                rows.append(strip.cpu().numpy())
            assert (np.concatenate(rows, 0) == ref[0]).all(), world


@pytest.mark.parametrize("scale,stages,H", [(2, 2, 37), (4, 2, 9), (2, 3, 23), (2, 1, 5), (3, 2, 2)])
def test_strip_sharding_small_frames(scale, stages, H):
    """infer_strip on frames where strips are thinner than the halo, world > rows, and host arrays."""
    import torch
    from mulut_b200.infer import LutEngine
    rng = np.random.default_rng(H)
    luts = O.random_luts(60 + H, stages, "sdy", scale)
    img = rng.integers(0, 256, (H, 40, 3), dtype=np.uint8)
    ref = CO.sr_u8(img, luts, stages, "sdy", scale)
    with LutEngine(luts, stages, "sdy", scale, 4, device=0) as eng:
        for world in (1, 2, 3, 8):
            dev = [eng.infer_strip(torch.from_numpy(img).cuda(), r, world)[2].cpu().numpy() for r in range(world)]
            assert (np.concatenate(dev, 0) == ref).all(), world
            host = [eng.infer_strip(img, r, world)[2] for r in range(world)]
            assert (np.concatenate(host, 0) == ref).all(), world


def test_cfg1_batch_of_64_256x256_x4_matches_c_oracle(shipped_luts):
    """BASELINE config 1 (the reference's CPU-runnable case) as bench.py times it: a 64-frame batch of
    256x256 RGB frames, shipped x4 LUTs, AUTO policy."""
    import torch
    from mulut_b200.infer import LutEngine
    rng = np.random.default_rng(11)
    frames = rng.integers(0, 256, (64, 256, 256, 3), dtype=np.uint8)
    ref = CO.sr_u8(frames, shipped_luts, 2, "sdy", 4)
    with LutEngine(shipped_luts, 2, "sdy", 4, 4, device=0) as eng:
        out = eng(torch.from_numpy(frames).cuda()).cpu().numpy()
        assert (out == ref).all(), int((out != ref).sum())
        assert (eng(frames) == ref).all()


@pytest.mark.parametrize("kernel", [-1, 1])
def test_one_handle_serves_concurrent_streams(kernel):
    """concurrent calls on different streams of ONE handle (the LUTs are read-only; every stream
    gets its own intermediates).  Different frames on three streams, interleaved without synchronising."""
    import torch
    from mulut_b200.infer import LutEngine
    rng = np.random.default_rng(17)
    luts = O.random_luts(18, 2, "sdy", 2)
    frames = [rng.integers(0, 256, (2, 400, 480, 3), dtype=np.uint8) for _ in range(3)]
    refs = [CO.sr_u8(f, luts, 2, "sdy", 2) for f in frames]
    with LutEngine(luts, 2, "sdy", 2, 4, device=0, kernel=kernel) as eng:
        streams = [torch.cuda.Stream() for _ in range(3)]
        d_in = [torch.from_numpy(f).cuda() for f in frames]
        torch.cuda.synchronize()
        outs = [None] * 3
        for rep in range(4):
            for i, st in enumerate(streams):
                with torch.cuda.stream(st):
                    outs[i] = eng(d_in[i])
        torch.cuda.synchronize()
        for i in range(3):
            assert (outs[i].cpu().numpy() == refs[i]).all(), i
        # a ninth distinct stream is refused with a clear error instead of sharing a workspace
        extra = [torch.cuda.Stream() for _ in range(8)]
        with pytest.raises(ValueError, match="distinct streams"):
            for st in extra:
                with torch.cuda.stream(st):
                    eng(d_in[0][:, :8, :16])
        torch.cuda.synchronize()


def test_call_compatible_single_pass(pass_cases):
    from mulut_b200.infer import FourSimplexInterpFaster
    meta, data = pass_cases
    for c in meta:
        lut = np.random.default_rng(c["lut_seed"]).integers(-127, 128, (83521, c["up"] ** 2), dtype=np.int8)
        out = FourSimplexInterpFaster(lut.astype(np.float32), data["x_%d" % c["i"]], c["h"], c["w"], 4, c["rot"],
                                      upscale=c["up"], mode=c["mode"])
        ref = data["out_%d" % c["i"]]
        assert out.dtype == np.float64 and out.shape == ref.shape
        assert (out == ref).all(), c


def test_error_mapping():
    from mulut_b200.infer import FourSimplexInterpFaster, LutEngine
    luts = O.random_luts(3, 1, "s", 2)
    with pytest.raises(ValueError, match="Mode e not implemented."):
        LutEngine({"s1_e": luts["s1_s"]}, 1, "e", 2)
    with pytest.raises(ValueError, match="Mode q not implemented."):
        FourSimplexInterpFaster(np.zeros((83521, 1), np.float32), np.zeros((1, 3, 3), np.float32), 2, 2, 4, 1, 1, "q")
    with pytest.raises(IndexError):
        LutEngine({"s1_s": luts["s1_s"][:1000]}, 1, "s", 2)
    with pytest.raises(ValueError):
        LutEngine(luts, 1, "s", 5)
    with LutEngine(luts, 1, "s", 2) as eng:
        out = eng(np.zeros((0, 8, 8, 3), np.uint8))
        assert out.shape == (0, 16, 16, 3)
        g = eng(np.random.default_rng(0).integers(0, 256, (6, 7), dtype=np.uint8))   # grey -> 3 channels
        assert g.shape == (12, 14, 3) and (g[..., 0] == g[..., 1]).all()
