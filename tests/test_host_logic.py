"""CPU-side tests: the C ABI library loads and exports every symbol the header
declares, file-name rules, options, sharding, the flat gradient bucket and the
multi-process (gloo, world_size 2) plumbing."""
import os
import re
import socket

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    txt = open(os.path.join(ROOT, "include", "mulut.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mulut_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from mulut_b200 import _lib
    names = _header_functions()
    assert len(names) >= 14
    assert sorted(_lib.SYMBOLS) == names            # binding table == header
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), n
    L = _lib.lib()
    assert L.mulut_version() >= 100
    assert isinstance(_lib.last_error(), str)


def test_header_cites_reference_for_each_entry_point():
    txt = open(os.path.join(ROOT, "include", "mulut.h")).read()
    for cite in ("sr/4_test_lut.py:279-306", "sr/4_test_lut.py:14-237", "sr/model.py:69-287",
                 "sr/4_test_lut.py:323-333"):
        assert cite in txt, cite


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_product_fails_loudly_without_gpu(shipped_luts):
    from mulut_b200 import _lib
    from mulut_b200.infer import LutEngine
    with pytest.raises(_lib.MulutError, match="CUDA error"):
        LutEngine(shipped_luts, 2, "sdy", 4)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mulut_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("# oracle-free", ""), os.path.join(dp, f)


def test_tools_never_import_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline / parity legs may execute the oracle tree:
    the helper scripts under tools/ must not (those that need a checker live under tests/)."""
    import re
    tools = os.path.join(ROOT, "tools")
    for dp, _, files in os.walk(tools):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(dp, f)


def test_nonlast_epilogue_without_clamps_is_round_half_even():
    """The integer form K1b uses for non-last stages (common.cuh: rhe_div_nonneg_magic): u = t + den/2,
    q' = mulhi(u, magic), minus one on an odd tie - against numpy's round-half-even, every t in [0, 254 den]."""
    for M in range(1, 9):
        D = 64 * M
        t = np.arange(0, 254 * D + 1, dtype=np.int64)
        magic = 0xFFFFFFFF // D + 1
        u = t + D // 2
        q = (u * magic) >> 32
        r = u - q * D
        res = q - ((r == 0) & (q & 1))
        assert (res == np.round(t / D)).all() and res.max() <= 254, M


def test_lut_naming_rules(gold_dir):
    from mulut_b200.infer import load_luts, lut_path
    # test path: 8 - interval (4_test_lut.py:331-332)
    assert os.path.basename(lut_path("e", "LUT_ft", 4, 4, 2, "y")) == "LUT_ft_x4_4bit_int8_s2_y.npy"
    assert os.path.basename(lut_path("e", "LUT_ft", 2, 5, 1, "s")) == "LUT_ft_x2_3bit_int8_s1_s.npy"
    luts = load_luts(os.path.join(gold_dir, "luts_x4"), 2, "sdy", 4, 4, "LUT_ft")
    assert sorted(luts) == ["s1_d", "s1_s", "s1_y", "s2_d", "s2_s", "s2_y"]
    assert luts["s1_s"].shape == (83521, 1) and luts["s2_d"].shape == (83521, 16) and luts["s2_d"].dtype == np.int8
    with pytest.raises(FileNotFoundError):
        load_luts(os.path.join(gold_dir, "luts_x4"), 2, "sdy", 2, 4, "LUT_ft")


def test_options_match_reference_flags():
    from mulut_b200.options import TestOptions, TrainOptions
    o = TestOptions().parse(["--stages", "2", "--modes", "sdy", "-e", "../models/sr_x2sdy"])
    assert (o.stages, o.modes, o.expDir, o.scale, o.interval, o.lutName) == (2, "sdy", "../models/sr_x2sdy", 4, 4, "LUT_ft")
    assert (o.testDir, o.resultRoot, o.loadIter, o.isTrain) == ("../data/SRBenchmark", "../results", 200000, False)
    o = TestOptions().parse(["-r", "2", "--interval", "5", "-i", "7"])
    assert (o.scale, o.interval, o.loadIter) == (2, 5, 7) and o.expDir.startswith("../models/debug/expr_")
    t = TrainOptions().parse(["-e", "x", "--batchSize", "256", "-g", "8"])
    assert (t.batchSize, t.cropSize, t.totalIter, t.lr0, t.lr1, t.gpuNum, t.isTrain) == (256, 48, 200000, 1e-3, 1e-4, 8, True)


def test_shard_range_partitions_exactly():
    from mulut_b200.dist import shard_range, shard_rows_with_halo
    for n in (0, 1, 7, 8, 64, 1001):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    b, e, lb, le = shard_rows_with_halo(1080, 1, 4, 4)
    assert (b, e, lb, le) == (270, 540, 266, 544)
    assert shard_rows_with_halo(1080, 0, 4, 4)[2] == 0 and shard_rows_with_halo(1080, 3, 4, 4)[3] == 1080
    with pytest.raises(ValueError):
        shard_range(4, 4, 4)


def test_mulut_module_parameters_and_export(tmp_path, shipped_luts):
    from mulut_b200.model import MuLUT
    net = MuLUT(None, 2, ["s", "d", "y"], upscale=4, interval=4, luts=shipped_luts)
    names = [n for n, _ in net.named_parameters()]
    assert names == ["weight_s1_s", "weight_s1_d", "weight_s1_y", "weight_s2_s", "weight_s2_d", "weight_s2_y"]
    assert sum(p.numel() for p in net.parameters()) == 4259571      # 17.04 MB fp32
    out = net.export_luts(str(tmp_path))
    for k, v in shipped_luts.items():
        assert (out[k] == v).all()                                   # int8/127 -> round(.*127) round trip
        assert (np.load(tmp_path / "LUT_ft_x4_4bit_int8_{}.npy".format(k)) == v).all()
    # loading by the finetune naming rule (model.py:53-54: `interval`, not 8-interval)
    for k, v in shipped_luts.items():
        np.save(tmp_path / "LUT_x4_4bit_int8_{}.npy".format(k), v)
    net2 = MuLUT(str(tmp_path), 2, ["s", "d", "y"], upscale=4, interval=4)
    assert torch.equal(net2.weight_s2_y, net.weight_s2_y)
    x = torch.tensor([0.4, 0.5, 1.5, 2.5, -0.5])
    assert torch.equal(MuLUT.round_func(x), torch.round(x))


def test_lr_schedule_matches_reference_formula():
    from mulut_b200.cli.finetune_lut import lr_lambda
    f = lr_lambda(2000, 1e-3, 1e-4)
    assert abs(f(0) - 1.0) < 1e-12 and abs(f(2000) - 0.1) < 1e-12 and abs(f(1000) - 0.55) < 1e-12
    g = lr_lambda(100, 1e-3, -1)
    assert abs(g(100) - 0.2) < 1e-12


def test_flat_grad_bucket_single_process():
    from mulut_b200.dist import FlatGradBucket
    ps = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(7))]
    b = FlatGradBucket(ps)
    (ps[0].sum() * 2 + (ps[1] * 3).sum()).backward()
    # every view starts on a 16-byte boundary (4 floats): 15 -> 16, 7 -> 8
    assert b.flat.numel() == 24 and b.offsets == [0, 16]
    assert torch.equal(b.flat[:15], torch.full((15,), 2.0)) and torch.equal(b.flat[16:23], torch.full((7,), 3.0))
    assert float(b.flat[15]) == 0 and float(b.flat[23]) == 0
    b.zero_()
    assert float(b.flat.abs().sum()) == 0 and ps[0].grad.data_ptr() == b.flat.data_ptr()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    from mulut_b200.dist import FlatGradBucket, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    w = [torch.nn.Parameter(torch.randn(11, 4)), torch.nn.Parameter(torch.randn(6))]
    data = torch.arange(40, dtype=torch.float32).reshape(10, 4)
    b, e = shard_range(10, rank, world)
    bucket = FlatGradBucket(w)
    x = data[b:e]
    # per-rank MEAN loss over an equal share, like F.mse_loss on the rank's batch
    loss = ((x @ w[0][:4] ).pow(2).mean() + w[1].sum() * x.mean())
    loss.backward()
    whole = bucket.flat.clone()
    bucket.all_reduce_mean()
    reduced = bucket.flat.clone()
    # the same reduction started range by range (what a step does under its backward pass): second table first,
    # then the first one is the gap all_reduce_mean fills
    bucket.flat.copy_(whole)
    assert bucket.range_of([w[1]]) == (bucket.offsets[1], bucket.flat.numel())
    assert bucket.range_of(w) == (0, bucket.flat.numel())
    bucket.begin_range(*bucket.range_of([w[1]]))
    bucket.all_reduce_mean()
    assert torch.equal(bucket.flat, reduced) and not bucket._pending
    q.put((rank, bucket.packed().clone().numpy(), (b, e)))
    dist.destroy_process_group()


def test_gloo_world2_allreduce_equals_full_batch():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    [p.join(60) for p in procs]
    assert res[0][2] == (0, 5) and res[1][2] == (5, 10)
    assert np.allclose(res[0][1], res[1][1])
    # single-process reference on the whole batch (equal shares => mean of means == global mean)
    torch.manual_seed(0)
    w = [torch.nn.Parameter(torch.randn(11, 4)), torch.nn.Parameter(torch.randn(6))]
    data = torch.arange(40, dtype=torch.float32).reshape(10, 4)
    loss = ((data @ w[0][:4]).pow(2).mean() + w[1].sum() * data.mean())
    loss.backward()
    full = torch.cat([w[0].grad.reshape(-1), w[1].grad.reshape(-1)]).numpy()
    assert np.allclose(res[0][1], full, rtol=1e-5, atol=1e-5)


def test_metrics_known_values():
    from mulut_b200.metrics import PSNR, cal_ssim, modcrop, rgb2ycbcr
    a = np.full((32, 32), 100.0)
    b = a + 5.0
    assert abs(PSNR(a, b, 4) - 20 * np.log10(255 / 5)) < 1e-4
    assert abs(cal_ssim(a, a) - 1.0) < 1e-12
    assert modcrop(np.zeros((10, 11, 3)), 4).shape == (8, 8, 3)
    y = rgb2ycbcr(np.array([[[255, 255, 255], [0, 0, 0]]], dtype=np.uint8))[..., 0]
    assert np.allclose(y, [[235.0, 16.0]])


def test_transfer_grid_enumeration_and_round_trip(tmp_path):
    """2_transfer_to_lut.py's grid: row a*L^3+b*L^2+c*L+d is the patch [[a,b],[c,d]] with the last grid
    point at 255; a 'network' that reads a table at the grid points reproduces the table (the
    enumeration order is the retrieval kernels' row order), file names and int8 quantisation as the
    reference writes them."""
    from mulut_b200 import transfer as T
    x = T.get_input_tensor(4)
    L = 17
    assert x.shape == (L ** 4, 1, 2, 2) and x.dtype == torch.float32
    base = np.concatenate([np.arange(0, 256, 16), [255]])
    for idx in (0, 1, 16, 17, 5000, 83520, 12345):
        a, r = divmod(idx, L ** 3); b, r = divmod(r, L ** 2); c, d = divmod(r, L)
        assert np.allclose(x[idx, 0].numpy() * 255.0, [[base[a], base[b]], [base[c], base[d]]]), idx
    for mode, pos in (("d", ((0, 0), (0, 2), (2, 0), (2, 2))), ("y", ((0, 0), (1, 1), (1, 2), (2, 1)))):
        xm = T.get_mode_input_tensor(x[:1000], mode)
        assert xm.shape == (1000, 1, 3, 3)
        for k, (yy, xx) in enumerate(pos):
            assert torch.equal(xm[:, 0, yy, xx], x[:1000, 0, k // 2, k % 2])
        assert float(xm.sum()) == pytest.approx(float(x[:1000].sum()))
    with pytest.raises(ValueError, match="Mode e not implemented."):
        T.get_mode_input_tensor(x[:4], "e")
    # round trip through a table-reading "network"
    rng = np.random.default_rng(0)
    table = rng.integers(-127, 128, (L ** 4, 4)).astype(np.int8)

    def net_fn(batch, stage, mode):
        taps = batch.reshape(batch.shape[0], -1) * 255.0
        taps = taps[:, [0, 1, 2, 3]] if mode == "s" else taps[:, [0, 2, 6, 8]]
        gi = torch.where(taps >= 255, torch.tensor(16.0), taps / 16).round().long()
        rows = ((gi[:, 0] * L + gi[:, 1]) * L + gi[:, 2]) * L + gi[:, 3]
        return torch.from_numpy(table.astype(np.float32))[rows].reshape(-1, 1, 2, 2) / 127.0

    out = T.transfer_to_lut(net_fn, 1, "sd", 2, 4, exp_dir=str(tmp_path))
    for m in "sd":
        assert out["s1_" + m].dtype == np.int8 and out["s1_" + m].shape == (L ** 4, 1, 2, 2)
        assert (out["s1_" + m].reshape(-1, 4) == table).all()
        assert (np.load(tmp_path / "LUT_x2_4bit_int8_s1_{}.npy".format(m)) == out["s1_" + m]).all()


def _plan(hist, n_tiles, G=148, cap=None, orphans=1):
    import ctypes
    from mulut_b200 import _lib
    h = (ctypes.c_ulonglong * 8)(*[int(x) for x in hist])
    g = (ctypes.c_int * 8)()
    mask = ctypes.c_uint()
    cap = sum(hist) // 4 + 1024 if cap is None else cap
    assert _lib.lib().mulut_plan_bins(h, n_tiles, G, cap, orphans, g, ctypes.byref(mask)) == 0
    return list(g), mask.value


def test_binned_kernel_plan_properties():
    """K1f's CTA allocation (host mirror of the device code): every CTA is dealt, empty bins get
    none, sparse bins become orphans only when allowed and only within the list's capacity, dense
    bins get CTAs roughly in proportion to their samples."""
    n_tiles, tile = 32640, 3072
    total = n_tiles * tile
    # the benchmark's stage-2 input: two dense bins, two at 0.5 %
    hist = [0, 0, int(.005 * total), int(.505 * total), int(.485 * total), int(.005 * total), 0, 0]
    g, mask = _plan(hist, n_tiles)
    assert sum(g) == 148 and mask == 0b00100100 and g[2] == g[5] == 0 and g[0] == g[7] == 0
    assert abs(g[3] - g[4]) <= 4 and g[3] + g[4] == 148
    g, mask = _plan(hist, n_tiles, orphans=0)                   # orphaning off: the sparse bins get CTAs
    assert mask == 0 and sum(g) == 148 and g[2] >= 1 and g[5] >= 1 and g[3] > 5 * g[2]
    # uniform: all eight bins resident and equal
    g, mask = _plan([total // 8] * 8, n_tiles)
    assert mask == 0 and sum(g) == 148 and max(g) - min(g) <= 1
    # a single populated bin takes everything; an empty input takes nothing
    g, mask = _plan([0, 0, 0, 0, 0, 0, total, 0], n_tiles)
    assert g == [0, 0, 0, 0, 0, 0, 148, 0] and mask == 0
    g, mask = _plan([0] * 8, n_tiles)
    assert g == [0] * 8 and mask == 0
    # tiny frame: everything is cheaper through the list ... unless the list cannot hold it
    g, mask = _plan([10, 0, 0, 5, 0, 0, 0, 1], 1)
    assert g == [0] * 8 and mask == 0b10001001
    g, mask = _plan([10, 0, 0, 5, 0, 0, 0, 1], 1, cap=4)
    assert mask == 0b10000000 and g[0] >= 1 and g[3] >= 1 and g[7] == 0 and sum(g) == 148
    # more non-empty bins than CTAs to floor-share: still exactly G in total
    rng = np.random.default_rng(0)
    for _ in range(50):
        hist = [int(x) for x in rng.integers(0, 10 ** int(rng.integers(1, 9)), 8)]
        G = int(rng.integers(8, 200))
        g, mask = _plan(hist, int(rng.integers(1, 50000)), G=G, orphans=int(rng.integers(0, 2)))
        resident = [b for b in range(8) if hist[b] and not (mask >> b) & 1]
        assert all((g[b] > 0) == (b in resident) for b in range(8)), (hist, g, mask)
        assert sum(g) == (G if resident else 0), (hist, g, mask)


def test_new_entry_points_validate_arguments_without_a_gpu():
    """Argument checks of the K4 / Adam / metrics entry points return before any CUDA call, so they
    are testable here: bad mode -> E_BAD_MODE with the reference's message, null buffers -> E_BAD_ARG."""
    import ctypes
    from mulut_b200 import _lib
    L = _lib.lib()
    # quantised rows + clamp flags per mode, the statistics tail, three scratch copies of the gradient tables
    assert L.mulut_stage_workspace_bytes(3, 83521, 4) == (3 * (83521 * 16 + ((83521 * 2 + 15) // 16) * 16) + 16 +
                                                          3 * 3 * 83521 * 16 * 4)
    assert L.mulut_stage_workspace_bytes(3, 83521, 1) > 0 and L.mulut_stage_workspace_bytes(0, 83521, 4) == 0
    one = ctypes.c_void_p(16)                  # a non-null, 16-byte aligned dummy: never dereferenced on these paths
    ptrs = (ctypes.c_void_p * 1)(16)
    rc = L.mulut_stage_fwd_f32(ptrs, 1, b"e", 83521, 1, 4, one, 1, 1, 4, 4, 4.0, 127.0, one, one, one, None)
    assert rc == _lib.E_BAD_MODE and "Mode e not implemented." in _lib.last_error()
    rc = L.mulut_stage_fwd_f32(ptrs, 1, b"s", 100, 1, 4, one, 1, 1, 4, 4, 4.0, 127.0, one, one, one, None)
    assert rc == _lib.E_LUT_SMALL
    rc = L.mulut_stage_fwd_f32(ptrs, 1, b"s", 83521, 1, 4, one, 1, 1, 4, 4, 4.0, 127.0, one, one, None, None)
    assert rc == _lib.E_BAD_ARG and "workspace" in _lib.last_error()
    rc = L.mulut_stage_bwd_f32(ptrs, 1, b"s", 83521, 5, 4, one, 1, 1, 4, 4, 4.0, 127.0, one, one, one, ptrs, None, None)
    assert rc == _lib.E_BAD_ARG
    rc = L.mulut_adam_step_f32(None, one, one, one, 10, one, 0.9, 0.999, 1e-8, 0.0, one, None)
    assert rc == _lib.E_BAD_ARG
    out = (ctypes.c_double * 2)()
    rc = L.mulut_eval_psnr_ssim_y_u8(one, one, 0, 8, 4, one, out, None)
    assert rc == _lib.E_BAD_ARG
    with pytest.raises(ValueError):
        _lib.check(rc)


def test_paired_threshold_form_equals_sorted_simplex_weights():
    """K1h (csrc/infer_stage1.cu) interpolates with a 3-key sort and a-paired table entries: vertex
    (u_j, a not stepped) weighs (g_j - g_{j+1}) - (c_j - c_{j+1}) and (u_j, a stepped) c_j - c_{j+1},
    c_j = min(g_j, f_a).  Exhaustively over all 16^4 fraction tuples this puts the same weight on every
    corner of the 4-D cell as the reference's sorted-fraction form (sr/4_test_lut.py:53-237; ties are
    zero-width intervals, so the tie order cannot matter)."""
    f = np.stack(np.meshgrid(*[np.arange(16)] * 4, indexing="ij"), -1).reshape(-1, 4)       # (65536, 4): fa fb fc fd
    n = f.shape[0]
    rows = np.arange(n)
    # reference form: sort the four fractions descending (stable), walk the vertex chain
    order = np.argsort(-f, axis=1, kind="stable")
    fs = np.take_along_axis(f, order, 1)
    w_ref = np.zeros((n, 16), dtype=np.int64)                 # weight per corner, corner = bitmask of stepped axes
    corner = np.zeros(n, dtype=np.int64)
    w = np.concatenate([16 - fs[:, :1], fs[:, :-1] - fs[:, 1:], fs[:, 3:]], 1)
    for k in range(5):
        np.add.at(w_ref, (rows, corner), w[:, k])
        if k < 4:
            corner = corner | (1 << order[:, k])
    # K1h form: sort b, c, d only; a enters through c_j = min(g_j, fa)
    fa = f[:, 0]
    o3 = np.argsort(-f[:, 1:], axis=1, kind="stable")
    g = np.take_along_axis(f[:, 1:], o3, 1)
    gj = np.concatenate([np.full((n, 1), 16), g, np.zeros((n, 1), dtype=g.dtype)], 1)      # g0..g4
    cj = np.minimum(gj, fa[:, None])
    q = gj + 255 * cj                                           # the kernel's packed Q_j = g_j + 255 c_j
    wp = q[:, :-1] - q[:, 1:]                                   # alpha | beta << 8
    alpha, beta = wp & 255, wp >> 8
    assert (alpha >= 0).all() and (alpha <= 16).all() and (beta >= 0).all() and (beta <= 16).all()
    w_new = np.zeros((n, 16), dtype=np.int64)
    corner = np.zeros(n, dtype=np.int64)
    for j in range(4):
        np.add.at(w_new, (rows, corner), alpha[:, j])
        np.add.at(w_new, (rows, corner | 1), beta[:, j])        # bit 0 = axis a stepped
        if j < 3:
            corner = corner | (1 << (o3[:, j] + 1))
    assert (w_ref.sum(1) == 16).all()
    assert (w_new == w_ref).all()


def test_header_is_plain_c99():
    """The drop-in boundary is a C ABI: include/mulut.h must compile as C (no C++ types in the signatures)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c",
                        os.path.join(ROOT, "include", "mulut.h")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_private_backward_halves_cover_every_row_a_pixel_touches():
    """K4p (interp_f32.cu) keeps slabs 0..8 or 8..16 of a x1 table in shared memory and deals the pixels by their
    centre tap (< 128 / >= 128).  Whatever the other three taps are, all five vertices of the simplex must then
    lie in that half - checked on the oracle's vertex arithmetic (oracle/mulut_oracle.py:simplex_vertices) for
    every centre value at the extreme and at random neighbour values."""
    from oracle import mulut_oracle as O
    L = 17
    slab, rows_half = L ** 3, 9 * L ** 3
    rng = np.random.default_rng(0)
    others = np.concatenate([np.array([[0, 0, 0], [255, 255, 255], [0, 255, 0], [255, 0, 255], [15, 16, 240]]),
                             rng.integers(0, 256, (200, 3))])
    for ta in range(256):
        half_base = 8 * slab if ta >= 128 else 0
        t = np.concatenate([np.full((len(others), 1), ta), others], axis=1).T        # (4, n)
        verts, w, _ = O.simplex_vertices(t, 4)
        assert verts.min() >= half_base and verts.max() < half_base + rows_half, ta
        assert verts.max() < L ** 4 and (w.sum(axis=0) == 16).all()
    # the extremes fill the halves exactly: the lower one ends on its last row, the upper one starts on its first
    assert O.simplex_vertices(np.array([[127], [255], [255], [255]]), 4)[0].max() == rows_half - 1
    assert O.simplex_vertices(np.array([[128], [0], [0], [0]]), 4)[0].min() == 8 * slab
