#!/usr/bin/env python
"""bench.py — headline benchmark of the MuLUT LUT-retrieval hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): output Mpix/s of sr_x2sdy 2-stage LUT inference,
1920x1080 RGB frames -> 3840x2160 (scale 2, modes s,d,y, interval 4), seeded
random int8 LUTs (no x2 LUT is shipped by the reference), synthetic uniform
uint8 frames.  One "step" = one pass of the path over --frames frames per GPU.
Frames are sharded across ranks with no collective (weak scaling).

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, C, SCALE, STAGES, MODES, INTERVAL = 1080, 1920, 3, 2, 2, "sdy", 4
WORKLOAD = "cfg2: sr_x2sdy 2-stage, 1920x1080x3 uint8 -> 3840x2160x3, scale 2, random int8 LUTs (seed 1)"
# algorithmic bytes per input sample (SURVEY.md 8d / DESIGN.md): 12*5 vertex rows of 1 B in
# stage 1, 12*5 rows of r^2 = 4 B in stage 2; HBM: 1 B read + r^2 written.
GATHER_B_STAGE1, GATHER_B_STAGE2, HBM_B = 60, 240, 5
SHIPPED_LUTS = False

# The headline (and the default) is cfg2; the other BASELINE.json configs can be timed with --config.
CONFIGS = {
    "cfg2": dict(H=1080, W=1920, scale=2, shipped=False, frames=16,
                 name="cfg2: sr_x2sdy 2-stage, 1920x1080x3 uint8 -> 3840x2160x3, scale 2, random int8 LUTs (seed 1)"),
    "cfg1": dict(H=256, W=256, scale=4, shipped=True, frames=64,
                 name="cfg1: sr_x2sdy 2-stage, 256x256x3 -> 1024x1024x3, scale 4, shipped LUTs"),
    "cfg3": dict(H=540, W=960, scale=4, shipped=True, frames=16,
                 name="cfg3: sr_x4sdy 2-stage, 960x540x3 -> 3840x2160x3, scale 4, shipped LUTs"),
    # calibration workload (not a BASELINE config): the x2 last stage ALONE on uniform frames,
    # i.e. all eight value bins equally populated - the worst case of the binned kernel
    "x2s1": dict(H=1080, W=1920, scale=2, shipped=False, frames=16, stages=1,
                 name="x2s1: ONE-stage sr_x2sdy, 1920x1080x3 uniform uint8 -> 3840x2160x3, random int8 LUTs"),
    "cfg5": dict(H=4320, W=7680, scale=2, shipped=False, frames=2,
                 name="cfg5: sr_x2sdy 2-stage, 7680x4320x3 -> 15360x8640x3, scale 2, random int8 LUTs (seed 1)"),
}


def select_config(name):
    global H, W, SCALE, WORKLOAD, GATHER_B_STAGE2, HBM_B, SHIPPED_LUTS, STAGES
    c = CONFIGS[name]
    STAGES = c.get("stages", 2)
    H, W, SCALE, WORKLOAD, SHIPPED_LUTS = c["H"], c["W"], c["scale"], c["name"], c["shipped"]
    GATHER_B_STAGE2, HBM_B = 60 * SCALE * SCALE, 1 + SCALE * SCALE
    return c["frames"]


def make_luts(seed=1):
    if SHIPPED_LUTS:                       # the reference's x4 tables (tests/golden/luts_x4)
        d = os.path.join(ROOT, "tests", "golden", "luts_x4")
        return {"s{}_{}".format(s, m): np.load(os.path.join(d, "LUT_ft_x4_4bit_int8_s{}_{}.npy".format(s, m))).reshape(
            -1, 1 if s == 1 else 16) for s in (1, 2) for m in "sdy"}
    rng = np.random.default_rng(seed)
    luts = {}
    for s in range(STAGES):
        cols = SCALE * SCALE if s + 1 == STAGES else 1
        for m in MODES:
            luts["s{}_{}".format(s + 1, m)] = rng.integers(-127, 128, (83521, cols), dtype=np.int8)
    return luts


DATA = "uniform"


def make_frames(n, seed):
    """uniform: i.i.d. uint8 (worst case: every sample in a different LUT cell).  natural: mirror-tiled
    crops of the reference's golden Set5 results (tests/golden/set5_x4.npz) - smooth regions, edges,
    a natural value histogram (SURVEY.md 8d asks for this set to be reported separately)."""
    rng = np.random.default_rng(seed)
    if DATA == "uniform":
        return rng.integers(0, 256, (n, H, W, C), dtype=np.uint8)
    d = np.load(os.path.join(ROOT, "tests", "golden", "set5_x4.npz"))
    srcs = [d[k] for k in d.files if k.startswith("sr_")]
    out = np.empty((n, H, W, C), np.uint8)
    for i in range(n):
        img = srcs[(i + seed) % len(srcs)]
        img = np.concatenate([img, img[::-1]], 0)
        img = np.concatenate([img, img[:, ::-1]], 1)           # mirror tile: seamless when repeated
        oy, ox = rng.integers(0, img.shape[0]), rng.integers(0, img.shape[1])
        ys = (np.arange(H) + oy) % img.shape[0]
        xs = (np.arange(W) + ox) % img.shape[1]
        out[i] = img[ys][:, xs]
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "MEASURED_PEAKS.json"
    return 6650.0, "fallback (B200_PROFILING.md)"


def gather_peak(device):
    """Measured denominators of the gather roofline (mulut_gather_bench, live, ~0.5 s): independent random
    gathers of one shape from one memory level on all SMs.  ONE fixed peak per primitive: the best of the
    launch shapes tried (not the shape of the kernel that is compared with it)."""
    from mulut_b200 import _lib
    L = _lib.lib()
    out = (ctypes.c_double * 3)()
    res = {}
    for name, table, shapes in [("quad_cell64", 12 << 20, [(4, 512)]), ("ldg_u32", 1 << 20, [(4, 512)]),
                                ("lds_u8", 83584, [(2, 512), (1, 1024), (2, 384)]),
                                ("lds_u32", 176976, [(1, 768), (1, 1024)]),
                                ("quad_cell256_3rows", 50 << 20, [(2, 384)])]:
        for bps, tpb in shapes:
            rc = L.mulut_gather_bench(device, _lib.GB_VARIANTS[name], table, 256, bps, tpb, 3, out)
            if rc == 0 and out[0] > res.get(name, {}).get("gathers_per_s", 0.0):
                res[name] = {"gathers_per_s": out[0], "useful_GBps": out[1] / 1e9, "blocks_per_sm": bps, "threads": tpb}
    return res


def cpu_port_throughput(min_seconds, max_frames, threads=0):
    """The C oracle (oracle/mulut_oracle.c, kind 'port') on this host's cores."""
    from oracle import c_oracle as CO
    luts = make_luts()
    frame = make_frames(1, 1234)
    CO.sr_u8(frame[:, :64], luts, STAGES, MODES, SCALE, INTERVAL, threads)      # warm (build + page in)
    n, t0 = 0, time.perf_counter()
    while True:
        CO.sr_u8(frame, luts, STAGES, MODES, SCALE, INTERVAL, threads)
        n += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds or n >= max_frames:
            break
    cores = threads if threads > 0 else CO.max_threads()
    return n * H * W * SCALE * SCALE / dt / 1e6, cores, n, dt


def _numpy_ref_worker(job):
    """One pool worker = one 256x256 frame through the UNMODIFIED reference numpy path: the reference's own
    FourSimplexInterpFaster (sr/4_test_lut.py:14-237) driven by its _worker loop (:279-306, restated in
    oracle/ref_import.ref_pipeline), single-threaded like the reference's Pool workers (:257-259)."""
    cfg, seed, idx = job
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    select_config(cfg)
    from oracle import ref_import as R
    R.test_lut_module()                                      # import outside the timed part
    luts = make_luts()
    img = np.random.default_rng(seed + idx).integers(0, 256, (256, 256, C), dtype=np.uint8)
    t0 = time.perf_counter()
    out = R.ref_pipeline(img, luts, STAGES, list(MODES), SCALE, INTERVAL)
    dt = time.perf_counter() - t0
    return dt, (img, out) if idx == 0 else None


def numpy_reference_throughput(cfg, max_workers=0):
    """cpu_baseline.numpy_ref: the reference's numpy CPU path on this host, the reference's way - a
    multiprocessing.Pool with one 256x256 frame (BASELINE config 1's frame size) per worker, one worker per
    host core.  Cost per pixel does not depend on the frame size (every pass is elementwise over the
    pixels), so Mpix/s carries over to 1080p frames linearly; a 1080p frame would need ~460 s and 2.9 GB
    per worker."""
    import multiprocessing as mp
    from oracle import ref_import as R
    if not R.available():
        return {"unavailable": "oracle/_ref not staged (python -m oracle.fetch_ref in the build container)"}
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    if max_workers > 0:
        cores = min(cores, max_workers)
    ctx = mp.get_context("spawn")                            # this process holds a CUDA context: no fork
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_numpy_ref_worker, [(cfg, 4242, i) for i in range(cores)])
    wall = time.perf_counter() - t0
    per_frame = [r[0] for r in res]
    busy = max(per_frame)                                    # all workers start together: the slowest one bounds the batch
    img, out = res[0][1]
    from oracle import c_oracle as CO
    same = bool(np.array_equal(CO.sr_u8(img, make_luts(), STAGES, MODES, SCALE, INTERVAL, 1), out))
    pix = 256 * 256 * SCALE * SCALE
    return {"value": cores * pix / busy / 1e6, "unit": "Mpix/s", "cores": cores, "kind": "reference",
            "per_core_Mpix_s": pix / (sum(per_frame) / len(per_frame)) / 1e6,
            "sample": "{} x one 256x256x3 frame (scale {}), one per core under multiprocessing.Pool({}), slowest worker {:.1f} s "
                      "(pool wall {:.1f} s incl. spawn + imports); unmodified reference FourSimplexInterpFaster from "
                      "oracle/_ref".format(cores, SCALE, cores, busy, wall),
            "scaling_note": "per-pixel cost is frame-size independent: carries over to 1080p linearly",
            "c_oracle_agrees_on_this_frame": same}


def reference_module_step(luts, batch, crop, dev, steps=3):
    """finetune.reference_module: the reference's own training step (3_finetune_lut.py:129-136) with its own
    UNMODIFIED model.MuLUT (oracle/_ref/sr/model.py) on `dev` through ATen - the bar K4 is measured against."""
    import tempfile
    import torch
    import torch.nn.functional as F
    from oracle import ref_import as R
    if not R.available():
        return {"unavailable": "oracle/_ref not staged (python -m oracle.fetch_ref in the build container)"}
    model = R.model_module()
    with tempfile.TemporaryDirectory() as tmp:
        for k, v in luts.items():
            np.save(os.path.join(tmp, "LUT_x4_4bit_int8_{}.npy".format(k)), v)
        net = model.MuLUT(lut_folder=tmp, stages=2, modes=["s", "d", "y"], upscale=4, interval=4).to(dev)
    opt = torch.optim.Adam([p for p in net.parameters() if p.requires_grad], lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    from mulut_b200.cli.finetune_lut import synthetic_batch
    im, lb = synthetic_batch(batch, crop, 4, 1000, dev)

    def step():
        opt.zero_grad()
        loss = F.mse_loss(net(im), lb)
        loss.backward()
        opt.step()
        return loss

    on_gpu = torch.device(dev).type == "cuda"
    loss0 = float(step().item())                       # warm-up (allocator, cuDNN-free: plain ATen indexing kernels)
    if on_gpu:
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    if on_gpu:
        e0.record()
    for _ in range(steps):
        step()
    if on_gpu:
        e1.record()
        torch.cuda.synchronize(dev)
    wall = (time.perf_counter() - t0) / steps * 1e3
    return {"ms_per_step": e0.elapsed_time(e1) / steps if on_gpu else wall, "wall_ms_per_step": wall, "batch": batch,
            "steps": steps, "first_loss": loss0,
            "module": "unmodified reference model.MuLUT (sr/model.py) + torch.optim.Adam, eager ATen, " +
                      ("same GPU" if on_gpu else "host CPU, {} torch threads".format(torch.get_num_threads()))}


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's CPU path.  The reference is pure Python
    and /root/reference does not exist on the GPU box, so this is the C port of
    the same algorithm (oracle/, pinned bit-exact to the reference's golden
    outputs), on all host threads.  Each step = one 1080p frame."""
    if rank != 0:
        return
    from oracle import c_oracle as CO
    luts = make_luts()
    frame = make_frames(1, 0)
    cores = CO.max_threads()
    for _ in range(args.warmup):
        CO.sr_u8(frame, luts, STAGES, MODES, SCALE, INTERVAL, 0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        CO.sr_u8(frame, luts, STAGES, MODES, SCALE, INTERVAL, 0)
    dt = time.perf_counter() - t0
    val = args.steps * H * W * SCALE * SCALE / dt / 1e6
    sample = "{} steps x 1 frame 1920x1080 (C port of the reference algorithm, pthreads over rows)".format(args.steps)
    line = {
        "impl": "reference", "metric": "output Mpix/s", "value": val, "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": 1, "device": "cpu"},
        "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", type=str, default="mulut_b200", choices=["mulut_b200", "reference"])
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU per step (default: per config, 16 for cfg2)")
    ap.add_argument("--config", type=str, default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--kernel", type=str, default="auto", choices=["auto", "generic", "tiled", "quad", "cell", "binned"])
    ap.add_argument("--data", type=str, default="uniform", choices=["uniform", "natural"],
                    help="synthetic frame content: i.i.d. uniform bytes (default, worst case) or tiled natural crops")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-finetune", action="store_true", help="skip the cfg4 finetune block")
    ap.add_argument("--no-numpy-ref", action="store_true", help="skip cpu_baseline.numpy_ref (the reference's numpy path under a Pool)")
    ap.add_argument("--finetune-steps", type=int, default=30)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    global DATA
    DATA = args.data
    default_frames = select_config(args.config)
    if args.frames <= 0:
        args.frames = default_frames

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from mulut_b200 import _lib
    from mulut_b200.infer import LutEngine, pinned_empty

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the communicator comes up: park stdout on
        # stderr until then so that rank 0's stdout carries ONE JSON line and nothing else
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    kernel = {"auto": _lib.KERNEL_AUTO, "generic": _lib.KERNEL_GENERIC, "tiled": _lib.KERNEL_TILED,
              "quad": _lib.KERNEL_TILED_QUAD, "cell": _lib.KERNEL_TILED_CELL,
              "binned": _lib.KERNEL_TILED_BINNED}[args.kernel]
    luts = make_luts()
    eng = LutEngine(luts, STAGES, MODES, SCALE, INTERVAL, device=local, kernel=kernel)
    F = args.frames
    # cfg5 on several GPUs: ONE stream of huge frames, every frame strip-sharded over the ranks with a
    # 2-rows-per-stage halo (LutEngine.infer_strip's arithmetic) - total work fixed ("strong").
    # Every other config shards whole frames: each rank has its own F frames ("weak").
    strip = args.config == "cfg5" and world > 1
    H_full = H
    row_b, row_e, load_b, load_e = 0, H, 0, H
    if strip:
        from mulut_b200.dist import shard_rows_with_halo
        row_b, row_e, load_b, load_e = shard_rows_with_halo(H_full, rank, world, 2 * STAGES)
    HL = load_e - load_b                                     # input rows this rank reads
    host_in = pinned_empty((F, HL, W, C))
    if strip:
        whole = make_frames(F, seed=1000)                    # the same frames on every rank
        host_in[...] = whole[:, load_b:load_e]
        frame0_whole = whole[:1].copy() if rank == 0 else None
        del whole
    else:
        host_in[...] = make_frames(F, seed=1000 + rank)
    host_out = pinned_empty((F, HL * SCALE, W * SCALE, C))
    H_eff = HL                                               # the H the engine sees
    d_in = torch.from_numpy(np.ascontiguousarray(host_in)).cuda()
    d_out = torch.empty((F, HL * SCALE, W * SCALE, C), dtype=torch.uint8, device="cuda")
    eng.reserve(F, HL, W, C)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- device-resident leg (value) ----------------
    for _ in range(args.warmup):
        eng.infer_device(d_in, d_out)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.profile(True)
    launches0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        eng.infer_device(d_in, d_out)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count - launches0
    prof = eng.profile_read()
    eng.profile(False)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    out_pix_per_step = (1 if strip else world) * F * H_full * W * SCALE * SCALE
    value = out_pix_per_step * args.steps / (ms_max * 1e-3) / 1e6

    # ---------------- end-to-end leg: host buffers through the C ABI ----------------
    # (a) one synchronous call per step: the pipeline fills and drains inside every step
    for _ in range(2):
        eng.infer_host(host_in, host_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.infer_host(host_in, host_out)       # synchronous: H2D + kernels + D2H inside
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_sync_value = out_pix_per_step * args.steps / float(t.item()) / 1e6
    sync_out = np.array(host_out) if rank == 0 else None       # kept for the parity check (the streaming leg reuses host_out)
    # (b) the streaming form (a video pipeline): the same steps enqueued back to back, outputs alternating
    # between two pinned buffers, one wait at the end - every step's H2D and D2H is still inside the region
    host_out2 = pinned_empty(host_out.shape)
    outs = (host_out, host_out2)
    for i in range(2):
        eng.infer_host_async(host_in, outs[i & 1])
    eng.host_sync()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        eng.infer_host_async(host_in, outs[i & 1])
    eng.host_sync()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = out_pix_per_step * args.steps / float(t.item()) / 1e6
    if rank == 0 and not np.array_equal(host_out2, host_out):
        raise SystemExit("streaming host path: the two output buffers differ")
    sanity = int(host_out[0, :8, :8].sum())    # device->host read of the step's result

    # ---------------- parity of what was timed (outside every timed region) ----------------
    # One whole frame of each timed leg - device-resident, blocking host call, streaming host call - against
    # the C oracle (the checker; oracle/mulut_oracle.c) on the same input frame; all frames of the legs must
    # also agree with each other byte for byte.  A mismatch fails the run.
    parity = None
    if rank == 0:
        from oracle import c_oracle as CO
        dev_out = d_out.cpu().numpy()
        if strip:      # the whole frame through the oracle; this rank's strip is rows [row_b, row_e) of it
            ref0 = CO.sr_u8(frame0_whole, luts, STAGES, MODES, SCALE, INTERVAL, 0)[:, row_b * SCALE:row_e * SCALE]
            k0, k1 = (row_b - load_b) * SCALE, (row_e - load_b) * SCALE
        else:
            ref0 = CO.sr_u8(np.asarray(host_in[:1]), luts, STAGES, MODES, SCALE, INTERVAL, 0)
            k0, k1 = 0, HL * SCALE
        legs = {"device": dev_out, "host_sync": sync_out, "host_async": host_out, "host_async (second buffer)": host_out2}
        for name, arr in legs.items():
            if not np.array_equal(arr[0, k0:k1], ref0[0]):
                raise SystemExit("PARITY FAILURE: leg '{}' differs from the C oracle in {} bytes".format(
                    name, int((arr[0, k0:k1] != ref0[0]).sum())))
            if not np.array_equal(arr, dev_out):
                raise SystemExit("PARITY FAILURE: leg '{}' differs from the device leg".format(name))
        parity = {"result": "bit-exact", "checked": "frame 0 of the device, host_sync and host_async legs vs the C oracle "
                  "({} bytes each); all {} frames of the three legs identical".format(ref0[0].size, F)}
        del dev_out, sync_out

    # ---------------- (c) the copy ceiling (after the parity check: it leaves stale bytes in the host buffers):
    # the SAME copies on the same lanes (per frame one cudaMemcpyAsync in, one out) with no kernels between
    # them - what this machine's host link allows the streaming path at this N
    for i in range(2):
        eng.host_copy_probe_async(host_in, outs[i & 1])
    eng.host_sync()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        eng.host_copy_probe_async(host_in, outs[i & 1])
    eng.host_sync()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    copy_ceiling = out_pix_per_step * args.steps / float(t.item()) / 1e6

    # ---------------- the same step on natural-like frames (reported beside the headline) ----------------
    natural = None
    if args.data == "uniform" and world == 1:
        DATA = "natural"
        d_nat = torch.from_numpy(make_frames(F, seed=77)).cuda()
        DATA = "uniform"
        for _ in range(3):
            eng.infer_device(d_nat, d_out)
        n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        n0.record()
        for _ in range(10):
            eng.infer_device(d_nat, d_out)
        n1.record()
        torch.cuda.synchronize()
        natural = {"value": F * HL * W * SCALE * SCALE * 10 / (n0.elapsed_time(n1) * 1e-3) / 1e6, "unit": "Mpix/s",
                   "steps": 10, "frames": "mirror-tiled crops of the reference's golden Set5 results, device-resident"}
        del d_nat

    # ---------------- cfg4: the LUT finetune step (every rank takes part: data-parallel, one all-reduce per step) ----------------
    finetune = None
    if not args.no_finetune:
        del d_in, d_out
        torch.cuda.empty_cache()
        from tools.finetune_bench import finetune_block
        try:
            finetune, ft_state = finetune_block(rank, world, local, dist if world > 1 else None, steps=args.finetune_steps,
                                                warmup=10, reference_fn=reference_module_step,
                                                clock_sampler=lambda: ClockSampler(local))
            ft_state.close()
            del ft_state                       # the captured graphs hold NCCL work: dropped before the group goes away
            # the same step on natural-like (smooth) patches: the realistic case for the LUT-gradient scatter
            smooth, ft_state = finetune_block(rank, world, local, dist if world > 1 else None, steps=args.finetune_steps,
                                              warmup=10, smooth=True, phases_wanted=False)
            ft_state.close()
            del ft_state
            finetune["smooth_patches"] = smooth
        except Exception as e:
            finetune = {"error": repr(e)[:400]}
        torch.cuda.synchronize()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel ----------------
    # Per kernel kind: vertex gathers per input sample, LUT bytes per gather (algorithmic), algorithmic HBM
    # bytes per sample, and which measured gather primitive is its ceiling.
    r2, M = SCALE * SCALE, len(MODES)
    KINFO = {
        "smem_stage":    (60, 1, 1 + 2 * M, "lds_u8", "K1h/K1a: vertex gathers from the shared-memory LUT (K1h fetches the 60 algorithmic 1-byte vertices as 48 two-byte pairs)"),
        "fused_stage":   (60, 1, 2, "lds_u8", "K1i: K1h (a-paired 16-bit table in shared memory: the 60 algorithmic 1-byte vertices as 48 two-byte gathers) with the mode combine and the stage epilogue fused; partial tiles are exchanged through an L2-resident ring"),
        "generic_stage": (60, 1, 2, "ldg_u32", "K0: vertex gathers through L1/L2"),
        "last_binned":   (60, r2, 1 + r2, "lds_u32", "K1f: 4-byte vertex-row gathers from shared-memory LUT slabs"),
        "last_tiled":    (60, r2, 1 + r2, "quad_cell256_3rows" if SCALE == 4 else "quad_cell64",
                          "K1e: three 64-byte row-blocks of a 256-byte cell per interpolation (L2)" if SCALE == 4 else
                          "K1c: 64-byte cell fetches from L1/L2 (one per 5 vertices)"),
        "generic_last":  (60, r2, 1 + r2, "ldg_u32", "K0: vertex gathers through L1/L2"),
        "combine":       (0, 0, 2 * M + 1, None, "K1b: streaming"),
        "bin_hist":      (0, 0, 2, None, "K1f preparation: streaming"),
        "bin_orphans":   (0, 0, 0, None, "K1f orphan list"),
    }
    samples_per_step = F * HL * W * C
    hbm_peak, peak_src = measured_peaks()
    dom = max(prof.items(), key=lambda kv: kv[1][0]) if prof else (None, (0.0, 0))
    dom_name, (dom_ms, dom_n) = dom
    n_gather, b_gather, hbm_b, peak_variant, kdesc = KINFO.get(dom_name, (0, 0, HBM_B, None, ""))
    per_launch_s = dom_ms * 1e-3 / max(dom_n, 1)
    step_s = ms * 1e-3 / args.steps
    gp = gather_peak(local)

    def peak_gathers(variant):
        g = gp.get(variant, {}).get("gathers_per_s")
        if not g:
            return None
        return g * 5 if variant.startswith("quad_cell") else g   # one cell fetch serves the 5 vertices of an interpolation

    # measured DRAM bytes per input sample and kernel (one ncu --set full capture per kernel, profiles/traffic.json)
    traffic_tab = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tp):
        try:
            traffic_tab = {k: v for k, v in json.load(open(tp)).items() if isinstance(v, dict)}
        except Exception:
            traffic_tab = {}

    # PRIMARY roofline: the path is bound by ON-CHIP gathers (300 B of LUT rows per input sample against 5 B of
    # HBM), so the dominant kernel's algorithmic vertex-row bytes per second are set against the measured rate
    # of the same gather primitive (ONE fixed peak per primitive: best launch shape of mulut_gather_bench).
    ach_gathers = samples_per_step * n_gather / per_launch_s if per_launch_s > 0 else 0.0
    pk = peak_gathers(peak_variant) if peak_variant else None
    per_kernel, t_peak_sum = {}, 0.0
    for kname, (kms, kn) in prof.items():
        ng, bg, kh, pv, _ = KINFO.get(kname, (0, 0, 0, None, ""))
        if not kn:
            continue
        k_s = kms * 1e-3 / kn
        if ng:
            g = samples_per_step * ng / k_s
            kp = peak_gathers(pv)
            per_kernel[kname] = {"Ggathers_per_s": g / 1e9, "GBps": g * bg / 1e9, "peak_primitive": pv,
                                 "frac": g / kp if kp else None, "launch_ms": k_s * 1e3}
            if kp:
                t_peak_sum += samples_per_step * ng / kp
        elif kh:                               # streaming kernels: their floor is the HBM time of their algorithmic bytes
            t_peak_sum += samples_per_step * kh / (hbm_peak * 1e9)
    dom_traffic = traffic_tab.get(dom_name, {}).get("dram_bytes_per_sample")
    roofline = {
        "bound": "on-chip gather (shared memory / L1-L2): not hbm, not tensor - see roofline_hbm for the HBM side",
        "kernel": dom_name, "kernel_desc": kdesc,
        "achieved": ach_gathers * b_gather / 1e9, "peak": pk * b_gather / 1e9 if pk else None, "unit": "GB/s",
        "frac": ach_gathers / pk if pk else None,
        "traffic": dom_traffic * samples_per_step if dom_traffic is not None else None,
        "traffic_def": "DRAM bytes of this kernel per launch (ncu dram__bytes_read+write per sample from profiles/traffic.json "
                       "x samples per launch); its algorithmic HBM bytes are {} B/sample".format(hbm_b),
        "achieved_Ggathers_per_s": ach_gathers / 1e9, "peak_Ggathers_per_s": pk / 1e9 if pk else None,
        "alg_gathers_per_sample": n_gather, "alg_bytes_per_gather": b_gather,
        "peak_def": "mulut_gather_bench '{}' measured live: independent random gathers of this shape from the same memory "
                    "level on all SMs, best launch shape ({})".format(peak_variant, gp.get(peak_variant, {})),
        "launch_ms": per_launch_s * 1e3, "share_of_step": dom_ms / ms if ms > 0 else None,
        "whole_step": {"frac": t_peak_sum / step_s if step_s > 0 else None,
                       "def": "sum over the step's kernels of (algorithmic gathers / primitive peak, or algorithmic HBM bytes / "
                              "HBM peak for the streaming kernels) divided by the measured step time",
                       "Ggathers_per_s": samples_per_step * 60 * STAGES / step_s / 1e9},
        "per_kernel": per_kernel,
        # the north star's own yardstick: the L2 gather roofline.  The shared-memory kernels are not bound by it - the
        # whole step delivers more vertex rows per second than L2 serves with its best fetch shape
        "whole_step_vs_l2_gather_roofline": {
            "l2_cell64_x5_Ggathers_per_s": (peak_gathers("quad_cell64") or 0) / 1e9,
            "l2_u32_Ggathers_per_s": (peak_gathers("ldg_u32") or 0) / 1e9,
            "frac_of_l2_cell64": (samples_per_step * 60 * STAGES / step_s) / peak_gathers("quad_cell64")
            if peak_gathers("quad_cell64") else None,
            "frac_of_l2_u32": (samples_per_step * 60 * STAGES / step_s) / peak_gathers("ldg_u32")
            if peak_gathers("ldg_u32") else None,
        },
        "microbench": gp,
    }
    # HBM side, WHOLE step: algorithmic = 1 B read + r^2 B written per input sample
    step_traffic = None
    if traffic_tab and prof and all(k in traffic_tab for k in prof):
        step_traffic = sum(traffic_tab[k]["dram_bytes_per_sample"] for k in prof) * samples_per_step
    ach_hbm = samples_per_step * HBM_B / step_s / 1e9 if step_s > 0 else 0.0
    roofline_hbm = {"bound": "hbm", "scope": "whole step", "achieved": ach_hbm, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach_hbm / hbm_peak, "alg_bytes_per_sample": HBM_B,
                    "traffic": step_traffic,
                    "traffic_bytes_per_sample": step_traffic / samples_per_step if step_traffic else None,
                    "traffic_GBps": step_traffic / step_s / 1e9 if step_traffic else None,
                    "peak_source": peak_src + " (burst copy figure)",
                    "note": "structurally small: 5 B of HBM per 300 B gathered on chip"}
    kernels = {k: {"ms_total": v[0], "launches": v[1], "ms_per_launch": v[0] / v[1]} for k, v in prof.items()}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        v, cores, n, dt_cpu = cpu_port_throughput(args.cpu_seconds, 64)
        cpu = {"value": v, "unit": "Mpix/s", "cores": cores, "kind": "port",
               "sample": "{} frame(s) {}x{} in {:.1f} s, C port of the reference algorithm (oracle/mulut_oracle.c), "
                         "pthreads over rows".format(n, W, H, dt_cpu)}
        if not args.no_numpy_ref:
            try:
                cpu["numpy_ref"] = numpy_reference_throughput(args.config)
            except Exception as e:
                cpu["numpy_ref"] = {"error": repr(e)[:300]}

    line = {
        "metric": "output Mpix/s", "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "strong" if strip else "weak",
        "vs_baseline": None, "dtype": "u8",
        "data": "synthetic" if DATA == "uniform" else "synthetic (mirror-tiled natural crops)",
        "config": {"workload": WORKLOAD, "frames_per_gpu_per_step": F, "kernel": args.kernel,
                   "l2": "per-step working set per GPU = {:.0f} MB in + {:.0f} MB out (> 126 MB L2), no flush".format(
                       F * HL * W * C / 1e6, F * HL * W * C * SCALE * SCALE / 1e6),
                   "parallelism": ("every frame strip-sharded over {} GPUs ({} + {} halo rows on rank 0), no collective".format(
                       world, row_e - row_b, HL - (row_e - row_b)) if strip else
                       "frames sharded over {} GPU(s), no collective".format(world))},
        "e2e": {"value": e2e_value, "unit": "Mpix/s", "h2d_bytes_per_step": F * HL * W * C,
                "d2h_bytes_per_step": F * HL * W * C * SCALE * SCALE, "api": "LutEngine.infer_host_async x steps + host_sync -> mulut_sr_infer_u8_host_async / mulut_sr_host_sync",
                "sync_per_step": {"value": e2e_sync_value, "unit": "Mpix/s", "api": "LutEngine.infer_host -> mulut_sr_infer_u8_host, one blocking call per step"},
                "copy_ceiling": {"value": copy_ceiling, "unit": "Mpix/s", "frac": e2e_value / copy_ceiling if copy_ceiling > 0 else None,
                                 "def": "the same per-frame H2D + D2H copies on the same lanes with no kernels (mulut_host_copy_probe_async), "
                                        "all {} rank(s) at once: the host-link ceiling of this machine at this N".format(world)},
                "host_memory": "pinned", "check": sanity},
        "gpu_launches": int(launches) * world,
        "parity": parity,
        "clocks": clocks,
        "roofline": roofline,
        "roofline_hbm": roofline_hbm,
        "kernels": kernels,
        "natural_frames": natural,
        "cpu_baseline": cpu,
        "finetune": finetune,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
