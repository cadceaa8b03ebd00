"""Host-side mirror of the reference's inference interface (sr/4_test_lut.py).

* :func:`load_luts`            the LUT-loading block, 4_test_lut.py:323-333
* :class:`LutEngine`           owns the device LUT replica (C ABI handle) and runs
                               the fused stages x modes x rotations path that
                               replaces eltr._worker's loop, 4_test_lut.py:279-306
* :func:`FourSimplexInterpFaster`  call-compatible single pass, 4_test_lut.py:14-237
* :class:`eltr`                dataset runner with the reference's class name, file
                               layout and printed summary, 4_test_lut.py:240-316

All compute goes through libmulut_b200.so; nothing here falls back to numpy.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Iterable, Optional

import numpy as np

from . import _lib

VALID_MODES = ("s", "d", "y")
MODE_PAD = {"s": 1, "d": 2, "y": 2}      # 4_test_lut.py:289-292


def _check_modes(modes: Iterable[str]) -> str:
    modes = "".join(modes)
    for m in modes:
        if m not in VALID_MODES:
            # same message as 4_test_lut.py:52-54
            raise ValueError("Mode {} not implemented.".format(m))
    return modes


def lut_path(exp_dir: str, lut_name: str, scale: int, interval: int, stage: int, mode: str) -> str:
    """File-name rule of the TEST path (4_test_lut.py:331-332): note `8 - interval`."""
    return os.path.join(exp_dir, "{}_x{}_{}bit_int8_s{}_{}.npy".format(lut_name, scale, 8 - interval, stage, mode))


def load_luts(exp_dir: str, stages: int = 2, modes="sdy", scale: int = 4, interval: int = 4,
              lut_name: str = "LUT_ft") -> Dict[str, np.ndarray]:
    """lutDict of 4_test_lut.py:323-333, kept as int8 (the reference casts to
    float32; values are identical)."""
    luts = {}
    for s in range(stages):
        v_num = scale * scale if (s + 1) == stages else 1
        for mode in modes:
            key = "s{}_{}".format(s + 1, mode)
            arr = np.load(lut_path(exp_dir, lut_name, scale, interval, s + 1, mode))
            luts[key] = np.ascontiguousarray(arr.reshape(-1, v_num).astype(np.int8))
    return luts


class LutEngine:
    """Device-resident MuLUT inference engine (one per GPU).

    luts: dict "s{stage}_{mode}" -> integer array (L^4, 1 | scale^2) in the
    reference's on-disk layout.
    """

    def __init__(self, luts: Dict[str, np.ndarray], stages: int = 2, modes="sdy", scale: int = 4,
                 interval: int = 4, device: int = 0, kernel: int = _lib.KERNEL_AUTO):
        self.modes = _check_modes(modes)
        self.stages, self.scale, self.interval, self.device = int(stages), int(scale), int(interval), int(device)
        self._h = ctypes.c_void_p()
        self._pending = []
        tabs = []
        rows = None
        for s in range(self.stages):
            cols = self.scale * self.scale if (s + 1) == self.stages else 1
            for m in self.modes:
                key = "s{}_{}".format(s + 1, m)
                if key not in luts:
                    raise KeyError(key)
                t = np.asarray(luts[key])
                t = np.ascontiguousarray(t.reshape(-1, cols).astype(np.int8))
                rows = t.shape[0] if rows is None else rows
                if t.shape[0] != rows:
                    raise ValueError("all LUTs must have the same number of rows")
                tabs.append(t)
        self._tabs = tabs
        ptrs = (ctypes.c_void_p * len(tabs))(*[t.ctypes.data for t in tabs])
        L = _lib.lib()
        _lib.check(L.mulut_create(ctypes.byref(self._h), self.device, self.stages, self.modes.encode(),
                                  self.scale, self.interval, ptrs, int(rows)))
        if kernel != _lib.KERNEL_AUTO:
            self.set_kernel(kernel)

    # -- lifetime -----------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().mulut_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- configuration --------------------------------------------------------
    def set_kernel(self, kernel: int) -> None:
        _lib.check(_lib.lib().mulut_set_kernel(self._h, int(kernel)))

    def reserve(self, N: int, H: int, W: int, C: int) -> None:
        _lib.check(_lib.lib().mulut_reserve(self._h, N, H, W, C))

    @property
    def launch_count(self) -> int:
        return int(_lib.lib().mulut_launch_count(self._h))

    def profile(self, on: bool) -> None:
        """Bracket every kernel launch with CUDA events (bench.py's roofline leg)."""
        _lib.check(_lib.lib().mulut_profile_enable(self._h, 1 if on else 0))

    def profile_read(self) -> dict:
        """{kernel kind: (total_ms, launches)} since profile(True); synchronises."""
        out = {}
        for name, kind in _lib.PROF_KINDS.items():
            ms, n = ctypes.c_double(), ctypes.c_longlong()
            _lib.check(_lib.lib().mulut_profile_read(self._h, kind, ctypes.byref(ms), ctypes.byref(n)))
            if n.value:
                out[name] = (ms.value, n.value)
        return out

    # -- hot path -------------------------------------------------------------
    def infer_device(self, frames, out=None):
        """frames: torch.uint8 CUDA tensor (N,H,W,C) or (H,W,C), contiguous.
        Returns a torch.uint8 CUDA tensor (N,H*r,W*r,C) (or 3-D).  Asynchronous on
        torch's current stream."""
        import torch
        if not (isinstance(frames, torch.Tensor) and frames.is_cuda and frames.dtype == torch.uint8):
            raise TypeError("infer_device expects a CUDA uint8 tensor")
        squeeze = frames.dim() == 3
        x = frames[None] if squeeze else frames
        if x.dim() != 4:
            raise ValueError("expected (N,H,W,C) or (H,W,C)")
        x = x.contiguous()
        N, H, W, C = x.shape
        r = self.scale
        if out is None:
            out = torch.empty((N, H * r, W * r, C), dtype=torch.uint8, device=x.device)
        elif tuple(out.shape) != (N, H * r, W * r, C) or not out.is_contiguous() or out.dtype != torch.uint8:
            raise ValueError("bad out tensor")
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(_lib.lib().mulut_sr_infer_u8(self._h, x.data_ptr(), out.data_ptr(), N, H, W, C,
                                                ctypes.c_void_p(stream)))
        return out[0] if squeeze else out

    def infer_host(self, frames: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        """frames: numpy uint8 (N,H,W,C) or (H,W,C) in host memory (pinned is
        faster: :func:`pinned_empty`).  H2D, kernels and D2H are pipelined per
        frame inside the library.  Synchronous."""
        a = np.ascontiguousarray(frames, dtype=np.uint8)
        squeeze = a.ndim == 3
        if squeeze:
            a = a[None]
        if a.ndim != 4:
            raise ValueError("expected (N,H,W,C) or (H,W,C)")
        N, H, W, C = a.shape
        r = self.scale
        if out is None:
            out = np.empty((N, H * r, W * r, C), dtype=np.uint8)
        o = out[None] if (squeeze and out.ndim == 3) else out
        if o.shape != (N, H * r, W * r, C) or o.dtype != np.uint8 or not o.flags.c_contiguous:
            raise ValueError("bad out array")
        _lib.check(_lib.lib().mulut_sr_infer_u8_host(self._h, a.ctypes.data, o.ctypes.data, N, H, W, C))
        return o[0] if squeeze else o

    def infer_host_async(self, frames: np.ndarray, out: np.ndarray) -> None:
        """Streaming form of :meth:`infer_host`: enqueue (N,H,W,C) uint8 `frames` -> `out` and return.
        Consecutive calls overlap (the copy-in and kernels of a call run under the copy-out of the one
        before).  Both arrays must be C-contiguous uint8, should be pinned (:func:`pinned_empty`) and must
        not be touched until :meth:`host_sync` returns."""
        if not (isinstance(frames, np.ndarray) and isinstance(out, np.ndarray) and frames.dtype == np.uint8 and
                out.dtype == np.uint8 and frames.flags.c_contiguous and out.flags.c_contiguous and frames.ndim == 4):
            raise ValueError("infer_host_async expects C-contiguous uint8 (N,H,W,C) arrays")
        N, H, W, C = frames.shape
        if out.shape != (N, H * self.scale, W * self.scale, C):
            raise ValueError("bad out array")
        self._pending.append((frames, out))                      # keep the buffers alive until host_sync
        _lib.check(_lib.lib().mulut_sr_infer_u8_host_async(self._h, frames.ctypes.data, out.ctypes.data, N, H, W, C))

    def host_copy_probe_async(self, frames: np.ndarray, out: np.ndarray) -> None:
        """The H2D / D2H copies of :meth:`infer_host_async` with no kernels between them (bench.py's
        e2e.copy_ceiling: what the host link of this machine allows the streaming path)."""
        N, H, W, C = frames.shape
        self._pending.append((frames, out))
        _lib.check(_lib.lib().mulut_host_copy_probe_async(self._h, frames.ctypes.data, out.ctypes.data, N, H, W, C))

    def infer_strip(self, frame, rank: int, world: int, out=None):
        """Strip-sharded inference of ONE frame ("for single huge frames: horizontal strips with
        a 4-row input halo"): computes rows [b*r, e*r) of the output, (b, e) = the `rank`-th of `world`
        contiguous row ranges of the input, from input rows [b - halo, e + halo) with halo = 2 rows per
        stage (each stage reads a 5x5 window).  No collective: concatenating the strips of all ranks along
        the row axis gives the whole-frame bytes exactly - at the frame's own top/bottom edge the strip ends
        at the edge, so the kernels' coordinate clamp supplies the reference's edge replication there, and
        an interior strip's halo rows keep every kept output row out of reach of the strip's artificial edge.
        frame: CUDA uint8 (H,W,C) tensor or host numpy array of that shape (only the strip's rows are read).
        Returns (row_begin, row_end, strip) with strip of shape ((e-b)*r, W*r, C), same kind as `frame`."""
        from .dist import shard_rows_with_halo
        if frame.ndim != 3:
            raise ValueError("infer_strip expects one (H,W,C) frame")
        H = int(frame.shape[0])
        r = self.scale
        b, e, lb, le = shard_rows_with_halo(H, rank, world, 2 * self.stages)
        if e <= b:
            empty = frame[0:0]
            return b, e, (empty.new_empty((0, frame.shape[1] * r, frame.shape[2])) if not isinstance(frame, np.ndarray)
                          else np.empty((0, frame.shape[1] * r, frame.shape[2]), np.uint8))
        piece = frame[lb:le]
        if isinstance(frame, np.ndarray):
            full = self.infer_host(np.ascontiguousarray(piece))
        else:
            full = self.infer_device(piece.contiguous())
        strip = full[(b - lb) * r:(e - lb) * r]
        if out is not None:
            out[...] = strip
            strip = out
        return b, e, strip

    def host_sync(self) -> None:
        """Wait for every :meth:`infer_host_async` call issued so far."""
        try:
            _lib.check(_lib.lib().mulut_sr_host_sync(self._h))
        finally:
            self._pending.clear()

    def __call__(self, frames, out=None):
        """numpy in -> numpy out (host path); CUDA tensor in -> CUDA tensor out."""
        if isinstance(frames, np.ndarray):
            if frames.ndim == 2:                       # grey -> 3 channels, 4_test_lut.py:268-270
                frames = np.stack([frames] * 3, axis=2)
            return self.infer_host(frames, out)
        return self.infer_device(frames, out)


class _PinnedBuf:
    """Owner of a cudaMallocHost block; numpy views keep it alive through `.base`."""

    def __init__(self, nbytes: int):
        self.ptr = _lib.lib().mulut_host_alloc(nbytes)
        if not self.ptr:
            raise MemoryError(_lib.last_error())
        self.__array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 3}

    def __del__(self):
        try:
            if self.ptr:
                _lib.lib().mulut_host_free(ctypes.c_void_p(self.ptr))
                self.ptr = None
        except Exception:
            pass


def pinned_empty(shape, dtype=np.uint8) -> np.ndarray:
    """numpy array backed by page-locked host memory (for LutEngine.infer_host)."""
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    raw = np.asarray(_PinnedBuf(max(count * dtype.itemsize, 1)))
    return raw[:count * dtype.itemsize].view(dtype).reshape(shape)


def FourSimplexInterpFaster(weight, img_in, h, w, interval, rot, upscale=4, mode="s", device: int = 0):
    """Call-compatible replacement of the reference function of the same name
    (sr/4_test_lut.py:14): numpy in, numpy float64 out, computed on the GPU."""
    import torch
    _check_modes(mode)
    if mode not in VALID_MODES:
        raise ValueError("Mode {} not implemented.".format(mode))
    dev = torch.device("cuda", device)
    wt = torch.as_tensor(np.ascontiguousarray(np.asarray(weight, dtype=np.float32).reshape(-1, upscale * upscale)),
                         device=dev)
    x = torch.as_tensor(np.ascontiguousarray(np.asarray(img_in, dtype=np.float32)), device=dev)
    C = x.shape[0]
    p = MODE_PAD[mode]
    if tuple(x.shape) != (C, h + p, w + p):
        raise ValueError("img_in must be (C, h+{p}, w+{p}) for mode {m}".format(p=p, m=mode))
    k = rot % 4
    oshape = (C, h * upscale, w * upscale) if k % 2 == 0 else (C, w * upscale, h * upscale)
    out = torch.empty(oshape, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(_lib.lib().mulut_interp_pass_f64(wt.data_ptr(), wt.shape[0], x.data_ptr(), C, h, w, interval, rot,
                                                upscale, mode.encode(), out.data_ptr(), ctypes.c_void_p(stream)))
    return out.cpu().numpy()


class eltr:
    """Dataset runner mirroring the reference class (4_test_lut.py:240-316):
    same directory layout (`testDir/<dataset>/{HR,LR_bicubic/X<scale>}`), same
    output names and the same printed summary.  Images go through one
    LutEngine instead of a multiprocessing.Pool of numpy workers."""

    def __init__(self, dataset, opt, lutDict, device: int = 0):
        folder = os.path.join(opt.testDir, dataset, "HR")
        files = sorted(os.listdir(folder))
        exp_name = opt.expDir.rstrip("/").split("/")[-1]
        result_path = os.path.join(opt.resultRoot, exp_name, dataset, "X{}".format(opt.scale))
        os.makedirs(result_path, exist_ok=True)
        self.result_path, self.dataset, self.files, self.opt = result_path, dataset, files, opt
        self.engine = LutEngine(lutDict, opt.stages, opt.modes, opt.scale, opt.interval, device=device)

    def run(self, num_worker=None):
        """PSNR / SSIM are computed on the GPU (mulut_eval_psnr_ssim_y_u8) from the device-resident
        SR frame and the uploaded HR frame; the SR frame comes back once, for the PNG."""
        import torch
        from .metrics import modcrop, psnr_ssim_device
        from PIL import Image
        dev = torch.device("cuda", self.engine.device)
        res = []
        for name in self.files:
            lr = np.array(Image.open(os.path.join(self.opt.testDir, self.dataset,
                                                  "LR_bicubic/X{}".format(self.opt.scale), name)))
            if lr.ndim == 2:
                lr = np.stack([lr] * 3, axis=2)
            gt = modcrop(np.array(Image.open(os.path.join(self.opt.testDir, self.dataset, "HR", name))),
                         self.opt.scale)
            if gt.ndim == 2:
                gt = np.stack([gt] * 3, axis=2)
            d_out = self.engine(torch.from_numpy(np.ascontiguousarray(lr[:, :, :3])).to(dev))
            d_gt = torch.from_numpy(np.ascontiguousarray(gt[:, :, :3])).to(dev)
            res.append(list(psnr_ssim_device(d_gt, d_out, self.opt.scale)))
            Image.fromarray(d_out.cpu().numpy()).save(os.path.join(
                self.result_path, "{}_{}_{}bit.png".format(name[:-4], self.opt.lutName, 8 - self.opt.interval)))
        res = np.asarray(res)
        print("Dataset {} | AVG LUT PSNR: {:.2f} SSIM: {:.4f}".format(self.dataset, res[:, 0].mean(), res[:, 1].mean()))
        return res
