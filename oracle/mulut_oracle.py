"""CPU oracle for the MuLUT LUT-retrieval hot path (TEST INFRASTRUCTURE ONLY).

This file is the checker, not the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product path (``mulut_b200``) never does.

It restates, in all-integer numpy, the algorithm of the reference's numpy
inference path:

* ``FourSimplexInterpFaster``      /root/reference/sr/4_test_lut.py:14-237
* the stage/mode/rotation loop     /root/reference/sr/4_test_lut.py:279-306
* LUT file naming / loading        /root/reference/sr/4_test_lut.py:323-333

The reference enumerates the 24 orderings of the four LSB fractions with
boolean masks; this restatement uses the equivalent closed form "sort the four
fractions descending and walk the simplex" (SURVEY.md §8-SPEC), and folds the
rot90 -> pad(edge) -> interpolate -> rot90-back sandwich into rotated tap
offsets with clamped coordinates, so no image is ever rotated.

Parity is PINNED: ``tests/test_oracle_pinned.py`` checks this file against the
reference's five golden Set5 PNGs (results/sr_x2sdy/Set5/X4/*.png, committed as
tests/golden/set5_x4.npz) and against fixtures produced by importing the
reference itself (oracle/make_golden.py).
"""
from __future__ import annotations

import numpy as np

# Tap offsets (dy, dx) of taps a, b, c, d in the un-rotated frame.
# /root/reference/sr/4_test_lut.py:18-51 (slices of img_in per mode).
MODE_TAPS = {
    "s": ((0, 0), (0, 1), (1, 0), (1, 1)),
    "d": ((0, 0), (0, 2), (2, 0), (2, 2)),
    "y": ((0, 0), (1, 1), (1, 2), (2, 1)),
}
# bottom/right edge padding per mode, /root/reference/sr/4_test_lut.py:289-292
MODE_PAD = {"s": 1, "d": 2, "y": 2}


def check_mode(mode: str) -> None:
    if mode not in MODE_TAPS:
        # same message as /root/reference/sr/4_test_lut.py:52-54
        raise ValueError("Mode {} not implemented.".format(mode))


def rotated_taps(mode: str, r: int):
    """Tap offsets seen from the un-rotated frame when the reference rotates
    the image by ``np.rot90(img, r)`` before sampling (4_test_lut.py:294)."""
    check_mode(mode)
    taps = []
    for dy, dx in MODE_TAPS[mode]:
        for _ in range(r % 4):
            dy, dx = dx, -dy
        taps.append((dy, dx))
    return tuple(taps)


def rotated_subpixel(u: int, v: int, up: int, r: int):
    """Where LUT column (u, v) of the up x up block lands after the reference
    rotates the interpolated image back by ``np.rot90(out, 4 - r)``
    (4_test_lut.py:235,297)."""
    for _ in range(r % 4):
        u, v = v, up - 1 - u
    return u, v


def lut_geometry(interval: int):
    q = 1 << interval                      # 4_test_lut.py:15
    L = (1 << (8 - interval)) + 1          # 4_test_lut.py:16
    return q, L


def simplex_vertices(t, interval: int):
    """Vectorised core: for tap values ``t`` (4, ...) ints in [0,255] return
    (vertex indices (5, ...), integer weights (5, ...), order (4, ...)).

    Sort descending by LSB fraction, ties -> higher tap index first (this tie
    rule only matters for d/d-input in the torch twin; the forward value is
    tie-independent because tied vertices get weight 0)."""
    q, L = lut_geometry(interval)
    t = np.asarray(t, dtype=np.int64)
    m = t >> interval
    f = t & (q - 1)
    stride = np.array([L * L * L, L * L, L, 1], dtype=np.int64)
    stride = stride.reshape((4,) + (1,) * (t.ndim - 1))
    v0 = (m * stride).sum(axis=0)
    tap = np.arange(4, dtype=np.int64).reshape((4,) + (1,) * (t.ndim - 1))
    key = f * 4 + tap                      # distinct keys; larger = earlier
    order = np.argsort(-key, axis=0, kind="stable")
    fs = np.take_along_axis(f, order, axis=0)
    ss = np.take_along_axis(np.broadcast_to(stride, t.shape), order, axis=0)
    verts = np.empty((5,) + t.shape[1:], dtype=np.int64)
    verts[0] = v0
    for j in range(4):
        verts[j + 1] = verts[j] + ss[j]
    w = np.empty((5,) + t.shape[1:], dtype=np.int64)
    w[0] = q - fs[0]
    w[1] = fs[0] - fs[1]
    w[2] = fs[1] - fs[2]
    w[3] = fs[2] - fs[3]
    w[4] = fs[3]
    return verts, w, order


def interp_unrotated(lut, img, mode: str, r: int, up: int, interval: int):
    """One (mode, rotation) interpolation pass in the un-rotated frame.

    lut : int array (n_rows, up*up) (int8 values)
    img : int array (H, W, C), values 0..255
    returns int64 (H*up, W*up, C) = q * FourSimplexInterpFaster(...) of the
    reference's rotate/pad/interp/rotate-back sandwich for rotation ``r``.
    """
    lut = np.asarray(lut).reshape(-1, up * up).astype(np.int64)
    img = np.asarray(img).astype(np.int64)
    H, W, C = img.shape
    ys = np.arange(H)
    xs = np.arange(W)
    taps = []
    for dy, dx in rotated_taps(mode, r):
        yy = np.clip(ys + dy, 0, H - 1)
        xx = np.clip(xs + dx, 0, W - 1)
        taps.append(img[yy][:, xx])
    t = np.stack(taps, axis=0)             # (4, H, W, C)
    verts, w, _ = simplex_vertices(t, interval)
    _, L = lut_geometry(interval)
    if verts.max(initial=0) >= lut.shape[0]:
        raise IndexError("LUT too small: need {} rows, have {}".format(L ** 4, lut.shape[0]))
    o = np.zeros((H, W, C, up * up), dtype=np.int64)
    for k in range(5):
        o += w[k][..., None] * lut[verts[k]]
    out = np.zeros((H * up, W * up, C), dtype=np.int64)
    for u in range(up):
        for v in range(up):
            uu, vv = rotated_subpixel(u, v, up, r)
            out[uu::up, vv::up, :] = o[..., u * up + v]
    return out


def round_half_even_div(num, den: int):
    """Integer round-half-to-even of num/den (den > 0); np.round semantics of
    4_test_lut.py:302 without floating point."""
    num = np.asarray(num, dtype=np.int64)
    qd = np.floor_divide(num, den)
    rm = num - qd * den
    up_ = (2 * rm > den) | ((2 * rm == den) & ((qd & 1) == 1))
    return qd + up_.astype(np.int64)


def stage_epilogue(S, n_modes: int, interval: int, last: bool):
    """4_test_lut.py:281-286,300-306 in integers.  S = sum over modes and
    rotations of q * interp."""
    q, _ = lut_geometry(interval)
    if last:
        x = round_half_even_div(S, q * n_modes)
    else:
        D = q * n_modes * 4
        x = round_half_even_div(S + 127 * D, D)
    return np.clip(x, 0, 255)


def sr_pipeline(img_u8, luts: dict, stages: int, modes, scale: int, interval: int = 4):
    """Whole inference path for one HWC uint8 image (4_test_lut.py:279-306).

    luts: dict "s{stage}_{mode}" -> int8 array (L^4, 1) for non-last stages,
    (L^4, scale^2) for the last stage."""
    img = np.asarray(img_u8)
    if img.ndim == 2:                      # 4_test_lut.py:268-270 (grey -> 3ch)
        img = np.stack([img, img, img], axis=2)
    x = img.astype(np.int64)
    modes = list(modes)
    for m in modes:
        check_mode(m)
    for s in range(stages):
        last = (s + 1) == stages
        up = scale if last else 1
        S = 0
        for mode in modes:
            lut = luts["s{}_{}".format(s + 1, mode)]
            for r in range(4):
                S = S + interp_unrotated(lut, x, mode, r, up, interval)
        x = stage_epilogue(S, len(modes), interval, last)
    return x.astype(np.uint8)


def sr_pipeline_batch(frames_u8, luts, stages, modes, scale, interval=4):
    return np.stack([sr_pipeline(f, luts, stages, modes, scale, interval) for f in frames_u8])


def four_simplex_interp(weight, img_in, h, w, interval, rot, upscale=4, mode="s"):
    """Call-compatible restatement of FourSimplexInterpFaster
    (4_test_lut.py:14-237): ``img_in`` is the already rotated and edge-padded
    (C, h+p, w+p) array; returns float64 (C, h*up, w*up) rotated by ``rot``
    quarter turns, divided by q."""
    check_mode(mode)
    q, _ = lut_geometry(interval)
    up = upscale
    lut = np.asarray(weight).reshape(-1, up * up).astype(np.int64)
    x = np.asarray(img_in).astype(np.int64)
    C = x.shape[0]
    t = np.stack([x[:, dy:dy + h, dx:dx + w] for dy, dx in MODE_TAPS[mode]], axis=0)
    verts, wts, _ = simplex_vertices(t, interval)
    o = np.zeros((C, h, w, up * up), dtype=np.int64)
    for k in range(5):
        o += wts[k][..., None] * lut[verts[k]]
    o = o.reshape(C, h, w, up, up).transpose(0, 1, 3, 2, 4).reshape(C, h * up, w * up)
    o = np.rot90(o, rot, [1, 2])
    return o.astype(np.float64) / q


# ---------------------------------------------------------------------------
# LUT file naming (the on-disk format is the reference's, unchanged)
# ---------------------------------------------------------------------------
def lut_file_name(lut_name: str, scale: int, interval: int, stage: int, mode: str) -> str:
    """4_test_lut.py:331-332 (note: test path names files with 8-interval)."""
    return "{}_x{}_{}bit_int8_s{}_{}.npy".format(lut_name, scale, 8 - interval, stage, mode)


def random_luts(seed: int, stages: int, modes, scale: int, interval: int = 4):
    """Seeded random int8 LUTs in the shipped layout."""
    rng = np.random.default_rng(seed)
    _, L = lut_geometry(interval)
    luts = {}
    for s in range(stages):
        cols = scale * scale if (s + 1) == stages else 1
        for mode in modes:
            luts["s{}_{}".format(s + 1, mode)] = rng.integers(-127, 128, (L ** 4, cols), dtype=np.int8)
    return luts
