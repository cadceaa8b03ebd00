"""One-off randomised parity sweep of the x2 inference kernels (K1g + K1f / K1c) against the C oracle:

    python tests/fuzz_infer.py <first seed> <last seed> [scale=2]      # on the GPU box

Random C in 1..4, TMA-mappable and arbitrary widths, 1-3 stages, mode subsets, five value distributions, orphaning on/off.
Round 1: scale 2 seeds 0..699, scales 1, 3 and 4 seeds 0..119, both kernel selections: 0 mismatches.
Round 2 (HEAD, after the K1b / K1f / K0 / K1e3 rewrites): scale 2 seeds 0..399, scales 1, 3 and 4 seeds 0..99: 0 mismatches."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mulut_b200.infer import LutEngine
from oracle import c_oracle as CO, mulut_oracle as O
bad = 0
SCALE = int(sys.argv[3]) if len(sys.argv) > 3 else 2
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.default_rng(5000 + seed)
    C = int(rng.choice([1, 2, 3, 4]))
    W = int(rng.integers(1, 40)) * (16 // np.gcd(16, C))
    if (W * C) % 16: W *= 16
    if seed % 3 == 0: W = 1 + (W * 7 + seed) % 211          # any width: rows TMA cannot map in place go through the pitched copy
    H = int(rng.integers(1, 300)); N = int(rng.integers(1, 5)); stages = int(rng.integers(1, 4))
    modes = ["sdy", "s", "dy", "ys", "d"][int(rng.integers(0, 5))]
    os.environ["MULUT_BN_ORPHANS"] = str(seed % 2)
    kind = seed % 5
    shp = (N, H, W, C)
    if kind == 0: img = rng.integers(0, 256, shp)
    elif kind == 1:
        lo = int(rng.integers(0, 200)); img = rng.integers(lo, lo + int(rng.integers(2, 56)), shp)
    elif kind == 2: img = np.where(rng.random(shp) < 0.5, rng.integers(0, 64, shp), rng.integers(192, 256, shp))
    elif kind == 3: img = np.clip(rng.normal(128, 20, shp) + (rng.random(shp) < 0.01) * rng.normal(0, 120, shp), 0, 255)
    else: img = np.full(shp, int(rng.integers(0, 256)))
    img = np.ascontiguousarray(img, dtype=np.uint8)
    luts = O.random_luts(900 + seed, stages, modes, SCALE)
    ref = CO.sr_u8(img, luts, stages, modes, SCALE)
    for kernel in (3, -1):
        with LutEngine(luts, stages, modes, SCALE, 4, device=0, kernel=kernel) as eng:
            out = eng(torch.from_numpy(img).cuda()).cpu().numpy()
        if not (out == ref).all():
            bad += 1
            print("MISMATCH", seed, kernel, shp, stages, modes, int((out != ref).sum()))
print("fuzz done", sys.argv[1], sys.argv[2], "scale", SCALE, "bad =", bad)
