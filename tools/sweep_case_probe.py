"""Per-plane error of one case of tests/test_gpu_configs.py::test_pyramid_parameter_sweep (seed on the command line)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, slamklt
from oracle import oracle as O

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 193
rng = np.random.default_rng(7300 + seed)
H = int(rng.choice([rng.integers(16, 64), rng.integers(64, 200), rng.integers(200, 420), rng.integers(420, 800), rng.integers(800, 1300)]))
W = int(rng.choice([rng.integers(16, 120), rng.integers(120, 700), rng.integers(700, 1400), rng.integers(1400, 2300)]))
levels = int(rng.integers(0, 5))
while levels > 0 and min((H + (1 << levels) - 1) >> levels, (W + (1 << levels) - 1) >> levels) < 4:
    levels -= 1
sigma = float(rng.choice([0.8, 1.0, 1.0, 1.5, 2.0]))
u8 = rng.integers(0, 256, (2, H, W)).astype(np.uint8)
if seed % 3 == 0:
    yy, xx = np.mgrid[0:H, 0:W]
    u8 = np.stack([(127 + 100 * np.sin(0.05 * yy + k) * np.cos(0.031 * xx)).astype(np.uint8) for k in range(2)])
f64 = u8.astype(np.float64) / 255.0
ctx = slamklt.Context(0)
op = O.LKPyramid(f64[0], levels, sigma=sigma, mode="ctor")
gp = slamklt.LKPyramid(ctx, f64[0], levels, sigma=sigma)
print(f"seed {seed}: {H}x{W}, levels {levels}, sigma {sigma}")
for tag in ("ctor", "update"):
    if tag == "update":
        op.update(f64[1], sigma=sigma); gp.update(f64[1], sigma=sigma)
    for l in range(levels + 1):
        row = []
        for name in ("layer", "Iy", "Ix", "Syy", "Sxx", "Syx"):
            a, b = gp.plane(l, name), op.plane(l, name)
            row.append(f"{name} {np.abs(a - b).max() / np.abs(b).max():.2e} (max|b| {np.abs(b).max():.2e})")
        print(tag, l, " | ".join(row))
