"""Aggregate agreement of the CUDA path with the oracle over the cases of tests/test_gpu_configs.py::test_tracking_parameter_sweep:
flags and positions of fb_tracking! and optflow! summed over all keypoints of seeds 0 .. n-1 (north_star: flags >= 99.9 %, positions
within 0.01 px)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, slamklt
from slamklt import synth
from oracle import oracle as O

n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 300
ctx = slamklt.Context(0)
T = dict(points=0, fb_flag_diff=0, fb_fwd_flag_diff=0, fb_both=0, fb_ge_001=0, fb_max=0.0, of_flag_diff=0, of_both=0, of_ge_001=0, of_max=0.0)
for seed in range(n_seeds):
    rng = np.random.default_rng(4200 + seed)
    H, W = int(rng.integers(60, 260)), int(rng.integers(80, 420))
    levels = int(rng.integers(0, 4))
    while min(H, W) >> levels < 8:
        levels -= 1
    window = int(rng.choice([3, 3, 3, 4, 5, 7, 9, 9, 11, 15])) if seed else 9
    iterations = int(rng.choice([1, 3, 10, 30]))
    eps = float(rng.choice([1e-3, 1e-2, 5e-2]))
    thr = float(rng.choice([1e-6, 1e-4, 1e-3]))
    max_distance = float(rng.choice([0.25, 0.5, 1.0, 2.0]))
    fr, _ = synth.make_sequence(900 + seed, 2, H=H, W=W)
    f = synth.to_f64(fr)
    n = 500
    pts = synth.random_keypoints(50 + seed, n, H, W, border=0.0)
    pts[:8] = [[1, 1], [H, W], [1, W], [H, 1], [1.49, 1.51], [H - 0.5, W - 0.5], [H / 2, 1.0], [1.0, W / 2]]
    disp = rng.uniform(-1.5, 1.5, (n, 2)) if seed % 2 else None
    o0, o1 = O.LKPyramid(f[0], levels), O.LKPyramid(f[1], levels)
    o1.update(f[1])
    g0, g1 = slamklt.LKPyramid(ctx, f[0], levels), slamklt.LKPyramid(ctx, f[1], levels)
    g1.update(f[1])
    kw = dict(iterations=iterations, window_size=window, pyramid_levels=levels, max_distance=max_distance, eigenvalue_threshold=thr, eps=eps)
    po, so, fo = O.fb_tracking(o0, o1, pts, displacement=None if disp is None else disp.copy(), **kw)
    pg, sg, fg = slamklt.fb_tracking(g0, g1, pts, displacement=None if disp is None else disp.copy(), **kw)
    so, sg, fo, fg = (np.asarray(a, bool) for a in (so, sg, fo, fg))
    T["points"] += n
    T["fb_flag_diff"] += int(np.sum(so != sg)); T["fb_fwd_flag_diff"] += int(np.sum(fo != fg))
    both = so & sg
    if both.any():
        d = np.abs(po[both] - pg[both]).max(axis=1)
        T["fb_both"] += int(both.sum()); T["fb_ge_001"] += int(np.sum(d >= 0.01)); T["fb_max"] = max(T["fb_max"], float(d.max()))
    d0 = np.zeros((n, 2)) if disp is None else disp
    lk = dict(iterations=iterations, window_size=window, pyramid_levels=levels, eigenvalue_threshold=thr, eps=eps)
    do, so2 = O.optflow(d0.copy(), o0, o1, pts, O.LucasKanade(**lk))[:2]
    dg, sg2 = slamklt.optflow(d0.copy(), g0, g1, pts, slamklt.LucasKanade(**lk))[:2]
    so2, sg2 = np.asarray(so2, bool), np.asarray(sg2, bool)
    T["of_flag_diff"] += int(np.sum(so2 != sg2))
    ok = so2 & sg2
    if ok.any():
        dd = np.abs(np.asarray(do)[ok] - np.asarray(dg)[ok]).max(axis=1)
        T["of_both"] += int(ok.sum()); T["of_ge_001"] += int(np.sum(dd >= 0.01)); T["of_max"] = max(T["of_max"], float(dd.max()))
P = T["points"]
print(f"{n_seeds} random configurations, {P} keypoints")
print(f"fb_tracking!: status flags differ on {T['fb_flag_diff']} ({100 * (1 - T['fb_flag_diff'] / P):.4f} % agree), forward flags differ on {T['fb_fwd_flag_diff']} "
      f"({100 * (1 - T['fb_fwd_flag_diff'] / P):.4f} % agree); tracked by both {T['fb_both']}: {T['fb_ge_001']} positions >= 0.01 px apart "
      f"({100 * (1 - T['fb_ge_001'] / max(1, T['fb_both'])):.4f} % within), max {T['fb_max']:.4f} px")
print(f"optflow!: status flags differ on {T['of_flag_diff']} ({100 * (1 - T['of_flag_diff'] / P):.4f} % agree); ok in both {T['of_both']}: {T['of_ge_001']} displacements >= 0.01 px apart "
      f"({100 * (1 - T['of_ge_001'] / max(1, T['of_both'])):.4f} % within), max {T['of_max']:.4f} px")
