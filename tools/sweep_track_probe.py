"""Outliers of one case of tests/test_gpu_configs.py::test_tracking_parameter_sweep (seed on the command line): where they sit and how
well conditioned their structure tensor is."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, slamklt
from slamklt import synth
from oracle import oracle as O

ctx = slamklt.Context(0)
for seed in [int(a) for a in sys.argv[1:]] or [270, 293]:
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
    print(f"seed {seed}: {H}x{W} {kw} disp={'yes' if disp is not None else 'no'}")
    po, so, fo = O.fb_tracking(o0, o1, pts, displacement=None if disp is None else disp.copy(), **kw)
    pg, sg, fg = slamklt.fb_tracking(g0, g1, pts, displacement=None if disp is None else disp.copy(), **kw)
    both = np.asarray(so, bool) & np.asarray(sg, bool)
    d = np.full(n, 0.0); d[both] = np.abs(po[both] - pg[both]).max(axis=1)
    print(f"  tracked by both {both.sum()}, flags differ {np.sum(np.asarray(so, bool) != np.asarray(sg, bool))}, >= 0.01 px: {np.sum(d >= 0.01)}, median {np.median(d[both]):.2e}")
    # optflow! with the same parameters: no forward-backward gate, so ill-conditioned points stay in the comparison
    d0 = np.zeros((n, 2)) if disp is None else disp
    lk = dict(iterations=iterations, window_size=window, pyramid_levels=levels, eigenvalue_threshold=thr, eps=eps)
    do, so2 = O.optflow(d0.copy(), o0, o1, pts, O.LucasKanade(**lk))[:2]
    dg, sg2 = slamklt.optflow(d0.copy(), g0, g1, pts, slamklt.LucasKanade(**lk))[:2]
    ok2 = np.asarray(so2, bool) & np.asarray(sg2, bool)
    d = np.full(n, 0.0); d[ok2] = np.abs(np.asarray(do)[ok2] - np.asarray(dg)[ok2]).max(axis=1)
    po, pg = pts + np.asarray(do), pts + np.asarray(dg)
    print(f"  optflow!: ok in both {ok2.sum()}, flags differ {np.sum(np.asarray(so2, bool) != np.asarray(sg2, bool))}, >= 0.01 px: {np.sum(d >= 0.01)}; of those the forward-backward gate keeps {int(np.sum((d >= 0.01) & both))}")
    S = {k: o0.plane(0, k) for k in ("Syy", "Sxx", "Syx")}
    for i in np.flatnonzero(d >= 0.005):
        y, x = int(np.floor(pts[i, 0])), int(np.floor(pts[i, 1]))
        y0, y1, x0, x1 = max(1, y - window), min(H, y + window), max(1, x - window), min(W, x + window)
        G = np.array([[S["Syy"][y0 - 1:y1, x0 - 1:x1].sum(), S["Syx"][y0 - 1:y1, x0 - 1:x1].sum()],
                      [S["Syx"][y0 - 1:y1, x0 - 1:x1].sum(), S["Sxx"][y0 - 1:y1, x0 - 1:x1].sum()]])
        ev = np.linalg.eigvalsh(G)
        npx = (y1 - y0 + 1) * (x1 - x0 + 1)
        print(f"  point {i}: {pts[i]} -> oracle {po[i]} gpu {pg[i]} |d| {d[i]:.4f}; level-0 G eigenvalues / window px {ev / npx} (threshold {thr}), cond {ev[1] / max(ev[0], 1e-300):.1f}")
