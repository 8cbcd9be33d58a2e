"""Timing of the drop-in calls on frames beyond the register-tiled kernels (general pyramid kernels), host wall clock per call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, slamklt
from slamklt import synth

ctx = slamklt.Context(0)
for (H, W, L) in ((1080, 1920, 5), (1920, 1080, 5), (2160, 3840, 5)):
    fr, _ = synth.make_sequence(1, 2, H=H, W=W)
    f = [np.asfortranarray(x) for x in synth.to_f64(fr)]   # column-major like Julia's matrices: no layout copy in the mirror
    fr = [np.asfortranarray(x) for x in fr]
    a, b = slamklt.LKPyramid(ctx, f[0], L), slamklt.LKPyramid(ctx, f[1], L)
    pts = synth.random_keypoints(3, 8000, H, W)
    e = slamklt.Extractor(8000, 17, (H // 35, W // 35), 35)
    res = {}
    for name, fn in (("update_f64", lambda: b.update(f[1])), ("update_u8", lambda: b.update(fr[1])),
                     ("fb_tracking_8000kp", lambda: slamklt.fb_tracking(a, b, pts, window_size=9, pyramid_levels=L, max_distance=1.0)),
                     ("detect", lambda: slamklt.detect(ctx, e, f[1], pts[:2000]))):
        fn(); ctx.sync()
        if name == "update_u8":
            ctx.profile(True); fn(); rep = ctx.profile_report(); ctx.profile(False)
            res["build_kernels_ms"] = round(sum(v[1] for v in rep.values()), 3)
            res["general_kernels_ms"] = round(sum(v[1] for k, v in rep.items() if k.startswith("k_gen") or k in ("k_convert", "k_resize_L0")), 3)
        t = time.perf_counter()
        for _ in range(10): fn()
        ctx.sync()
        res[name] = round((time.perf_counter() - t) / 10 * 1e3, 3)
    print(f"{H}x{W} L={L}: {res}")
