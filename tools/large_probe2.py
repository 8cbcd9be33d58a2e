import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, slamklt
from slamklt import synth
ctx = slamklt.Context(0)
for (H, W, L) in ((1920, 1080, 5), (2160, 3840, 5)):
    fr, _ = synth.make_sequence(1, 2, H=H, W=W)
    f = synth.to_f64(fr)
    b = slamklt.LKPyramid(ctx, fr[1], L)
    b.update(fr[1]); ctx.sync()
    ctx.profile(True)
    for _ in range(3): b.update(fr[1])
    rep = ctx.profile_report(); ctx.profile(False)
    print(H, W, {k: v for k, v in rep.items()})
    e = slamklt.Extractor(8000, 17, (H // 35, W // 35), 35)
    pts = synth.random_keypoints(3, 2000, H, W)
    slamklt.detect(ctx, e, fr[1], pts); ctx.sync()
    for src, nm in ((fr[1], "u8"), (f[1], "f64")):
        t = time.perf_counter()
        for _ in range(5): slamklt.detect(ctx, e, src, pts)
        ctx.sync(); print("detect", nm, (time.perf_counter() - t) / 5 * 1e3, "ms")
    ctx.profile(True)
    slamklt.detect(ctx, e, fr[1], pts)
    print(ctx.profile_report()); ctx.profile(False)
