import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, slamklt
from slamklt import synth
fr, aff = synth.make_sequence(2000, 65)
f64 = synth.to_f64(fr)
ctx = slamklt.Context(0)
batch = slamklt.StreamBatch(ctx, 376, 1241, 3, 64, 2000)
batch.prime(f64[0])
e = slamklt.Extractor(2376, 17, (11, 36), 35)
cur = np.stack([synth.random_keypoints(i, 1000, 376, 1241) for i in range(64)])
for name, frames in (("f64", f64[1:]), ("u8", fr[1:])):
    batch.upload(slamklt.StreamBatch.pack_frames(frames), np.zeros((64, 1, 2)) + 5)
    for c in (None, cur):
        batch.detect(e, c)
        ctx.profile(True)
        t0 = time.perf_counter()
        for _ in range(5): out = batch.detect(e, c)
        dt = (time.perf_counter() - t0) / 5
        rep = ctx.profile_report(); ctx.profile(False)
        print(name, "masked" if c is not None else "plain", "wall %.2f ms" % (dt * 1e3), {k: round(v[1] / v[0], 3) for k, v in rep.items()}, sum(len(o) for o in out))
