"""Run one stage of the headline step in isolation (for ncu captures and quick A/B timings).
    python tools/stage_bench.py track|build|step [iters]
64 KITTI-shaped frames x 2000 keypoints, device-resident (the bench.py workload)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, slamklt
from slamklt import synth

stage = sys.argv[1] if len(sys.argv) > 1 else "track"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
nf = int(os.environ.get("STAGE_FRAMES", "64"))
SH, SW = int(os.environ.get("STAGE_H", "376")), int(os.environ.get("STAGE_W", "1241"))   # frame size (default: KITTI)
fr, aff = synth.make_sequence(2000, nf + 1, H=SH, W=SW)
f64 = synth.to_f64(fr)
ctx = slamklt.Context(0)
batch = slamklt.StreamBatch(ctx, SH, SW, 3, nf, 2000)
batch.prime(f64[0])
e = slamklt.Extractor(6 * (SH // 35 + 1) * (SW // 35 + 1), 17, (SH // 35 + 1, SW // 35 + 1), 35)
batch.upload(slamklt.StreamBatch.pack_frames(f64[:-1]), np.zeros((nf, 1, 2)) + 5)
kps = batch.detect(e)
rng = np.random.default_rng(7)
pts = np.empty((nf, 2000, 2))
for i, kp in enumerate(kps):
    kp = kp.astype(np.float64)[:2000]
    if len(kp) < 2000:
        kp = np.vstack([kp, synth.random_keypoints(1000 + i, 2000 - len(kp), SH, SW)])
    pts[i] = np.clip(kp + rng.uniform(-0.5, 0.5, kp.shape), 1.0, [SH, SW])
# STAGE_U8=1: the frames sit on the device as UInt8 (what a step leaves there when it repacks 8-bit data) instead of Float64
batch.upload(slamklt.StreamBatch.pack_frames(fr[1:] if os.environ.get("STAGE_U8") else f64[1:]), pts)
alg = slamklt.LucasKanade(iterations=30, window_size=9, pyramid_levels=3)
batch.build(); batch.track(alg, 1.0); ctx.sync()
ctx.stats(reset=True)
ctx.timer_start()
for _ in range(iters):
    if stage in ("build", "step"): batch.build()
    if stage in ("track", "step"): batch.track(alg, 1.0)
ms = ctx.timer_stop() / iters
st = ctx.stats()
_, status = batch.download()
print(f"{stage}: {ms:.4f} ms per pass, tracked {(status & 1).mean():.4f}, iters/kp {st['lk_iters'] / max(1, iters * nf * 2000):.2f}")
batch.close(); ctx.close()
