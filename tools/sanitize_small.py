"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): every kernel family on tiny inputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import slamklt
from slamklt import synth

fr, aff = synth.make_sequence(1, 3, H=75, W=131)
f = synth.to_f64(fr)
ctx = slamklt.Context(0)
for levels, shape in ((2, (75, 131)), (1, (33, 47)), (2, (72, 128))):  # the last one takes the group-aligned column kernels
    a = slamklt.LKPyramid(ctx, f[0][:shape[0], :shape[1]], levels)
    b = slamklt.LKPyramid(ctx, fr[1][:shape[0], :shape[1]], levels)
    b.update(f[1][:shape[0], :shape[1]])
    pts = np.vstack([synth.random_keypoints(3, 40, shape[0], shape[1], border=1.0), [[1.0, 1.0], [shape[0], shape[1]], [shape[0] - 0.5, 2.0]]])
    for w in (9, 11, 15):
        slamklt.fb_tracking(a, b, pts, window_size=w, pyramid_levels=levels, max_distance=1.0)
    slamklt.optflow(np.zeros_like(pts), a, b, pts, slamklt.LucasKanade(pyramid_levels=levels))
    slamklt.optical_flow_matching(a, b, pts, np.ones_like(pts), np.arange(len(pts)) % 2, pyramid_levels=levels, pyramid_levels_3d=1)
    a.plane(1, "Syy"); a.plane(0, "Iyy"); a.plane(0, "Ix")
    sc = synth.matching_scene(4, pts, pts + 0.7, camera=dict(synth.KITTI_CAMERA, fx=90.0, fy=90.0, cx=shape[1] / 2, cy=shape[0] / 2,
                                                              height=shape[0], width=shape[1]), baseline=0.3)
    cam = slamklt.Camera(**sc["camera"]); rcam = slamklt.Camera(**sc["camera"], Ti0=sc["Ti0"])
    slamklt.optical_flow_matching_frame(a, b, pts, sc["is_3d"], sc["world"], sc["cw"], cam, pyramid_levels=levels)
    slamklt.optical_flow_matching_frame(a, b, pts, sc["is_3d"], sc["world"], sc["cw"], cam, right_camera=rcam, undistorted=pts,
                                        stereo=True, pyramid_levels=levels)
# more keypoints than warp slots (148 x 16): the tracking kernel runs as a persistent grid drawing indices from the work counter
many = synth.random_keypoints(11, 3000, 72, 128, border=1.0)
slamklt.fb_tracking(a, b, many, window_size=9, pyramid_levels=2, max_distance=1.0)
e = slamklt.Extractor(100, 8, (3, 4), 35)
slamklt.detect(ctx, e, f[0], np.array([[10.0, 10.0], [60.0, 100.0]]))
slamklt.detect(ctx, e, fr[0], np.zeros((0, 2)))
batch = slamklt.StreamBatch(ctx, 75, 131, 2, 2, 50)
batch.prime(f[0])
pts = np.stack([synth.random_keypoints(5 + i, 50, 75, 131, border=2.0) for i in range(2)])
batch.step(slamklt.StreamBatch.pack_frames(f[1:3]), pts, slamklt.LucasKanade(pyramid_levels=2))
batch.upload(slamklt.StreamBatch.pack_frames(fr[1:3]), pts); batch.process(slamklt.LucasKanade(pyramid_levels=2)); batch.download()
batch.detect(e)
# round 2: masked detect through the register-tiled kernel (row binning, chunks of current points), a cell size other than 35, a
# mask blur that takes the first kernel; two batches in flight with page-locked Float64 frames (two upload engines); detect after
# such a step; stereo matching of two batches
cur = np.stack([synth.random_keypoints(40 + i, 30, 75, 131, border=0.0) for i in range(2)])
batch.detect(e, cur)
slamklt.detect(ctx, slamklt.Extractor(100, 5, (4, 6), 24), f[0], cur[0])
slamklt.detect(ctx, e, f[0], cur[0], sigma_mask=2.0)
slamklt.detect(ctx, e, fr[0], cur[0])
nb = 8
frs, _ = synth.make_sequence(2, nb + 1, H=376, W=400)
fs64 = synth.to_f64(frs)
pin = slamklt.PinnedArray((nb, 400, 376), np.float64)
pin.array[...] = np.transpose(fs64[1:], (0, 2, 1))
bx, by = slamklt.StreamBatch(ctx, 376, 400, 2, nb, 60), slamklt.StreamBatch(ctx, 376, 400, 2, nb, 60)
bx.prime(fs64[0]); by.prime(fs64[0])
p8 = np.stack([synth.random_keypoints(70 + i, 60, 376, 400, border=3.0) for i in range(nb)])
alg2 = slamklt.LucasKanade(pyramid_levels=2)
for _ in range(2):
    bx.step_begin(pin.array, p8, alg2); by.step_begin(pin.array, p8, alg2)
    bx.step_end(); by.step_end()
bx.detect(slamklt.Extractor(300, 8, (11, 12), 35), np.stack([synth.random_keypoints(90 + i, 40, 376, 400, border=0.0) for i in range(nb)]))
by.track_cross(bx, alg2); bx.download()
bx.close(); by.close(); pin.free()
# round 2c: frames taller than 512 rows (two warps per column: 12 and 20 rows per lane, odd height = per-row predicates, every
# source type; level 0 up to 1280 rows) and levels beyond the tiled kernels (general per-line kernels: tall, wide), tracking on such pyramids
for (Ht, Wt, src_kind) in ((600, 40, "f64"), (1080, 24, "u8"), (771, 20, "f32"), (1130, 36, "f64"), (1280, 24, "u8"), (1400, 28, "f64"), (40, 2080, "u8")):
    frt, _ = synth.make_sequence(9, 2, H=Ht, W=Wt)
    ft = synth.to_f64(frt)
    srct = {"f64": ft, "u8": frt, "f32": ft.astype(np.float32)}[src_kind]
    pa = slamklt.LKPyramid(ctx, srct[0], 2)
    pb = slamklt.LKPyramid(ctx, srct[1], 2); pb.update(srct[1])
    ptst = synth.random_keypoints(21, 60, Ht, Wt, border=1.0)
    slamklt.fb_tracking(pa, pb, ptst, window_size=9, pyramid_levels=2, max_distance=1.0)
    pb.plane(0, "Sxx"); pb.plane(0, "Ryx")
batch.close(); ctx.close()
print("sanitize_small done")
