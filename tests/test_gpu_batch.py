"""GPU tests of the batched stream API (configs 2/4/5): the batch path must give exactly what the single-pair entry
points give (same kernels), which in turn are parity-checked against the oracle in test_gpu_parity.py."""
import numpy as np
import pytest

import slamklt
from slamklt import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small_seq():
    fr, aff = synth.make_sequence(909, 9, H=120, W=200)
    return fr, synth.to_f64(fr), aff


@pytest.mark.parametrize("dtype", ["f64", "u8"])
def test_batch_step_matches_single_calls_and_oracle(ctx, small_seq, dtype):
    fr, f64, aff = small_seq
    H, W, L, NF, NP = 120, 200, 2, 4, 150
    alg = slamklt.LucasKanade(pyramid_levels=L, window_size=9)
    batch = slamklt.StreamBatch(ctx, H, W, L, NF, NP)
    src = f64 if dtype == "f64" else fr
    batch.prime(src[0], mode=slamklt.MODE_UPDATE)
    pts = np.stack([synth.random_keypoints(50 + i, NP, H, W, border=2.0) for i in range(8)])
    singles = [slamklt.LKPyramid(ctx, f64[0], L)]
    singles[0].update(f64[0])
    for step in range(2):  # two consecutive steps: the second relies on slamklt_batch_rotate carrying the last frame
        frames = src[1 + step * NF: 1 + (step + 1) * NF]
        p = pts[step * NF:(step + 1) * NF]
        out, st = batch.step(slamklt.StreamBatch.pack_frames(frames), p, alg, max_distance=1.0)
        for i in range(NF):
            t = step * NF + i
            cur = slamklt.LKPyramid(ctx, f64[t + 1], L)
            cur.update(f64[t + 1])
            singles.append(cur)
            ref_pts, ref_st, ref_fst = slamklt.fb_tracking(singles[t], cur, p[i], window_size=9, pyramid_levels=L, max_distance=1.0)
            assert np.array_equal(st[i] & 1, ref_st.astype(np.uint8))
            assert np.array_equal((st[i] >> 1) & 1, ref_fst.astype(np.uint8))
            assert np.array_equal(out[i][ref_fst], ref_pts[ref_fst])  # same kernels => bit-identical
            assert np.all(np.isnan(out[i][~ref_fst]))
        # slot planes equal the single pyramids (slot 0 now holds the last frame of this step)
        a = batch.slot(0).plane(1, "layer")
        assert np.array_equal(a, singles[(step + 1) * NF].plane(1, "layer"))
    # and against the oracle for the very last pair
    o0 = O.LKPyramid(f64[7], L); o0.update(f64[7]); o1 = O.LKPyramid(f64[8], L); o1.update(f64[8])
    po, so, fo = O.fb_tracking(o0, o1, pts[7], window_size=9, pyramid_levels=L, max_distance=1.0)
    assert np.mean(so == (st[3] & 1).astype(bool)) >= 0.99
    both = so & (st[3] & 1).astype(bool)
    assert np.mean(np.abs(po[both] - out[3][both]).max(axis=1) < 0.01) >= 0.99
    batch.close()


def test_batch_phases_equal_step(ctx, small_seq):
    fr, f64, aff = small_seq
    H, W, L, NF, NP = 120, 200, 2, 4, 100
    alg = slamklt.LucasKanade(pyramid_levels=L)
    pts = np.stack([synth.random_keypoints(80 + i, NP, H, W, border=4.0) for i in range(NF)])
    packed = slamklt.StreamBatch.pack_frames(f64[1:1 + NF])
    b1 = slamklt.StreamBatch(ctx, H, W, L, NF, NP); b1.prime(f64[0])
    b2 = slamklt.StreamBatch(ctx, H, W, L, NF, NP); b2.prime(f64[0])
    o1, s1 = b1.step(packed, pts, alg)
    b2.upload(packed, pts); b2.build(); b2.track(alg, 1.0)
    o2, s2 = b2.download()
    assert np.array_equal(s1, s2) and np.array_equal(np.nan_to_num(o1), np.nan_to_num(o2))
    b3 = slamklt.StreamBatch(ctx, H, W, L, NF, NP); b3.prime(f64[0])
    b3.upload(packed, pts); b3.process(alg, 1.0); b3.process(alg, 1.0)   # repeated: same slots rebuilt, same answer
    o3, s3 = b3.download()
    assert np.array_equal(s1, s3) and np.array_equal(np.nan_to_num(o1), np.nan_to_num(o3))
    b3.close()
    assert ctx.stats()["kernel_launches"] > 0
    b1.close(); b2.close()


def test_batch_detect_matches_single(ctx, small_seq):
    fr, f64, aff = small_seq
    H, W, NF = 120, 200, 4
    e = slamklt.Extractor(300, 8, (4, 6), 35)
    batch = slamklt.StreamBatch(ctx, H, W, 2, NF, 10)
    cur = np.stack([synth.random_keypoints(5 + i, 20, H, W) for i in range(NF)])
    for src in (f64, fr):
        batch.upload(slamklt.StreamBatch.pack_frames(src[:NF]), np.zeros((NF, 1, 2)) + 5)
        got = batch.detect(e, cur)
        for i in range(NF):
            assert np.array_equal(got[i], slamklt.detect(ctx, e, f64[i], cur[i]))
            assert np.array_equal(got[i], O.detect(O.Extractor(300, 8, (4, 6), 35), f64[i], cur[i]))
    batch.close()


def test_pyramid_copy_clone_swap(ctx, small_seq):
    _, f64, _ = small_seq
    a = slamklt.LKPyramid(ctx, f64[0], 2)
    b = slamklt.LKPyramid(ctx, f64[1], 2)
    la, lb = a.plane(2, "layer"), b.plane(2, "layer")
    c = a.deepcopy()                      # SLAM.jl:216-219
    assert np.array_equal(c.plane(2, "layer"), la) and c.has_gradients()
    a.swap(b)                             # front_end.jl:459-461 as a pointer swap
    assert np.array_equal(a.plane(2, "layer"), lb) and np.array_equal(b.plane(2, "layer"), la)
    b.copy_from(a)                        # pyramid.jl:28-38
    assert np.array_equal(b.plane(1, "Ix"), a.plane(1, "Ix"))
    empty = slamklt.LKPyramid(ctx, None, 2, shape=(120, 200))
    assert not empty.has_gradients()
    with pytest.raises(slamklt.SlamKltError):
        slamklt.fb_tracking(empty, a, np.array([[30.0, 30.0]]))


def test_two_host_threads_share_a_context(ctx, small_seq):
    """The mapper task calls into the path concurrently with the front-end task (mapper.jl:51-60); calls on one context are
    serialised by its mutex and must give the same answers as serial calls."""
    import threading
    _, f64, _ = small_seq
    pts = synth.random_keypoints(123, 200, 120, 200, border=3.0)
    ref_a = slamklt.LKPyramid(ctx, f64[0], 2); ref_b = slamklt.LKPyramid(ctx, f64[2], 2)
    ref_b.update(f64[1])  # same build path (update!) as the workers
    ref = slamklt.fb_tracking(ref_a, ref_b, pts, window_size=9, pyramid_levels=2, max_distance=1.0)
    results, errors = {}, []

    def worker(tid):
        try:
            a = slamklt.LKPyramid(ctx, f64[0], 2)
            b = slamklt.LKPyramid(ctx, f64[2], 2)
            for _ in range(10):
                b.update(f64[1])
                results[tid] = slamklt.fb_tracking(a, b, pts, window_size=9, pyramid_levels=2, max_distance=1.0)
                b.update(f64[2])
        except Exception as e:  # pragma: no cover
            errors.append(e)

    ths = [threading.Thread(target=worker, args=(i,)) for i in range(4)]
    for t in ths: t.start()
    for t in ths: t.join()
    assert not errors
    for tid in range(4):
        assert np.array_equal(results[tid][1], ref[1])
        assert np.array_equal(results[tid][0][ref[1]], ref[0][ref[1]])


def test_full_size_batch64_against_oracle(ctx):
    """BASELINE.json configs[1] at full size: 64 frames of 1241x376, 2000 keypoints each, through slamklt_batch_step;
    three of the 64 pairs are re-done by the Float64 oracle (north_star tolerances), and the batch is self-consistent:
    tracking frame i -> i+1 and back lands on the start (forward-backward property) for every accepted point."""
    fr, aff = synth.make_sequence(2000, 65)
    f64 = synth.to_f64(fr)
    H, W, L, NF, NP = 376, 1241, 3, 64, 2000
    alg = slamklt.LucasKanade(pyramid_levels=L, window_size=9)
    batch = slamklt.StreamBatch(ctx, H, W, L, NF, NP)
    batch.prime(f64[0], mode=slamklt.MODE_UPDATE)
    ext = slamklt.Extractor(2376, 17, (11, 36), 35)
    batch.upload(slamklt.StreamBatch.pack_frames(fr[:NF]), np.zeros((NF, 1, 2)) + 5)
    kps = batch.detect(ext)
    pts = np.stack([np.vstack([k.astype(np.float64), synth.random_keypoints(i, NP, H, W)])[:NP] + 0.25 for i, k in enumerate(kps)])
    batch.prime(f64[0], mode=slamklt.MODE_UPDATE)
    out, st = batch.step(slamklt.StreamBatch.pack_frames(fr[1:]), pts, alg, max_distance=1.0)   # u8 frames in
    ok = (st & 1).astype(bool)
    assert ok.mean() > 0.85
    gt = np.stack([synth.true_flow(aff, i, i + 1, pts[i]) for i in range(NF)])
    assert np.median(np.linalg.norm(out[ok] - gt[ok], axis=1)) < 0.15
    for i in (0, 31, 63):
        o0 = O.LKPyramid(f64[i], L); o0.update(f64[i])
        o1 = O.LKPyramid(f64[i + 1], L); o1.update(f64[i + 1])
        po, so, fo = O.fb_tracking(o0, o1, pts[i], window_size=9, pyramid_levels=L, max_distance=1.0)
        assert np.mean(so == ok[i]) >= 0.999
        both = so & ok[i]
        d = np.abs(po[both] - out[i][both]).max(axis=1)
        assert np.mean(d < 0.01) >= 0.999 and d.max() < 0.02
    batch.close()


def test_step_repacks_8bit_frames_losslessly_and_falls_back(ctx, monkeypatch):
    """slamklt_batch_step ships Float64 host frames whose pixels are all exact k/255 as 8 bits per pixel (lossless: the device
    rebuilds the identical Float64).  The result must be bit-identical to the plain Float64 upload (SLAMKLT_NO_PACK=1), frames
    that are not 8-bit data must take the Float64 path, and a batch that turns non-8-bit half way must switch over cleanly."""
    H, W, L, NF, NP = 376, 1241, 3, 16, 300       # 16 x 466 616 px: above the 1 Mpx threshold, 8 chunks of 2 frames
    fr, _ = synth.make_sequence(2020, NF + 1, H=H, W=W)
    f64 = synth.to_f64(fr)
    alg = slamklt.LucasKanade(pyramid_levels=L)
    pts = np.stack([synth.random_keypoints(300 + i, NP, H, W, border=4.0) for i in range(NF)])

    def run(frames, no_pack):
        if no_pack:
            monkeypatch.setenv("SLAMKLT_NO_PACK", "1")
        else:
            monkeypatch.delenv("SLAMKLT_NO_PACK", raising=False)
        b = slamklt.StreamBatch(ctx, H, W, L, NF, NP)
        b.prime(frames[0])
        h0 = ctx.stats()["h2d_bytes"]
        out, st = b.step(slamklt.StreamBatch.pack_frames(frames[1:]), pts, alg)
        sent = ctx.stats()["h2d_bytes"] - h0
        layer = b.slot(0).plane(2, "layer")       # after the rotation slot 0 is the last frame of the batch
        b.close()
        return out, st, sent, layer

    o_plain, s_plain, sent_plain, l_plain = run(f64, True)
    o_pack, s_pack, sent_pack, l_pack = run(f64, False)
    assert np.array_equal(s_plain, s_pack) and np.array_equal(np.nan_to_num(o_plain), np.nan_to_num(o_pack)) and np.array_equal(l_plain, l_pack)
    assert sent_plain > NF * H * W * 8 and sent_pack < NF * H * W * 1.2 + NF * NP * 16 + 1024
    assert (s_pack & 1).mean() > 0.8
    # not 8-bit data at all: plain upload, and the same answer as the explicit plain path
    g = f64 + 1e-9
    o1, s1, sent1, _ = run(g, False)
    o2, s2, sent2, _ = run(g, True)
    assert sent1 == sent2 and np.array_equal(s1, s2) and np.array_equal(np.nan_to_num(o1), np.nan_to_num(o2))
    # 8-bit until frame 9, then not: the first chunks go packed, the rest plain
    m = f64.copy(); m[10:] += 1e-9
    o3, s3, sent3, _ = run(m, False)
    o4, s4, sent4, _ = run(m, True)
    assert sent_pack < sent3 < sent4 and np.array_equal(s3, s4) and np.array_equal(np.nan_to_num(o3), np.nan_to_num(o4))


def test_step_hybrid_upload_and_two_batches_in_flight(ctx, monkeypatch):
    """Round 2: (1) page-locked Float64 host frames travel through two engines at once -- host threads repack some chunks to 8
    bits, the copy engine ships the others as plain Float64 (SLAMKLT_UPLOAD_PLAN forces the split here; by default it follows
    the two measured rates) -- and the result is bit-identical to the all-plain upload; (2) slamklt_batch_step_begin / _end keep
    two batches in flight and give what two synchronous steps give; (3) misuse is refused."""
    H, W, L, NF, NP = 376, 1241, 3, 16, 300
    fr, _ = synth.make_sequence(2021, NF + 1, H=H, W=W)
    f64 = synth.to_f64(fr)
    alg = slamklt.LucasKanade(pyramid_levels=L)
    pts = np.stack([synth.random_keypoints(500 + i, NP, H, W, border=4.0) for i in range(NF)])
    pin = slamklt.PinnedArray((NF, W, H), np.float64)
    pin.array[...] = np.transpose(f64[1:], (0, 2, 1))
    pin_rev = slamklt.PinnedArray((NF, W, H), np.float64)
    pin_rev.array[...] = np.transpose(f64[:-1][::-1], (0, 2, 1))

    def run(plan):
        for k in ("SLAMKLT_NO_PACK", "SLAMKLT_UPLOAD_PLAN", "SLAMKLT_NO_HYBRID"):
            monkeypatch.delenv(k, raising=False)
        if plan == "plain":
            monkeypatch.setenv("SLAMKLT_NO_PACK", "1")
        elif plan == "packed":
            monkeypatch.setenv("SLAMKLT_NO_HYBRID", "1")
        elif plan != "auto":
            monkeypatch.setenv("SLAMKLT_UPLOAD_PLAN", plan)
        b = slamklt.StreamBatch(ctx, H, W, L, NF, NP)
        b.prime(f64[0])
        h0 = ctx.stats()["h2d_bytes"]
        out, st = b.step(pin.array, pts, alg)
        sent = ctx.stats()["h2d_bytes"] - h0
        out2, st2 = b.step(pin_rev.array, pts[::-1], alg)     # a second step on the rotated batch (auto: measured rates now exist)
        layer = b.slot(0).plane(1, "layer")
        b.close()
        return np.nan_to_num(out), st, sent, np.nan_to_num(out2), st2, layer

    ref = run("plain")
    full = NF * H * W * 8
    for plan in ("packed", "PRPRPRRP", "RP", "RRRRRRRR", "auto"):
        got = run(plan)
        for a, b_ in zip(ref[:2] + ref[3:], got[:2] + got[3:]):
            assert np.array_equal(a, b_), plan
        if plan == "packed":
            assert got[2] < full * 0.2
        if plan == "PRPRPRRP":
            assert full * 0.5 < got[2] < full * 0.75       # 4 of 8 chunks plain + 4 packed
        if plan == "RRRRRRRR":
            assert got[2] > full
    for k in ("SLAMKLT_NO_PACK", "SLAMKLT_UPLOAD_PLAN", "SLAMKLT_NO_HYBRID"):
        monkeypatch.delenv(k, raising=False)

    # two batches in flight
    bx, by = slamklt.StreamBatch(ctx, H, W, L, NF, NP), slamklt.StreamBatch(ctx, H, W, L, NF, NP)
    bx.prime(f64[0]); by.prime(f64[NF])
    bx.step_begin(pin.array, pts, alg)
    by.step_begin(pin_rev.array, pts[::-1], alg)
    with pytest.raises(slamklt.SlamKltError):
        bx.step_begin(pin.array, pts, alg)                 # a step is already in flight on bx
    ox, sx = bx.step_end()
    oy, sy = by.step_end()
    with pytest.raises(slamklt.SlamKltError):
        bx.step_end()                                      # nothing in flight any more
    assert np.array_equal(sx, ref[1]) and np.array_equal(np.nan_to_num(ox), ref[0])
    bs = slamklt.StreamBatch(ctx, H, W, L, NF, NP)
    bs.prime(f64[NF])
    oy_ref, sy_ref = bs.step(pin_rev.array, pts[::-1], alg)
    assert np.array_equal(sy, sy_ref) and np.array_equal(np.nan_to_num(oy), np.nan_to_num(oy_ref))
    # and again on the rotated batches, the other way round
    by.step_begin(pin.array, pts, alg)
    bx.step_begin(pin_rev.array, pts[::-1], alg)
    ox2, sx2 = bx.step_end()
    oy2, sy2 = by.step_end()
    assert np.array_equal(sx2, ref[4]) and np.array_equal(np.nan_to_num(ox2), ref[3])
    # detect on the frames of a step that travelled in two formats equals detect on the same frames uploaded plainly
    ext = slamklt.Extractor(600, 17, (11, 36), 35)
    cur = np.stack([synth.random_keypoints(900 + i, 200, H, W, border=0.0) for i in range(NF)])
    got = by.detect(ext, cur)                              # by's last step: pin.array through step_begin (two engines)
    bs.upload(pin.array, pts)
    want = bs.detect(ext, cur)
    assert all(np.array_equal(g, w_) for g, w_ in zip(got, want)) and sum(len(g) for g in got) > 1000
    for b in (bx, by, bs):
        b.close()
    pin.free(); pin_rev.free()
