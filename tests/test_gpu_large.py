"""GPU parity tests of frames beyond the register-tiled pyramid kernels (more than 1088 rows or 2048 columns: portrait 1080p, 4K):
their levels are built by the general kernels of pyramid.cu (one thread per line, Float64 recursion), coarser levels that fit go
back to the tiled kernels.  Same oracle, same tolerances as tests/test_gpu_parity.py; SLAMKLT_FORCE_GENERIC=1 also sends a
KITTI-sized frame through the general kernels so that both builds of the same frame can be compared with each other."""
import numpy as np
import pytest

import slamklt
from slamklt import synth
from oracle import oracle as O
from test_gpu_parity import rel_err, elem_frac, PYR_RTOL, ELEM_FRAC, POS_TOL, FLAG_AGREE

pytestmark = pytest.mark.gpu

PLANES = ("layer", "Iy", "Ix", "Syy", "Sxx", "Syx")


def _check_planes(gp, op, levels):
    for l in range(levels + 1):
        for name in PLANES:
            a, b = gp.plane(l, name), op.plane(l, name)
            assert a.shape == b.shape
            assert rel_err(a, b) < PYR_RTOL, (l, name, rel_err(a, b))
            assert elem_frac(a, b) >= ELEM_FRAC, (l, name, elem_frac(a, b))
        if l < levels:
            assert rel_err(gp.plane(l, "blur"), op.plane(l, "blur")) < PYR_RTOL, l
        # the planes the tracking kernel reads: 19-column window row sums of the raw device prefix planes
        for rname, sname in (("Ryy", "Syy"), ("Rxx", "Sxx"), ("Ryx", "Syx")):
            R, S = gp.plane(l, rname), op.plane(l, sname)
            assert R.shape == (S.shape[0], S.shape[1] + 1) and np.all(R[:, 0] == 0.0)
            ref = np.concatenate([np.zeros((S.shape[0], 1)), np.cumsum(S, axis=1)], axis=1)
            k = min(19, S.shape[1])
            rd, rr = R[:, k:] - R[:, :-k], ref[:, k:] - ref[:, :-k]
            kr = min(19, S.shape[0])
            cd = np.concatenate([np.zeros((1, rd.shape[1])), np.cumsum(rd, axis=0)])
            cr = np.concatenate([np.zeros((1, rr.shape[1])), np.cumsum(rr, axis=0)])
            assert rel_err(cd[kr:] - cd[:-kr], cr[kr:] - cr[:-kr]) < PYR_RTOL, (l, rname)


@pytest.mark.parametrize("mode", ["ctor", "update"])
def test_general_kernels_on_a_kitti_frame(ctx, monkeypatch, mode):
    """Every level through the general kernels (forced): planes against the oracle, and against the tiled build of the same frame."""
    fr, _ = synth.make_sequence(2101, 2)
    f = synth.to_f64(fr)
    op = O.LKPyramid(f[0], 3, mode="ctor")
    tiled = slamklt.LKPyramid(ctx, f[0], 3)
    monkeypatch.setenv("SLAMKLT_FORCE_GENERIC", "1")
    gen = slamklt.LKPyramid(ctx, f[0], 3)
    if mode == "update":
        gen.update(f[1])
    monkeypatch.delenv("SLAMKLT_FORCE_GENERIC")
    if mode == "update":
        op.update(f[1]); tiled.update(f[1])
    _check_planes(gen, op, 3)
    for l in range(4):
        for name in PLANES:
            a, b = gen.plane(l, name), tiled.plane(l, name)
            assert rel_err(a, b) < 2e-6, (l, name, rel_err(a, b))   # two fp32 evaluations of the same Float64 quantity


@pytest.mark.parametrize("shape,levels,dtype", [((1200, 300), 3, "f64"), ((120, 2100), 3, "f64"), ((1920, 1080), 4, "u8"),
                                                ((2160, 3840), 5, "f64")])
def test_large_frames(ctx, shape, levels, dtype):
    """Tall, wide, portrait 1080p (UInt8 frames) and 4K: pyramids (both border regimes), tracking and extraction against the oracle."""
    H, W = shape
    fr, aff = synth.make_sequence(2200 + H, 2, H=H, W=W)
    f = synth.to_f64(fr)
    src = fr if dtype == "u8" else f
    o0 = O.LKPyramid(f[0], levels, mode="ctor")
    g0 = slamklt.LKPyramid(ctx, src[0], levels)
    _check_planes(g0, o0, levels)
    o1 = O.LKPyramid(f[1], levels, mode="ctor"); o1.update(f[1])
    g1 = slamklt.LKPyramid(ctx, src[1], levels); g1.update(src[1])
    _check_planes(g1, o1, levels)
    o0.update(f[0]); g0.update(src[0])
    # forward-backward tracking over all levels
    n = 3000
    pts = synth.random_keypoints(77, n, H, W, border=3.0)
    ko, so, fo = O.fb_tracking(o0, o1, pts, window_size=9, pyramid_levels=levels, max_distance=1.0)
    kg, sg, fg = slamklt.fb_tracking(g0, g1, pts, window_size=9, pyramid_levels=levels, max_distance=1.0)
    assert np.mean(so == sg) >= FLAG_AGREE and np.mean(fo == fg) >= FLAG_AGREE, (np.mean(so == sg), np.mean(fo == fg))
    both = so & sg
    assert both.sum() > 0.5 * n
    d = np.abs(ko[both] - kg[both]).max(axis=1)
    assert np.mean(d < POS_TOL) >= FLAG_AGREE and d.max() < 0.03, (np.mean(d < POS_TOL), d.max())
    # extraction on the raw frame with an avoidance mask
    cells = (max(1, H // 35), max(1, W // 35))
    e_args = (cells[0] * cells[1] * 4, 17, cells, 35)
    cur = synth.random_keypoints(78, 500, H, W, border=0.0)
    ko = O.detect(O.Extractor(*e_args), f[0], cur)
    kg = slamklt.detect(ctx, slamklt.Extractor(*e_args), src[0], cur)
    assert ko.shape == kg.shape and np.array_equal(ko, kg)


@pytest.mark.parametrize("dtype", ["f64", "u8"])
def test_batch_step_on_tall_frames(ctx, dtype):
    """The batched stream API on frames whose level 0 goes through the general kernels: one step of three 1200 x 260 frames gives
    exactly what single pyramids and fb_tracking! calls give, and its extraction equals the single-frame call."""
    H, W, L, NF, NP = 1200, 260, 3, 3, 400
    fr, _ = synth.make_sequence(2300, NF + 1, H=H, W=W)
    f64 = synth.to_f64(fr)
    src = f64 if dtype == "f64" else fr
    alg = slamklt.LucasKanade(pyramid_levels=L, window_size=9)
    batch = slamklt.StreamBatch(ctx, H, W, L, NF, NP)
    batch.prime(src[0], mode=slamklt.MODE_UPDATE)
    pts = np.stack([synth.random_keypoints(90 + i, NP, H, W, border=2.0) for i in range(NF)])
    out, st = batch.step(slamklt.StreamBatch.pack_frames(src[1:]), pts, alg, max_distance=1.0)
    prev = slamklt.LKPyramid(ctx, f64[0], L); prev.update(f64[0])
    for i in range(NF):
        cur = slamklt.LKPyramid(ctx, f64[i + 1], L); cur.update(f64[i + 1])
        ref_pts, ref_st, ref_fst = slamklt.fb_tracking(prev, cur, pts[i], window_size=9, pyramid_levels=L, max_distance=1.0)
        assert np.array_equal(st[i] & 1, ref_st.astype(np.uint8)) and np.array_equal((st[i] >> 1) & 1, ref_fst.astype(np.uint8))
        assert np.array_equal(out[i][ref_fst], ref_pts[ref_fst])
        assert ref_st.mean() > 0.5
        prev = cur
    e = slamklt.Extractor(1000, 17, (H // 35, W // 35), 35)
    batch.upload(slamklt.StreamBatch.pack_frames(src[1:]), pts)
    kps = batch.detect(e, pts)
    for i in range(NF):
        assert np.array_equal(kps[i], slamklt.detect(ctx, e, f64[i + 1], pts[i]))
    batch.close()


@pytest.mark.parametrize("shape,dtype,mode", [((600, 200), "f64", "update"), ((768, 130), "u8", "ctor"), ((1080, 320), "u8", "update"),
                                              ((1080, 320), "f64", "ctor"), ((1087, 150), "f64", "update"), ((771, 90), "f32", "ctor"),
                                              ((1088, 64), "u8", "update")])
def test_two_warps_per_column(ctx, monkeypatch, shape, dtype, mode):
    """Frames taller than 512 rows: the fused column kernel runs with two warps per column (12 or 18 rows per lane, states handed
    across the pair).  Planes against the oracle, and against the one-warp kernels (SLAMKLT_COLS_PAIR=0) on the same frame."""
    H, W = shape
    fr, _ = synth.make_sequence(2400 + H + W, 2, H=H, W=W)
    f = synth.to_f64(fr)
    src = {"f64": f, "u8": fr, "f32": f.astype(np.float32)}[dtype]
    ref = f if dtype != "f32" else src.astype(np.float64)
    L = 2
    op = O.LKPyramid(ref[0], L, mode="ctor")
    pair = slamklt.LKPyramid(ctx, src[0], L)
    monkeypatch.setenv("SLAMKLT_COLS_PAIR", "0")
    single = slamklt.LKPyramid(ctx, src[0], L)
    if mode == "update":
        single.update(src[1])
    monkeypatch.delenv("SLAMKLT_COLS_PAIR")
    if mode == "update":
        op.update(ref[1]); pair.update(src[1])
    _check_planes(pair, op, L)
    for l in range(L + 1):
        for name in PLANES + (("blur",) if l < L else ()):
            a, b = pair.plane(l, name), single.plane(l, name)
            assert rel_err(a, b) < 2e-6, (l, name, rel_err(a, b))


@pytest.mark.parametrize("shape,dtype,mode", [((1280, 200), "u8", "update"), ((1100, 130), "f64", "ctor"), ((1200, 96), "f32", "update"),
                                              ((1089, 72), "f64", "update")])
def test_level0_up_to_1280_rows(ctx, monkeypatch, shape, dtype, mode):
    """Level 0 of frames with 1089 ... 1280 rows (portrait 720p) is built by the two-warps-per-column kernel (20 rows per lane),
    the coarser levels by the ordinary tiled kernels.  Planes against the oracle, and against the general per-line kernels
    (SLAMKLT_FORCE_GENERIC=1) on the same frame."""
    H, W = shape
    fr, _ = synth.make_sequence(3100 + H + W, 2, H=H, W=W)
    f = synth.to_f64(fr)
    src = {"f64": f, "u8": fr, "f32": f.astype(np.float32)}[dtype]
    ref = f if dtype != "f32" else src.astype(np.float64)
    L = 2
    op = O.LKPyramid(ref[0], L, mode="ctor")
    tiled = slamklt.LKPyramid(ctx, src[0], L)
    monkeypatch.setenv("SLAMKLT_FORCE_GENERIC", "1")
    general = slamklt.LKPyramid(ctx, src[0], L)
    if mode == "update":
        general.update(src[1])
    monkeypatch.delenv("SLAMKLT_FORCE_GENERIC")
    if mode == "update":
        op.update(ref[1]); tiled.update(src[1])
    _check_planes(tiled, op, L)
    for l in range(L + 1):
        for name in PLANES + (("blur",) if l < L else ()):
            a, b = tiled.plane(l, name), general.plane(l, name)
            assert rel_err(a, b) < 4e-6, (l, name, rel_err(a, b))


def test_batch_step_1080p_equals_single_calls(ctx):
    """Config 5 geometry through the batched stream API (two warps per column at levels 0 and 1): one step of three 1080p UInt8 frames
    equals single pyramids + fb_tracking! bit for bit, and the step's tracks agree with the oracle."""
    H, W, L, NF, NP = 1080, 1920, 5, 3, 1500
    fr, _ = synth.make_sequence(2500, NF + 1, H=H, W=W)
    f64 = synth.to_f64(fr)
    alg = slamklt.LucasKanade(pyramid_levels=L, window_size=9)
    batch = slamklt.StreamBatch(ctx, H, W, L, NF, NP)
    batch.prime(fr[0], mode=slamklt.MODE_UPDATE)
    pts = np.stack([synth.random_keypoints(130 + i, NP, H, W, border=2.0) for i in range(NF)])
    out, st = batch.step(slamklt.StreamBatch.pack_frames(fr[1:]), pts, alg, max_distance=1.0)
    prev = slamklt.LKPyramid(ctx, fr[0], L); prev.update(fr[0])
    for i in range(NF):
        cur = slamklt.LKPyramid(ctx, fr[i + 1], L); cur.update(fr[i + 1])
        ref_pts, ref_st, ref_fst = slamklt.fb_tracking(prev, cur, pts[i], window_size=9, pyramid_levels=L, max_distance=1.0)
        assert np.array_equal(st[i] & 1, ref_st.astype(np.uint8)) and np.array_equal((st[i] >> 1) & 1, ref_fst.astype(np.uint8))
        assert np.array_equal(out[i][ref_fst], ref_pts[ref_fst])
        prev = cur
    o0 = O.LKPyramid(f64[NF - 1], L); o0.update(f64[NF - 1]); o1 = O.LKPyramid(f64[NF], L); o1.update(f64[NF])
    po, so, fo = O.fb_tracking(o0, o1, pts[NF - 1], window_size=9, pyramid_levels=L, max_distance=1.0)
    sg = (st[NF - 1] & 1).astype(bool)
    assert np.mean(so == sg) >= FLAG_AGREE and so.mean() > 0.5
    both = so & sg
    assert np.mean(np.abs(po[both] - out[NF - 1][both]).max(axis=1) < POS_TOL) >= FLAG_AGREE
    batch.close()
