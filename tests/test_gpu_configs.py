"""GPU parity at the other BASELINE.json configurations: config 3 (stereo left->right + temporal, 3000 keypoints,
4-level pyramid), config 5 (1920x1080, 8000 keypoints, 5-level pyramid, re-extraction), plus window sizes that take the
other kernel instantiations (w = 11 wrapper default, w = 15 maximum) and ragged / tiny images."""
import numpy as np
import pytest

import slamklt
from slamklt import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def check_tracks(ro, rg, min_n):
    (po, so, fo), (pg, sg, fg) = ro, rg
    assert np.mean(so == sg) >= 0.999 or np.sum(so != sg) <= 1, np.mean(so == sg)
    assert np.mean(fo == fg) >= 0.999 or np.sum(fo != fg) <= 1
    both = so & sg
    assert both.sum() >= min_n
    d = np.abs(po[both] - pg[both]).max(axis=1)
    # an epsilon-stop flip (|step| within rounding of 1e-2) moves a point by at most ~0.0141 px; allow one per test
    assert np.mean(d < 0.01) >= 0.999 or np.sum(d >= 0.01) <= 1, (np.mean(d < 0.01), d.max())
    assert d.max() < 0.02


def test_config3_stereo_and_temporal_L4(ctx):
    left, right, disp = synth.stereo_pair(3000)
    fl, fr = synth.to_f64(left[None])[0], synth.to_f64(right[None])[0]
    e = O.Extractor(3000, 17, (11, 36), 35)
    pts = O.detect(e, fl, np.zeros((0, 2))).astype(np.float64)[:3000]
    assert len(pts) >= 2500
    L = 4
    ol, orr = O.LKPyramid(fl, L), O.LKPyramid(fr, L)
    gl, gr = slamklt.LKPyramid(ctx, fl, L), slamklt.LKPyramid(ctx, fr, L)
    for l in range(L + 1):
        assert rel_err(gr.plane(l, "layer"), orr.plane(l, "layer")) < 1e-5
        assert rel_err(gr.plane(l, "Sxx"), orr.plane(l, "Sxx")) < 1e-5
    # left -> right matching (mapper.jl:58-60): disparities up to 40 px need the 4-level pyramid
    ro = O.fb_tracking(ol, orr, pts, window_size=9, pyramid_levels=L, max_distance=1.0)
    rg = slamklt.fb_tracking(gl, gr, pts, window_size=9, pyramid_levels=L, max_distance=1.0)
    check_tracks(ro, rg, 1000)
    ok = rg[1]
    d_est = pts[ok, 1] - rg[0][ok, 1]
    d_true = disp[np.clip(np.rint(pts[ok, 0]).astype(int) - 1, 0, 375), np.clip(np.rint(pts[ok, 1]).astype(int) - 1, 0, 1240)]
    assert np.median(np.abs(d_est - d_true)) < 0.5
    # 3-D points with a prior, pyramid_levels = 1 (map_manager.jl:458-521)
    prior = np.zeros_like(pts); prior[:, 1] = -0.5 * d_true_full(disp, pts)
    ro = O.fb_tracking(ol, orr, pts, displacement=prior, window_size=9, pyramid_levels=1, max_distance=1.0)
    rg = slamklt.fb_tracking(gl, gr, pts, displacement=prior, window_size=9, pyramid_levels=1, max_distance=1.0)
    check_tracks(ro, rg, 1000)


def d_true_full(disp, pts):
    return disp[np.clip(np.rint(pts[:, 0]).astype(int) - 1, 0, disp.shape[0] - 1),
                np.clip(np.rint(pts[:, 1]).astype(int) - 1, 0, disp.shape[1] - 1)]


def test_config5_1080p_L5_8000kp(ctx):
    fr, aff = synth.make_sequence(5000, 2, H=1080, W=1920)
    f = synth.to_f64(fr)
    L = 5
    o0, o1 = O.LKPyramid(f[0], L), O.LKPyramid(f[1], L)
    o1.update(f[1])
    g0, g1 = slamklt.LKPyramid(ctx, f[0], L), slamklt.LKPyramid(ctx, f[1], L)
    g1.update(f[1])
    assert [g1.level_shape(l) for l in range(L + 1)] == [(1080, 1920), (540, 960), (270, 480), (135, 240), (68, 120), (34, 60)]
    for l in range(L + 1):
        for name in ("layer", "Iy", "Ix", "Syy", "Syx"):
            assert rel_err(g1.plane(l, name), o1.plane(l, name)) < 1e-5, (l, name)
    # re-extraction every frame: 8000 keypoints over a 31 x 55 grid of 35-px cells
    eo, eg = O.Extractor(8000, 17, (31, 55), 35), slamklt.Extractor(8000, 17, (31, 55), 35)
    kp_o = O.detect(eo, f[0], np.zeros((0, 2)))
    kp_g = slamklt.detect(ctx, eg, f[0], np.zeros((0, 2)))
    assert np.array_equal(kp_o, kp_g) and len(kp_g) > 6000
    pts = kp_g.astype(np.float64)[:8000] + 0.3
    ro = O.fb_tracking(o0, o1, pts, window_size=9, pyramid_levels=L, max_distance=1.0)
    rg = slamklt.fb_tracking(g0, g1, pts, window_size=9, pyramid_levels=L, max_distance=1.0)
    check_tracks(ro, rg, 4000)
    cur = rg[0][rg[1]]
    assert np.array_equal(O.detect(eo, f[1], cur), slamklt.detect(ctx, eg, f[1], cur))


@pytest.mark.parametrize("window", [4, 11, 15])
def test_other_window_sizes(ctx, window):
    fr, _ = synth.make_sequence(77, 2, H=200, W=300)
    f = synth.to_f64(fr)
    pts = synth.random_keypoints(9, 400, 200, 300, border=1.0)
    o0, o1 = O.LKPyramid(f[0], 2), O.LKPyramid(f[1], 2)
    g0, g1 = slamklt.LKPyramid(ctx, f[0], 2), slamklt.LKPyramid(ctx, f[1], 2)
    ro = O.fb_tracking(o0, o1, pts, window_size=window, pyramid_levels=2, max_distance=0.5)
    rg = slamklt.fb_tracking(g0, g1, pts, window_size=window, pyramid_levels=2, max_distance=0.5)
    check_tracks(ro, rg, 100)


@pytest.mark.parametrize("shape", [(33, 47), (8, 300), (300, 9), (65, 1290)])
def test_ragged_and_tiny_images(ctx, shape):
    H, W = shape
    img = np.random.default_rng(H * W).uniform(0, 1, shape)
    levels = 1
    o = O.LKPyramid(img, levels); o.update(img)
    g = slamklt.LKPyramid(ctx, img, levels); g.update(img)
    for l in range(levels + 1):
        for name in ("layer", "Iy", "Ix", "Sxx", "Syy", "Syx"):
            assert rel_err(g.plane(l, name), o.plane(l, name)) < 1e-5, (l, name)
    pts = np.array([[1.0, 1.0], [H, W], [H / 2, W / 2], [1.5, W - 0.5], [H - 0.2, 1.3]], dtype=np.float64)
    ro = O.fb_tracking(o, o, pts, window_size=9, pyramid_levels=levels, max_distance=1.0)
    rg = slamklt.fb_tracking(g, g, pts, window_size=9, pyramid_levels=levels, max_distance=1.0)
    assert np.array_equal(ro[1], rg[1]) and np.array_equal(ro[2], rg[2])
    both = ro[1] & rg[1]
    assert np.all(np.abs(ro[0][both] - rg[0][both]) < 0.01)


def test_invalid_arguments(ctx):
    with pytest.raises(slamklt.SlamKltError) as ei:
        slamklt.LKPyramid(ctx, np.zeros((3, 50)), 1)
    assert ei.value.code == slamklt.E_INVALID          # recursive filter needs more than 3 samples per line
    with pytest.raises(slamklt.SlamKltError):
        slamklt.LKPyramid(ctx, np.zeros((40, 40)), 5)  # level 4 would be 3 x 3
    g = slamklt.LKPyramid(ctx, np.random.default_rng(0).uniform(size=(64, 64)), 1)
    with pytest.raises(slamklt.SlamKltError):
        slamklt.fb_tracking(g, g, np.array([[5.0, 5.0]]), window_size=16)
    with pytest.raises(slamklt.SlamKltError):
        g.update(np.zeros((64, 65)))


def test_optical_flow_matching_two_pass(ctx):
    """the tracking part of optical_flow_matching! (map_manager.jl:451-564) in one launch equals the
    reference's two fb_tracking! calls (3-D keypoints with a prior on 1 level, then failures + 2-D keypoints on 3 levels)."""
    fr, aff = synth.make_sequence(4100, 2)
    f = synth.to_f64(fr)
    rng = np.random.default_rng(8)
    pts = synth.random_keypoints(12, 1500, 376, 1241, border=2.0)
    gt = synth.true_flow(aff, 0, 1, pts)
    is3d = rng.random(len(pts)) < 0.5
    prior = 0.5 * (gt - pts) + rng.normal(0, 0.4, pts.shape)
    bad = is3d & (rng.random(len(pts)) < 0.3)          # wrong map points: prior far off => prior pass fails, retried as 2-D
    prior[bad] += rng.choice([-6.0, 6.0], size=(bad.sum(), 2))
    o0, o1 = O.LKPyramid(f[0], 3), O.LKPyramid(f[1], 3)
    o0.update(f[0]); o1.update(f[1])
    g0, g1 = slamklt.LKPyramid(ctx, f[0], 3), slamklt.LKPyramid(ctx, f[1], 3)
    g0.update(f[0]); g1.update(f[1])
    # reference composition
    exp_pts = np.full_like(pts, np.nan); exp_st = np.zeros(len(pts), bool); exp_3d = np.zeros(len(pts), bool)
    i3 = np.flatnonzero(is3d)
    p3, s3, _ = O.fb_tracking(o0, o1, pts[i3], displacement=prior[i3], window_size=9, pyramid_levels=1, max_distance=1.0)
    exp_pts[i3[s3]] = p3[s3]; exp_st[i3[s3]] = True; exp_3d[i3[s3]] = True
    i2 = np.concatenate([np.flatnonzero(~is3d), i3[~s3]])
    p2, s2, _ = O.fb_tracking(o0, o1, pts[i2], window_size=9, pyramid_levels=3, max_distance=1.0)
    exp_pts[i2[s2]] = p2[s2]; exp_st[i2[s2]] = True
    got_pts, got_st, got_3d = slamklt.optical_flow_matching(g0, g1, pts, prior, is3d, window_size=9, pyramid_levels=3,
                                                            pyramid_levels_3d=1, max_distance=1.0)
    assert np.mean(got_st == exp_st) >= 0.999 and np.mean(got_3d == exp_3d) >= 0.999
    both = got_st & exp_st & (got_3d == exp_3d)
    assert both.sum() > 1000 and exp_3d.sum() > 200 and (i3[~s3]).size > 50
    d = np.abs(got_pts[both] - exp_pts[both]).max(axis=1)
    assert np.mean(d < 0.01) >= 0.999 and d.max() < 0.02
