"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Float64 CPU restatement of the SLAM.jl KLT front-end path (see slam_oracle.c header;
PARITY UNPINNED: the reference holds no tests or golden vectors and Julia is absent).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.

Conventions mirror the reference: images are (H, W) float64 (stored column-major, y contiguous,
like Julia's Matrix{Gray{Float64}}), points are (N, 2) float64 in 1-based (y, x) order.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

PLANES = {"layer": 0, "Iy": 1, "Ix": 2, "Iyy": 3, "Ixx": 4, "Iyx": 5, "Syy": 6, "Sxx": 7, "Syx": 8, "blur": 9}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "slam_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp, ip, u8p, i64p = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_uint8), C.POINTER(C.c_int64)
        L.orc_iir2d.argtypes = [dp, dp, C.c_int, C.c_int, C.c_double, C.c_int]
        L.orc_iir1d.argtypes = [dp, C.c_int, C.c_double, C.c_double, C.c_double]
        L.orc_resize.argtypes = [dp, C.c_int, C.c_int, dp, C.c_int, C.c_int]
        L.orc_scharr.argtypes = [dp, C.c_int, C.c_int, dp, dp, C.c_int]
        L.orc_integral.argtypes = [dp, dp, C.c_int, C.c_int]
        L.orc_pyr_create.restype = C.c_void_p
        L.orc_pyr_create.argtypes = [C.c_int, C.c_int, C.c_int]
        L.orc_pyr_destroy.argtypes = [C.c_void_p]
        L.orc_pyr_build.argtypes = [C.c_void_p, dp, C.c_double, C.c_int]
        L.orc_pyr_plane.restype = dp
        L.orc_pyr_plane.argtypes = [C.c_void_p, C.c_int, C.c_int, ip, ip]
        L.orc_pinv2x2.argtypes = [dp, dp, dp]
        L.orc_optflow.restype = C.c_int
        L.orc_optflow.argtypes = [C.c_void_p, C.c_void_p, dp, dp, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_double, C.c_double, u8p]
        L.orc_fb_track.restype = C.c_int
        L.orc_fb_track.argtypes = [C.c_void_p, C.c_void_p, dp, dp, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_double, C.c_double, C.c_double, dp, u8p]
        L.orc_shi_tomasi.argtypes = [dp, C.c_int, C.c_int, dp]
        L.orc_mask.argtypes = [C.c_int, C.c_int, dp, C.c_int, C.c_int, C.c_double, dp]
        L.orc_detect.restype = C.c_int
        L.orc_detect.argtypes = [dp, C.c_int, C.c_int, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_double, C.c_double, i64p, C.c_int]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f(img):
    return np.asfortranarray(img, dtype=np.float64)


def iir2d(img, sigma, border="replicate"):
    a = _f(img)
    out = np.empty_like(a, order="F")
    lib().orc_iir2d(_dp(a), _dp(out), a.shape[0], a.shape[1], float(sigma), 0 if border == "replicate" else 1)
    return out


def iir1d(x, sigma, iminus=None, iplus=None):
    x = np.array(x, dtype=np.float64, copy=True)
    lib().orc_iir1d(_dp(x), len(x), float(sigma), float(x[0] if iminus is None else iminus),
                    float(x[-1] if iplus is None else iplus))
    return x


def resize(img, Ho, Wo):
    a = _f(img)
    out = np.empty((Ho, Wo), dtype=np.float64, order="F")
    lib().orc_resize(_dp(a), a.shape[0], a.shape[1], _dp(out), Ho, Wo)
    return out


def scharr(img, border="replicate"):
    a = _f(img)
    Iy = np.empty_like(a, order="F")
    Ix = np.empty_like(a, order="F")
    lib().orc_scharr(_dp(a), a.shape[0], a.shape[1], _dp(Iy), _dp(Ix), 0 if border == "replicate" else 1)
    return Iy, Ix


def integral(img):
    a = _f(img)
    out = np.empty_like(a, order="F")
    lib().orc_integral(_dp(a), _dp(out), a.shape[0], a.shape[1])
    return out


def pinv2x2(m):
    m = np.ascontiguousarray(m, dtype=np.float64)
    out = np.empty((2, 2))
    sv = np.empty(2)
    lib().orc_pinv2x2(_dp(m), _dp(out), _dp(sv))
    return out, sv


def shi_tomasi(img):
    a = _f(img)
    out = np.empty_like(a, order="F")
    lib().orc_shi_tomasi(_dp(a), a.shape[0], a.shape[1], _dp(out))
    return out


def mask(H, W, pts, radius, sigma=3.0):
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 2)
    out = np.empty((H, W), dtype=np.float64, order="F")
    lib().orc_mask(H, W, _dp(pts), len(pts), int(radius), float(sigma), _dp(out))
    return out


class LucasKanade:
    """lucas_kanade.jl:1-7"""

    def __init__(self, iterations=30, window_size=9, pyramid_levels=3, eigenvalue_threshold=1e-4, eps=1e-2):
        self.iterations, self.window_size, self.pyramid_levels = iterations, window_size, pyramid_levels
        self.eigenvalue_threshold, self.eps = eigenvalue_threshold, eps


class LKPyramid:
    """pyramid.jl:16-96.  mode 'ctor' = LKPyramid(image, levels; sigma, reusable=true); 'update' = update!(pyr, image)."""

    def __init__(self, image, levels, sigma=1.0, mode="ctor"):
        image = _f(image)
        self.H, self.W = image.shape
        self.levels = levels
        self._h = lib().orc_pyr_create(self.H, self.W, levels)
        self.update(image, sigma=sigma, mode=mode)

    def update(self, image, sigma=1.0, mode="update"):
        image = _f(image)
        assert image.shape == (self.H, self.W)
        lib().orc_pyr_build(self._h, _dp(image), float(sigma), 1 if mode == "ctor" else 0)
        return self

    def plane(self, level, name):
        h, w = C.c_int(), C.c_int()
        p = lib().orc_pyr_plane(self._h, level, PLANES[name], C.byref(h), C.byref(w))
        if not p:
            raise IndexError(level)
        arr = np.ctypeslib.as_array(p, shape=(w.value, h.value))  # column-major (H,W) == C-order (W,H)
        return arr.T.copy(order="F")

    def __del__(self):
        try:
            if self._h:
                lib().orc_pyr_destroy(self._h)
                self._h = None
        except Exception:
            pass


def optflow(displacement, first, second, points, alg: LucasKanade):
    """optflow! (lucas_kanade.jl:9-100).  Returns (displacement, status, n_good); displacement is a new array."""
    pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 2)
    disp = np.array(displacement, dtype=np.float64, copy=True).reshape(-1, 2)
    st = np.zeros(len(pts), dtype=np.uint8)
    n = lib().orc_optflow(first._h, second._h, _dp(pts), _dp(disp), len(pts), alg.iterations, alg.window_size,
                          alg.pyramid_levels, alg.eigenvalue_threshold, alg.eps,
                          st.ctypes.data_as(C.POINTER(C.c_uint8)))
    if n < 0:
        raise RuntimeError("Not enough layers in pyramids.")
    return disp, st.astype(bool), n


def fb_tracking(prev, cur, keypoints, displacement=None, iterations=30, window_size=11, pyramid_levels=3,
                max_distance=0.5, eigenvalue_threshold=1e-4, eps=1e-2):
    """fb_tracking! (tracker.jl:17-82).  Returns (new_keypoints, status, forward_status)."""
    pts = np.ascontiguousarray(keypoints, dtype=np.float64).reshape(-1, 2)
    n = len(pts)
    if n == 0:
        return None
    out = np.full((n, 2), np.nan)
    st = np.zeros(n, dtype=np.uint8)
    d = None if displacement is None else np.ascontiguousarray(displacement, dtype=np.float64).reshape(-1, 2)
    r = lib().orc_fb_track(prev._h, cur._h, _dp(pts), None if d is None else _dp(d), n, iterations, window_size,
                           pyramid_levels, eigenvalue_threshold, eps, float(max_distance), _dp(out),
                           st.ctypes.data_as(C.POINTER(C.c_uint8)))
    if r < 0:
        raise RuntimeError("Not enough layers in pyramids.")
    return out, (st & 1).astype(bool), ((st >> 1) & 1).astype(bool)


class Extractor:
    """extractor.jl:7-22 (descriptor omitted: describe is out of scope)."""

    def __init__(self, max_points, radius, grid_resolution, cell_size):
        self.max_points, self.radius = int(max_points), int(radius)
        self.grid_resolution, self.cell_size = (int(grid_resolution[0]), int(grid_resolution[1])), int(cell_size)


def detect(e: Extractor, image, current_points, sigma_mask=3.0, min_response=1e-4):
    """detect (extractor.jl:63-95).  Returns (n, 2) int64 array of 1-based (y, x)."""
    a = _f(image)
    cur = np.ascontiguousarray(current_points, dtype=np.float64).reshape(-1, 2)
    cap = a.shape[0] * a.shape[1]
    out = np.empty((cap, 2), dtype=np.int64)
    n = lib().orc_detect(_dp(a), a.shape[0], a.shape[1], _dp(cur), len(cur), e.max_points, e.radius,
                         e.grid_resolution[0], e.grid_resolution[1], e.cell_size, float(sigma_mask),
                         float(min_response), out.ctypes.data_as(C.POINTER(C.c_int64)), cap)
    return out[:n].copy()


class Camera:
    """camera.jl:1-46: intrinsics, distortion, image size and Ti0 (4 x 4, this camera <- camera 0)."""

    def __init__(self, fx, fy, cx, cy, k1=0.0, k2=0.0, p1=0.0, p2=0.0, height=0, width=0, Ti0=None):
        self.fx, self.fy, self.cx, self.cy = float(fx), float(fy), float(cx), float(cy)
        self.k1, self.k2, self.p1, self.p2 = float(k1), float(k2), float(p1), float(p2)
        self.height, self.width = int(height), int(width)
        self.Ti0 = np.eye(4) if Ti0 is None else np.array(Ti0, dtype=np.float64).reshape(4, 4)


def undistort_pdn_point(c: Camera, py, px):
    """undistort_pdn_point (camera.jl:106-128) on arrays of pre-divided (y, x); evaluation order of the reference."""
    sy, sx = py * py, px * px
    r2 = sy + sx
    rd = (1.0 + c.k1 * r2) + c.k2 * (r2 * r2)
    p = py * px
    dtx = (2 * c.p1) * p + c.p2 * (r2 + 2 * sy)
    dty = c.p1 * (r2 + 2 * sx) + (2 * c.p2) * p
    return (rd * py + dty) * c.fy + c.cy, (rd * px + dtx) * c.fx + c.cx


def undistort_point(c: Camera, pts_yx):
    """undistort_point (camera.jl:97-104)."""
    pts_yx = np.asarray(pts_yx, dtype=np.float64).reshape(-1, 2)
    y, x = undistort_pdn_point(c, (pts_yx[:, 0] - c.cy) / c.fy, (pts_yx[:, 1] - c.cx) / c.fx)
    return np.stack([y, x], axis=1)


def backproject(c: Camera, pts_yx):
    """backproject (camera.jl:130-143): (x, y, 1)."""
    pts_yx = np.asarray(pts_yx, dtype=np.float64).reshape(-1, 2)
    return np.stack([(pts_yx[:, 1] - c.cx) / c.fx, (pts_yx[:, 0] - c.cy) / c.fy, np.ones(len(pts_yx))], axis=1)


def _matmul4(A, B):
    """SMatrix * SMatrix: every element a left-to-right sum of products (no BLAS, no reassociation)."""
    out = np.empty((4, 4))
    for i in range(4):
        for j in range(4):
            acc = A[i, 0] * B[0, j]
            for k in range(1, 4):
                acc = acc + A[i, k] * B[k, j]
            out[i, j] = acc
    return out


def project_world_distort(c: Camera, T, world_xyz):
    """project_world_to_image_distort / ..._right_image_distort (frame.jl:458-484): T * [X;1], then project_undistort
    (camera.jl:73-85).  T is frame.cw or _matmul4(right_camera.Ti0, frame.cw)."""
    w = np.asarray(world_xyz, dtype=np.float64).reshape(-1, 3)
    X, Y, Z = w[:, 0], w[:, 1], w[:, 2]
    xc = ((T[0, 0] * X + T[0, 1] * Y) + T[0, 2] * Z) + T[0, 3]
    yc = ((T[1, 0] * X + T[1, 1] * Y) + T[1, 2] * Z) + T[1, 3]
    zc = ((T[2, 0] * X + T[2, 1] * Y) + T[2, 2] * Z) + T[2, 3]
    with np.errstate(all="ignore"):
        y, x = undistort_pdn_point(c, yc / zc, xc / zc)
    return np.stack([y, x], axis=1)


def in_image(c: Camera, pts_yx):
    """in_image (camera.jl:87-95)."""
    return (1 <= pts_yx[:, 0]) & (pts_yx[:, 0] <= c.height) & (1 <= pts_yx[:, 1]) & (pts_yx[:, 1] <= c.width)


def optical_flow_matching(from_pyr, to_pyr, pixels, is_3d, world, undist, cw, cam: Camera, right_cam: Camera = None,
                          stereo=False, window_size=9, pyramid_levels=3, max_distance=1.0, pyramid_levels_3d=1,
                          epipolar_error=2.0):
    """optical_flow_matching! (map_manager.jl:451-564) on arrays instead of the keypoint / map-point dictionaries, with
    update_keypoint! (frame.jl:252-270) or maybe_stereo_update! + update_stereo_keypoint! (map_manager.jl:579-590,
    frame.jl:272-287) applied to the tracked keypoints.  Returns (pixel, undistorted, position, status) with
    status bit0 updated, bit2 tracked by the prior pass, bit3 projection outside the image, bit4 epipolar reject."""
    pixels = np.asarray(pixels, dtype=np.float64).reshape(-1, 2)
    n = len(pixels)
    is_3d = np.asarray(is_3d).astype(bool)
    cw = np.asarray(cw, dtype=np.float64).reshape(4, 4)
    out_pix = np.full((n, 2), np.nan); out_und = np.full((n, 2), np.nan); out_pos = np.full((n, 3), np.nan)
    status = np.zeros(n, dtype=np.uint8)
    scale = 1.0 / 2.0 ** pyramid_levels_3d
    T = _matmul4(right_cam.Ti0, cw) if stereo else cw
    proj = project_world_distort(cam, T, world)
    inside = in_image(right_cam if stereo else cam, proj)
    i3 = np.flatnonzero(is_3d & inside)
    status[is_3d & ~inside] = 8
    i2 = np.flatnonzero(~is_3d)

    def update(idx, new_px, prior_pass):
        if len(idx) == 0:
            return
        if not stereo:
            und = undistort_point(cam, new_px)
            out_pix[idx], out_und[idx], out_pos[idx] = new_px, und, backproject(cam, und)
            status[idx] |= 1 | (4 if prior_pass else 0)
            return
        rp = undistort_point(right_cam, new_px)
        good = ~(np.abs(undist[idx, 0] - rp[:, 0]) > epipolar_error)
        status[idx[~good]] |= 16 | (4 if prior_pass else 0)
        g = idx[good]
        corrected = np.stack([pixels[g, 0], new_px[good, 1]], axis=1)
        und = undistort_point(right_cam, corrected)
        out_pix[g], out_und[g], out_pos[g] = corrected, und, backproject(right_cam, und)
        status[g] |= 1 | (4 if prior_pass else 0)

    if len(i3):
        disp = scale * (proj[i3] - pixels[i3])
        p3, s3, _ = fb_tracking(from_pyr, to_pyr, pixels[i3], displacement=disp, window_size=window_size,
                                pyramid_levels=pyramid_levels_3d, max_distance=max_distance)
        update(i3[s3], p3[s3], True)
        i2 = np.concatenate([i2, i3[~s3]])
    if len(i2):
        p2, s2, _ = fb_tracking(from_pyr, to_pyr, pixels[i2], window_size=window_size, pyramid_levels=pyramid_levels,
                                max_distance=max_distance)
        update(i2[s2], p2[s2], False)
    return out_pix, out_und, out_pos, status


def triangulate_stereo(und_yx, rund_yx, cam: Camera, rcam: Camera, wc, max_error=3.0):
    """triangulate_stereo! (mapper.jl:142-183) restated with numpy; the 4 x 4 eigen-problem goes to LAPACK (np.linalg.eigh) like
    the reference's geev.  [3P RecoverPose.triangulate]: rows x * P[3,:] - P[1,:], y * P[3,:] - P[2,:] per view (pixel units,
    points as (x, y), mapper.jl:162-164), homogeneous point = eigenvector of A'A for the smallest eigenvalue.
    Returns (world (n, 3) -- NaN unless status 1 --, status (n,) uint8: 1 update_mappoint!, 2 / 3 depth < 0.1 left / right,
    4 / 5 reprojection error left / right > max_error -> remove_stereo_keypoint!)."""
    und = np.asarray(und_yx, dtype=np.float64).reshape(-1, 2)
    rund = np.asarray(rund_yx, dtype=np.float64).reshape(-1, 2)
    n = len(und)
    K1 = np.array([[cam.fx, 0, cam.cx, 0], [0, cam.fy, cam.cy, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]])
    K2 = np.array([[rcam.fx, 0, rcam.cx, 0], [0, rcam.fy, rcam.cy, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]])
    Ti0 = np.asarray(rcam.Ti0, dtype=np.float64).reshape(4, 4)
    P1, P2 = K1, K2 @ Ti0                                       # mapper.jl:151-152
    wc = np.asarray(wc, dtype=np.float64).reshape(4, 4)
    world = np.full((n, 3), np.nan)
    status = np.zeros(n, dtype=np.uint8)
    for i in range(n):
        y1, x1 = und[i]; y2, x2 = rund[i]
        A = np.stack([x1 * P1[2] - P1[0], y1 * P1[2] - P1[1], x2 * P2[2] - P2[0], y2 * P2[2] - P2[1]])
        w, V = np.linalg.eigh(A.T @ A)
        X = V[:, int(np.argmin(w))]
        X = X * (1.0 / X[3])                                    # mapper.jl:165
        if not X[2] >= 0.1: status[i] = 2; continue
        R = Ti0 @ X
        if not R[2] >= 0.1: status[i] = 3; continue
        lp = np.array([cam.fy * X[1] / X[2] + cam.cy, cam.fx * X[0] / X[2] + cam.cx])      # project, camera.jl:62-67
        if np.linalg.norm(und[i] - lp) > max_error: status[i] = 4; continue
        rp = np.array([rcam.fy * R[1] / R[2] + rcam.cy, rcam.fx * R[0] / R[2] + rcam.cx])
        if np.linalg.norm(rund[i] - rp) > max_error: status[i] = 5; continue
        world[i] = (wc @ X)[:3]                                  # project_camera_to_world, frame.jl:452-456
        status[i] = 1
    return world, status


def describe(image, keypoints, pairs, window=9, sigma=2 ** 0.5):
    """describe (extractor.jl:103-105) -> ImageFeatures.create_descriptor(img, keypoints, BRIEF) [3P brief.jl], restated with numpy /
    scipy: imfilter(img, Kernel.gaussian(sigma)) = separable correlation with the normalised 4*ceil(sigma)+1 taps, replicated
    border; keypoints closer than ceil(window / 2) to the border are dropped; bit b = smoothed[k + s1[b]] < smoothed[k + s2[b]].
    Returns (descriptors (m, n_bits / 32) uint32 packed LSB first, kept keypoints (m, 2))."""
    import math
    from scipy import ndimage
    img = np.asarray(image, dtype=np.float64)
    H, W = img.shape
    hw = 2 * int(math.ceil(sigma))
    taps = np.exp(-np.arange(-hw, hw + 1) ** 2 / (2 * sigma * sigma)); taps /= taps.sum()
    sm = ndimage.correlate1d(ndimage.correlate1d(img, taps, axis=0, mode="nearest"), taps, axis=1, mode="nearest")
    kps = np.asarray(keypoints, dtype=np.int64).reshape(-1, 2)
    pr = np.asarray(pairs, dtype=np.int64).reshape(-1, 4)
    lim = -(-window // 2)
    keep = (kps[:, 0] - lim >= 1) & (kps[:, 1] - lim >= 1) & (kps[:, 0] + lim <= H) & (kps[:, 1] + lim <= W)
    k = kps[keep]
    a = sm[k[:, 0, None] - 1 + pr[None, :, 0], k[:, 1, None] - 1 + pr[None, :, 1]]
    b = sm[k[:, 0, None] - 1 + pr[None, :, 2], k[:, 1, None] - 1 + pr[None, :, 3]]
    bits = (a < b).astype(np.uint32).reshape(len(k), -1, 32)
    desc = (bits << np.arange(32, dtype=np.uint32)[None, None, :]).sum(axis=2).astype(np.uint32)
    return desc, k


def hamming_bits(d1, d2):
    """ImageFeatures.hamming_distance * n_bits: number of differing bits of two packed descriptors."""
    return int(sum(bin(int(x) ^ int(y)).count("1") for x, y in zip(d1, d2)))


def find_best_match(descriptors, set_offsets, target_sets, cand_offsets, candidates, max_distance):
    """The descriptor side of find_best_match (mapper.jl:392-462) with mappoint_min_distance (map_point.jl:165-174), loops as in the
    reference: candidates in order, empty descriptor sets skipped, `<=` so that a later tie replaces the best."""
    nt = len(target_sets)
    best_pos, best_dist, second_dist = np.full(nt, -1, np.int32), np.zeros(nt, np.int32), np.zeros(nt, np.int32)
    for t in range(nt):
        best = second = int(max_distance)
        pos = -1
        s1 = target_sets[t]
        for c in range(cand_offsets[t], cand_offsets[t + 1]):
            s2 = candidates[c]
            if set_offsets[s2 + 1] == set_offsets[s2]:
                continue
            d = 2 ** 31 - 1
            for i in range(set_offsets[s1], set_offsets[s1 + 1]):
                for j in range(set_offsets[s2], set_offsets[s2 + 1]):
                    d = min(d, hamming_bits(descriptors[i], descriptors[j]))
            if d <= best:
                second, best, pos = best, d, c - cand_offsets[t]
            elif d <= second:
                second = d
        best_pos[t], best_dist[t], second_dist[t] = pos, best, second
    return best_pos, best_dist, second_dist


def num_threads():
    return lib().orc_num_threads()


def set_threads(n):
    lib().orc_set_threads(int(n))
