"""Seeded synthetic KITTI-shaped frames.  numpy + OpenCV only; no device code.

A "world" texture (band-limited noise + random-contrast rectangles and discs, so that every
35-px cell holds corners) is viewed through a slowly moving affine camera.  Frames are
quantised to u8 and returned both as u8 and as Float64 = u8/255, which is exactly what the
reference sees after `Gray{Float64}.(load(png))` (example/kitty/main.jl:36-40), so the f64 and
u8 entry points of the library receive identical pixel values.
"""
from __future__ import annotations

import cv2
import numpy as np

KITTI_H, KITTI_W = 376, 1241


def make_world(seed: int, H: int, W: int, margin: int = 72) -> np.ndarray:
    rng = np.random.default_rng(seed)
    Hh, Ww = H + 2 * margin, W + 2 * margin
    tex = rng.standard_normal((Hh, Ww)).astype(np.float32)
    tex = cv2.GaussianBlur(tex, (0, 0), 1.5)
    tex /= max(float(tex.std()), 1e-6)
    world = 0.5 + 0.08 * tex
    n_shapes = max(8, (Hh * Ww) // 1100)
    for _ in range(n_shapes):
        cy, cx = int(rng.integers(0, Hh)), int(rng.integers(0, Ww))
        val = float(rng.uniform(0.05, 0.95))
        if rng.random() < 0.6:
            hh, ww = int(rng.integers(4, 26)), int(rng.integers(4, 26))
            ang = float(rng.uniform(0, 90))
            box = cv2.boxPoints(((cx, cy), (ww, hh), ang)).astype(np.int32)
            cv2.fillConvexPoly(world, box, val, lineType=cv2.LINE_AA)
        else:
            cv2.circle(world, (cx, cy), int(rng.integers(3, 12)), val, -1, lineType=cv2.LINE_AA)
    world = cv2.GaussianBlur(world, (0, 0), 0.8)
    world += 0.03 * cv2.GaussianBlur(rng.standard_normal((Hh, Ww)).astype(np.float32), (0, 0), 1.0)
    return np.clip(world, 0.05, 0.95).astype(np.float32)


def camera_affine(rng_params, t: int, H: int, W: int, margin: int):
    """2x3 matrix mapping frame pixel (x, y) -> world pixel (x, y) at time t."""
    ay, ax, wy, wx, py, px, rot_a, rot_w, sc_a, sc_w = rng_params
    ty = ay * np.sin(wy * t + py)
    tx = ax * np.sin(wx * t + px)
    th = np.deg2rad(rot_a * np.sin(rot_w * t))
    s = 1.0 + sc_a * np.sin(sc_w * t + 1.0)
    c, sn = s * np.cos(th), s * np.sin(th)
    cx, cy = (W - 1) / 2.0, (H - 1) / 2.0
    # rotate/scale about the frame centre, then translate into the world (offset by margin)
    M = np.array([[c, -sn, cx - c * cx + sn * cy + margin + tx],
                  [sn, c, cy - sn * cx - c * cy + margin + ty]], dtype=np.float64)
    return M


def make_sequence(seed: int, n_frames: int, H: int = KITTI_H, W: int = KITTI_W, max_step: float = 10.0,
                  noise: float = 0.005, margin: int = 72):
    """Returns (frames_u8 [n,H,W] uint8, affines [n,2,3]).  Frame t samples the world through affines[t]."""
    world = make_world(seed, H, W, margin)
    rng = np.random.default_rng(seed + 7919)
    amp = margin - 24.0
    wy, wx = rng.uniform(0.12, max_step / amp, size=2)
    params = (amp, amp, wy, wx, rng.uniform(0, 6.28), rng.uniform(0, 6.28),
              rng.uniform(0.1, 0.5), rng.uniform(0.05, 0.2), rng.uniform(0.001, 0.006), rng.uniform(0.05, 0.2))
    frames = np.empty((n_frames, H, W), dtype=np.uint8)
    affs = np.empty((n_frames, 2, 3))
    for t in range(n_frames):
        M = camera_affine(params, t, H, W, margin)
        affs[t] = M
        f = cv2.warpAffine(world, M, (W, H), flags=cv2.INTER_CUBIC | cv2.WARP_INVERSE_MAP,
                           borderMode=cv2.BORDER_REFLECT)
        f = f + noise * np.random.default_rng(seed * 131 + t).standard_normal((H, W)).astype(np.float32)
        frames[t] = np.clip(np.rint(f * 255.0), 0, 255).astype(np.uint8)
    return frames, affs


def to_f64(frames_u8: np.ndarray) -> np.ndarray:
    """u8 -> Float64 exactly like Gray{Float64}(N0f8): i/255."""
    return frames_u8.astype(np.float64) / 255.0


def true_flow(affs: np.ndarray, t0: int, t1: int, pts_yx: np.ndarray) -> np.ndarray:
    """Ground-truth positions in frame t1 of 1-based (y, x) points given in frame t0."""
    A0 = np.vstack([affs[t0], [0, 0, 1]])
    A1 = np.vstack([affs[t1], [0, 0, 1]])
    T = np.linalg.inv(A1) @ A0  # frame t0 pixel -> world -> frame t1 pixel
    xy = np.stack([pts_yx[:, 1] - 1.0, pts_yx[:, 0] - 1.0, np.ones(len(pts_yx))], axis=0)
    out = T @ xy
    return np.stack([out[1] + 1.0, out[0] + 1.0], axis=1)


def stereo_pair(seed: int, H: int = KITTI_H, W: int = KITTI_W, disparity=(2.0, 40.0), noise: float = 0.005,
                margin: int = 72):
    """Left/right pair: right = left shifted along x by a smooth disparity field in [disparity]."""
    world = make_world(seed, H, W, margin)
    rng = np.random.default_rng(seed + 104729)
    left = world[margin:margin + H, margin:margin + W]
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float32)
    d = disparity[0] + (disparity[1] - disparity[0]) * (0.5 + 0.5 * np.sin(xs / W * 2.1 + rng.uniform(0, 6)) *
                                                        np.cos(ys / H * 1.3 + rng.uniform(0, 6)))
    right = cv2.remap(world, xs + margin + d.astype(np.float32), ys + margin, cv2.INTER_CUBIC,
                      borderMode=cv2.BORDER_REFLECT)
    out = []
    for i, f in enumerate((left, right)):
        f = f + noise * np.random.default_rng(seed * 17 + i).standard_normal((H, W)).astype(np.float32)
        out.append(np.clip(np.rint(f * 255.0), 0, 255).astype(np.uint8))
    return out[0], out[1], d


def right_views(frames_u8: np.ndarray, seed: int, disparity=(2.0, 40.0), noise: float = 0.005) -> np.ndarray:
    """Right-camera views of a left sequence: every frame resampled along x by a smooth disparity field in [disparity]
    (a scene point at left x appears at right x - d), plus sensor noise.  Same shape and dtype as the input."""
    n, H, W = frames_u8.shape
    rng = np.random.default_rng(seed + 15485863)
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float32)
    d = disparity[0] + (disparity[1] - disparity[0]) * (0.5 + 0.5 * np.sin(xs / W * 2.1 + rng.uniform(0, 6)) *
                                                        np.cos(ys / H * 1.3 + rng.uniform(0, 6)))
    out = np.empty_like(frames_u8)
    for t in range(n):
        r = cv2.remap(frames_u8[t].astype(np.float32) / 255.0, xs + d.astype(np.float32), ys, cv2.INTER_CUBIC, borderMode=cv2.BORDER_REFLECT)
        r = r + noise * np.random.default_rng(seed * 257 + t).standard_normal((H, W)).astype(np.float32)
        out[t] = np.clip(np.rint(r * 255.0), 0, 255).astype(np.uint8)
    return out


def random_keypoints(seed: int, n: int, H: int, W: int, border: float = 12.0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    y = rng.uniform(1 + border, H - border, size=n)
    x = rng.uniform(1 + border, W - border, size=n)
    return np.stack([y, x], axis=1)


KITTI_CAMERA = dict(fx=718.856, fy=718.856, cx=607.1928, cy=185.2157, k1=-0.08, k2=0.015, p1=8e-4, p2=-5e-4,
                    height=KITTI_H, width=KITTI_W)


def matching_scene(seed: int, pixels_yx: np.ndarray, targets_yx: np.ndarray, camera: dict = None, baseline: float = 0.0,
                   frac_3d: float = 0.5, frac_bad: float = 0.25, frac_outside: float = 0.05, prior_noise: float = 0.6):
    """Inputs of optical_flow_matching! for keypoints `pixels_yx` whose true positions in the target image are
    `targets_yx`: a frame pose cw, map-point positions of the 3-D keypoints that project (through the distorting camera,
    camera.jl:106-128) close to the targets -- some far off (wrong map points) and some outside the image -- and the
    right camera's Ti0 for a stereo rig with `baseline`.  Returns a dict of arrays."""
    cam = dict(KITTI_CAMERA if camera is None else camera)
    rng = np.random.default_rng(seed)
    n = len(pixels_yx)
    is_3d = rng.random(n) < frac_3d
    tgt = targets_yx + rng.normal(0.0, prior_noise, targets_yx.shape)
    bad = is_3d & (rng.random(n) < frac_bad)
    tgt[bad] += rng.choice([-14.0, 14.0], size=(int(bad.sum()), 2))
    outside = is_3d & ~bad & (rng.random(n) < frac_outside)
    tgt[outside, 1] += rng.choice([-2.0, 2.0], size=int(outside.sum())) * cam["width"]
    # invert the distortion model by fixed-point iteration: distorted normalised (yd, xd) -> undistorted (yn, xn)
    yd, xd = (tgt[:, 0] - cam["cy"]) / cam["fy"], (tgt[:, 1] - cam["cx"]) / cam["fx"]
    yn, xn = yd.copy(), xd.copy()
    for _ in range(20):
        r2 = yn * yn + xn * xn
        rd = 1.0 + cam["k1"] * r2 + cam["k2"] * r2 * r2
        p = yn * xn
        dtx = 2 * cam["p1"] * p + cam["p2"] * (r2 + 2 * yn * yn)
        dty = cam["p1"] * (r2 + 2 * xn * xn) + 2 * cam["p2"] * p
        yn, xn = (yd - dty) / rd, (xd - dtx) / rd
    depth = rng.uniform(4.0, 60.0, n)
    # camera-space point of the camera the target image belongs to (the right one when baseline != 0)
    pc = np.stack([xn * depth, yn * depth, depth, np.ones(n)], axis=0)
    th = rng.uniform(-0.3, 0.3, 3)
    Rx = np.array([[1, 0, 0], [0, np.cos(th[0]), -np.sin(th[0])], [0, np.sin(th[0]), np.cos(th[0])]])
    Ry = np.array([[np.cos(th[1]), 0, np.sin(th[1])], [0, 1, 0], [-np.sin(th[1]), 0, np.cos(th[1])]])
    Rz = np.array([[np.cos(th[2]), -np.sin(th[2]), 0], [np.sin(th[2]), np.cos(th[2]), 0], [0, 0, 1]])
    cw = np.eye(4)
    cw[:3, :3] = Rz @ Ry @ Rx
    cw[:3, 3] = rng.uniform(-3.0, 3.0, 3)
    Ti0 = np.eye(4)
    Ti0[0, 3] = -baseline
    world = (np.linalg.inv(Ti0 @ cw) @ pc)[:3].T.copy()
    world[~is_3d] = 0.0
    return dict(is_3d=is_3d, world=world, cw=cw, Ti0=Ti0, camera=cam, bad=bad, outside=outside)
