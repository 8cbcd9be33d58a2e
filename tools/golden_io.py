"""Exchange of golden vectors with the real SLAM.jl (julia/dump_golden.jl).

    python tools/golden_io.py export DIR     writes the inputs as raw Float64 .bin files
    python tools/golden_io.py check DIR      compares every array dump_golden.jl wrote with the CPU oracle

.bin layout: Int64 ndims, Int64 dims..., payload in column-major order (Julia's own); *_i64 payloads are Int64.
Until `check` has been run against a real Julia dump, parity of the oracle with the reference is UNPINNED.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from slamklt import synth  # noqa: E402

LEVELS, WINDOW, MAXD = 3, 9, 1.0
EXT = (300, 8, 4, 6, 35)
CAMERA = dict(synth.KITTI_CAMERA, fx=180.0, fy=180.0, cx=100.0, cy=72.0, height=144, width=200)


def wr(path, a, dtype=np.float64):
    a = np.asarray(a, dtype=dtype)
    with open(path, "wb") as f:
        np.array([a.ndim], dtype=np.int64).tofile(f)
        np.array(a.shape, dtype=np.int64).tofile(f)
        np.asfortranarray(a).ravel(order="F").tofile(f)


def rd(path, dtype=np.float64):
    with open(path, "rb") as f:
        nd = int(np.fromfile(f, dtype=np.int64, count=1)[0])
        dims = tuple(int(d) for d in np.fromfile(f, dtype=np.int64, count=nd))
        return np.fromfile(f, dtype=dtype).reshape(dims, order="F")


def scene():
    fr, aff = synth.make_sequence(9090, 2, H=144, W=200)
    f = synth.to_f64(fr)
    pts = synth.random_keypoints(3, 120, 144, 200, border=2.0)
    sc = synth.matching_scene(5, pts, synth.true_flow(aff, 0, 1, pts), camera=CAMERA)
    return f, pts, sc


def export(d):
    os.makedirs(d, exist_ok=True)
    f, pts, sc = scene()
    wr(os.path.join(d, "img0.bin"), f[0]); wr(os.path.join(d, "img1.bin"), f[1])
    wr(os.path.join(d, "pts.bin"), pts.T)
    wr(os.path.join(d, "meta.bin"), [LEVELS, WINDOW, MAXD, *EXT])
    wr(os.path.join(d, "camera.bin"), [CAMERA[k] for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2", "height", "width")])
    wr(os.path.join(d, "cw.bin"), sc["cw"]); wr(os.path.join(d, "world.bin"), sc["world"].T)
    print("inputs written to", d)


def check(d):
    f, pts, sc = scene()
    worst = {}

    def cmp(name, ours, tol, dtype=np.float64, exact=False):
        theirs = rd(os.path.join(d, name + ".bin"), dtype)
        ours = np.asarray(ours)
        if theirs.shape != ours.shape:
            worst[name] = f"shape {theirs.shape} != {ours.shape}"
            return
        if exact:
            worst[name] = "equal" if np.array_equal(theirs, ours) else f"{int(np.sum(theirs != ours))} entries differ"
            return
        m = np.isfinite(theirs) & np.isfinite(ours)
        err = float(np.max(np.abs(theirs[m] - ours[m]))) if m.any() else 0.0
        worst[name] = f"{err:.3e}" + ("" if err <= tol and np.array_equal(np.isfinite(theirs), np.isfinite(ours)) else "  <-- MISMATCH")

    p0, p1 = O.LKPyramid(f[0], LEVELS, mode="ctor"), O.LKPyramid(f[1], LEVELS, mode="ctor")
    for tag, p in (("ctor0", p0), ("ctor1", p1)):
        for l in range(LEVELS + 1):
            for n in ("layer", "Iy", "Ix", "Iyy", "Ixx", "Iyx"):
                cmp(f"{tag}_{n}{l}", p.plane(l, n), 1e-9 if n[:2] != "Iy" and n[:2] != "Ix" or len(n) == 2 else 1e-7)
    p1.update(f[1])
    for l in range(LEVELS + 1):
        for n in ("layer", "Iy", "Ix", "Iyy", "Ixx", "Iyx"):
            cmp(f"upd1_{n}{l}", p1.plane(l, n), 1e-7)
    alg = O.LucasKanade(30, WINDOW, LEVELS)
    disp, st, _ = O.optflow(np.zeros_like(pts), p0, p1, pts, alg)
    cmp("optflow_status_i64", st.astype(np.int64), 0, np.int64, exact=True)
    cmp("optflow_disp", disp.T, 1e-6)
    new, fst, _ = O.fb_tracking(p0, p1, pts, window_size=WINDOW, pyramid_levels=LEVELS, max_distance=MAXD)
    cmp("fb_status_i64", fst.astype(np.int64), 0, np.int64, exact=True)
    cmp("fb_new", np.where(fst[:, None], new, np.nan).T, 1e-6)
    e = O.Extractor(EXT[0], EXT[1], (EXT[2], EXT[3]), EXT[4])
    cmp("detect_i64", O.detect(e, f[0], pts[:10]).T, 0, np.int64, exact=True)
    cmp("detect_nomask_i64", O.detect(e, f[0], np.zeros((0, 2))).T, 0, np.int64, exact=True)
    cam = O.Camera(**CAMERA)
    cmp("cam_proj", O.project_world_distort(cam, sc["cw"], sc["world"]).T, 1e-9)
    und = O.undistort_point(cam, pts)
    cmp("cam_undist", und.T, 1e-10); cmp("cam_backproject", O.backproject(cam, und).T, 1e-12)
    bad = 0
    for k, v in worst.items():
        print(f"{k:24s} {v}")
        bad += ("MISMATCH" in v) or ("differ" in v) or ("shape" in v)
    print("PINNED: the oracle reproduces the reference on this scene" if not bad else f"{bad} arrays do not match")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(export(sys.argv[2]) if sys.argv[1] == "export" else check(sys.argv[2]))
