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


def num_threads():
    return lib().orc_num_threads()


def set_threads(n):
    lib().orc_set_threads(int(n))
