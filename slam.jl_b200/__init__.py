"""slam.jl_b200 -- host-side mirror of SLAM.jl's KLT front-end interface over libslamklt.so.

The reference's host language is Julia, which this image does not have; this Python layer mirrors the
reference's operator interface for the path (same names, argument meaning and error behaviour) so the
parity tests read like calls into the reference:

    LKPyramid(image, levels; sigma, reusable)   pyramid.jl:40      -> LKPyramid(ctx, image, levels, sigma=...)
    update!(pyr, image)                         pyramid.jl:81      -> pyr.update(image)
    copy!(dst, src)                             pyramid.jl:28      -> dst.copy_from(src)
    deepcopy(pyr)                               SLAM.jl:216        -> pyr.deepcopy()
    optflow!(disp, first, second, points, alg)  lucas_kanade.jl:9  -> optflow(disp, first, second, points, alg)
    fb_tracking!(prev, cur, keypoints; ...)     tracker.jl:70      -> fb_tracking(prev, cur, keypoints, ...)
    Extractor(max_points, radius, grid, cell)   extractor.jl:21    -> Extractor(...)
    detect(e, image, current_points)            extractor.jl:63    -> detect(ctx, e, image, current_points)
    optical_flow_matching!(mm, frame, from, to, stereo)  map_manager.jl:451 -> optical_flow_matching_frame(from, to, pixels, is_3d,
                                                                           world, cw, camera, ...)  (arrays instead of dictionaries)
    Camera(fx, fy, cx, cy, k1, k2, p1, p2, h, w; Ti0)   camera.jl:31      -> Camera(...)
    preprocess! + klt_tracking! over many frames         front_end.jl:454  -> StreamBatch(ctx, H, W, levels, n_frames, max_points)

Everything below is ctypes plumbing over the C ABI declared in include/slamklt.h.  There is NO CPU path:
if the CUDA library is missing or no B200 is visible, construction fails loudly.

Images are (H, W) arrays (converted to column-major like Julia's Matrix); points are (N, 2) float64 in
1-based (y, x) order.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("SLAMKLT_LIB", os.path.join(_CSRC, "libslamklt.so"))  # override for kernel experiments only

F64, F32, U8 = 0, 1, 2
MODE_UPDATE, MODE_CTOR = 0, 1
PLANES = {"layer": 0, "Iy": 1, "Ix": 2, "Iyy": 3, "Ixx": 4, "Iyx": 5, "Syy": 6, "Sxx": 7, "Syx": 8, "blur": 9,
          "Ryy": 10, "Rxx": 11, "Ryx": 12}  # R*: the device's row-prefix planes, (H, W + 1)

E_INVALID, E_CUDA, E_LAYERS, E_NODEVICE, E_CAPACITY = -1, -2, -3, -4, -5


class SlamKltError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[slamklt {code}] {msg}")
        self.code = code


class LKParams(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("window_size", C.c_int32), ("pyramid_levels", C.c_int32),
                ("reserved", C.c_int32), ("eigenvalue_threshold", C.c_double), ("epsilon", C.c_double),
                ("max_distance", C.c_double)]


class DetectParams(C.Structure):
    _fields_ = [("max_points", C.c_int32), ("radius", C.c_int32), ("grid_h", C.c_int32), ("grid_w", C.c_int32),
                ("cell_size", C.c_int32), ("reserved", C.c_int32), ("sigma_mask", C.c_double),
                ("min_response", C.c_double)]


class CameraC(C.Structure):
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double), ("k1", C.c_double),
                ("k2", C.c_double), ("p1", C.c_double), ("p2", C.c_double), ("height", C.c_int64), ("width", C.c_int64),
                ("Ti0", C.c_double * 16)]


class MatchingParams(C.Structure):
    _fields_ = [("lk", LKParams), ("stereo", C.c_int32), ("pyramid_levels_3d", C.c_int32), ("epipolar_error", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("lk_window_iters", C.c_uint64), ("lk_iters", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]


def build_library(force: bool = False) -> str:
    """Compile libslamklt.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(_CSRC, f) for f in ("api.cu", "pyramid.cu", "lk.cu", "lk_patch.cu", "lk_tma.cu", "detect.cu", "match.cu", "common.cuh")]
    srcs.append(os.path.join(_HERE, "..", "include", "slamklt.h"))
    stale = force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", _CSRC, "-j4", "libslamklt.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None

# every symbol include/slamklt.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "slamklt_last_error", "slamklt_version", "slamklt_device_count", "slamklt_ctx_create", "slamklt_ctx_destroy",
    "slamklt_ctx_sync", "slamklt_get_stats", "slamklt_profile", "slamklt_profile_report", "slamklt_timer_start", "slamklt_timer_stop", "slamklt_pyr_create",
    "slamklt_pyr_destroy", "slamklt_pyr_build", "slamklt_pyr_copy", "slamklt_pyr_clone", "slamklt_pyr_swap",
    "slamklt_pyr_info", "slamklt_pyr_level_dims", "slamklt_pyr_download", "slamklt_optflow", "slamklt_fb_track",
    "slamklt_flow_matching", "slamklt_optical_flow_matching", "slamklt_triangulate_stereo", "slamklt_describe", "slamklt_find_best_match", "slamklt_detect", "slamklt_batch_create", "slamklt_batch_destroy", "slamklt_batch_prime", "slamklt_batch_upload",
    "slamklt_batch_build", "slamklt_batch_track", "slamklt_batch_track_cross", "slamklt_batch_process", "slamklt_batch_download", "slamklt_batch_rotate",
    "slamklt_batch_step", "slamklt_batch_step_begin", "slamklt_batch_step_end", "slamklt_upload_rates", "slamklt_batch_slot", "slamklt_batch_detect", "slamklt_host_alloc", "slamklt_host_free",
]


def lib():
    """Load libslamklt.so.  Raises if it has not been built: the product never falls back to a CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SlamKltError(E_NODEVICE, f"{LIB_PATH} is missing: run __graft_entry__.build() (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        vp, dp, u8p, ip, i64p = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint8), C.POINTER(C.c_int), C.POINTER(C.c_int64)
        L.slamklt_last_error.restype = C.c_char_p
        L.slamklt_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.slamklt_ctx_destroy.argtypes = [vp]
        L.slamklt_ctx_sync.argtypes = [vp]
        L.slamklt_get_stats.argtypes = [vp, C.POINTER(Stats), C.c_int]
        L.slamklt_profile.argtypes = [vp, C.c_int]
        L.slamklt_profile_report.argtypes = [vp, C.c_char_p, C.c_size_t]
        L.slamklt_timer_start.argtypes = [vp]
        L.slamklt_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
        L.slamklt_pyr_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
        L.slamklt_pyr_destroy.argtypes = [vp, vp]
        L.slamklt_pyr_build.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_double, C.c_int]
        L.slamklt_pyr_copy.argtypes = [vp, vp, vp]
        L.slamklt_pyr_clone.argtypes = [vp, vp, C.POINTER(vp)]
        L.slamklt_pyr_swap.argtypes = [vp, vp, vp]
        L.slamklt_pyr_info.argtypes = [vp, ip, ip, ip, ip]
        L.slamklt_pyr_level_dims.argtypes = [vp, C.c_int, ip, ip]
        L.slamklt_pyr_download.argtypes = [vp, vp, C.c_int, C.c_int, dp]
        L.slamklt_optflow.argtypes = [vp, vp, vp, dp, dp, C.c_int, C.POINTER(LKParams), u8p, ip]
        L.slamklt_fb_track.argtypes = [vp, vp, vp, dp, dp, C.c_int, C.POINTER(LKParams), dp, u8p]
        L.slamklt_flow_matching.argtypes = [vp, vp, vp, dp, dp, u8p, C.c_int, C.POINTER(LKParams), C.c_int, dp, u8p]
        L.slamklt_optical_flow_matching.argtypes = [vp, vp, vp, dp, u8p, dp, dp, C.c_int, dp, C.POINTER(CameraC), C.POINTER(CameraC),
                                                    C.POINTER(MatchingParams), dp, dp, dp, u8p]
        L.slamklt_triangulate_stereo.argtypes = [vp, dp, dp, C.c_int, C.POINTER(CameraC), C.POINTER(CameraC), dp, C.c_double, dp, u8p]
        i32p, u32p = C.POINTER(C.c_int32), C.POINTER(C.c_uint32)
        L.slamklt_describe.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, i64p, C.c_int, i32p, C.c_int, C.c_int, C.c_double, u32p, u8p]
        L.slamklt_find_best_match.argtypes = [vp, u32p, C.c_int, C.c_int, i32p, C.c_int, i32p, i32p, i32p, C.c_int, C.c_int, i32p, i32p, i32p]
        L.slamklt_detect.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, dp, C.c_int, C.POINTER(DetectParams),
                                     i64p, C.c_int, ip]
        L.slamklt_batch_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
        L.slamklt_batch_destroy.argtypes = [vp, vp]
        L.slamklt_batch_prime.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_double, C.c_int]
        L.slamklt_batch_upload.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_size_t, dp, C.c_int]
        L.slamklt_batch_build.argtypes = [vp, vp, C.c_double, C.c_int]
        L.slamklt_batch_track.argtypes = [vp, vp, C.POINTER(LKParams)]
        L.slamklt_batch_track_cross.argtypes = [vp, vp, vp, C.POINTER(LKParams)]
        L.slamklt_batch_process.argtypes = [vp, vp, C.c_double, C.c_int, C.POINTER(LKParams)]
        L.slamklt_batch_download.argtypes = [vp, vp, dp, u8p]
        L.slamklt_batch_rotate.argtypes = [vp, vp]
        L.slamklt_batch_step.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_size_t, dp, C.c_int, C.c_double, C.c_int,
                                         C.POINTER(LKParams), dp, u8p]
        L.slamklt_batch_step_begin.argtypes = L.slamklt_batch_step.argtypes
        L.slamklt_batch_step_end.argtypes = [vp, vp]
        L.slamklt_upload_rates.argtypes = [vp, dp, dp]
        L.slamklt_batch_slot.argtypes = [vp, C.c_int, C.POINTER(vp)]
        L.slamklt_batch_detect.argtypes = [vp, vp, dp, C.c_int, C.POINTER(DetectParams), i64p, C.c_int, ip]
        L.slamklt_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
        L.slamklt_host_free.argtypes = [vp]
        _lib = L
    return _lib


def _ck(rc):
    if rc != 0:
        raise SlamKltError(rc, lib().slamklt_last_error().decode())


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _image(img):
    """(array in column-major order, dtype code).  float64 / float32 / uint8 are passed through unconverted."""
    a = np.asarray(img)
    if a.dtype == np.float64:
        code = F64
    elif a.dtype == np.float32:
        code = F32
    elif a.dtype == np.uint8:
        code = U8
    else:
        a, code = a.astype(np.float64), F64
    if a.ndim != 2:
        raise SlamKltError(E_INVALID, "image must be 2-D")
    return np.asfortranarray(a), code


class Context:
    """One CUDA device + stream.  Fails loudly when no B200 is visible."""

    def __init__(self, device: int = 0):
        h = C.c_void_p()
        _ck(lib().slamklt_ctx_create(int(device), C.byref(h)))
        self._h = h
        self.device = device

    def sync(self):
        _ck(lib().slamklt_ctx_sync(self._h))

    def stats(self, reset=False) -> dict:
        s = Stats()
        _ck(lib().slamklt_get_stats(self._h, C.byref(s), 1 if reset else 0))
        return {k: int(getattr(s, k)) for k, _ in Stats._fields_}

    def upload_rates(self) -> dict:
        """Measured rates of the two upload engines of StreamBatch.step (GB of Float64 source per second; 0 = not measured)."""
        a, b = C.c_double(0.0), C.c_double(0.0)
        _ck(lib().slamklt_upload_rates(self._h, C.byref(a), C.byref(b)))
        return {"repack_GBps": a.value / 1e9, "plain_copy_GBps": b.value / 1e9}

    def profile(self, enable: bool):
        _ck(lib().slamklt_profile(self._h, 1 if enable else 0))

    def profile_report(self) -> dict:
        """{kernel name: (launches, total_ms)} accumulated since profile(True)."""
        buf = C.create_string_buffer(1 << 16)
        _ck(lib().slamklt_profile_report(self._h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, n, ms = line.split()
            out[name] = (int(n), float(ms))
        return out

    def timer_start(self):
        _ck(lib().slamklt_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        _ck(lib().slamklt_timer_stop(self._h, C.byref(ms)))
        return float(ms.value)

    def close(self):
        if getattr(self, "_h", None):
            lib().slamklt_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LucasKanade:
    """lucas_kanade.jl:1-7"""

    def __init__(self, iterations=30, window_size=9, pyramid_levels=3, eigenvalue_threshold=1e-4, eps=1e-2):
        self.iterations, self.window_size, self.pyramid_levels = iterations, window_size, pyramid_levels
        self.eigenvalue_threshold, self.eps = eigenvalue_threshold, eps

    def _c(self, max_distance=0.5):
        return LKParams(self.iterations, self.window_size, self.pyramid_levels, 0, self.eigenvalue_threshold,
                        self.eps, float(max_distance))


class LKPyramid:
    """LKPyramid (pyramid.jl:16-96), device resident."""

    def __init__(self, ctx: Context, image=None, levels: int = 3, sigma: float = 1.0, shape=None, _handle=None,
                 _owner=None):
        self.ctx = ctx
        self._owner = _owner
        if _handle is not None:
            self._h = _handle
        else:
            if image is not None:
                image, _ = _image(image)
                shape = image.shape
            h = C.c_void_p()
            _ck(lib().slamklt_pyr_create(ctx._h, int(shape[0]), int(shape[1]), int(levels), C.byref(h)))
            self._h = h
            if image is not None:
                self._build(image, sigma, MODE_CTOR)  # the constructor path: NA blur border, Fill(0) Scharr

    def _build(self, image, sigma, mode):
        a, code = _image(image)
        H, W = self.shape
        if a.shape != (H, W):
            raise SlamKltError(E_INVALID, f"image shape {a.shape} != pyramid shape {(H, W)}")
        _ck(lib().slamklt_pyr_build(self.ctx._h, self._h, a.ctypes.data_as(C.c_void_p), code, H, float(sigma), mode))
        return self

    def update(self, image, sigma: float = 1.0):
        """update!(lk, img; sigma) pyramid.jl:81-96"""
        return self._build(image, sigma, MODE_UPDATE)

    def _info(self):
        H, W, L, b = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _ck(lib().slamklt_pyr_info(self._h, C.byref(H), C.byref(W), C.byref(L), C.byref(b)))
        return H.value, W.value, L.value, bool(b.value)

    @property
    def shape(self):
        return self._info()[:2]

    @property
    def levels(self):
        return self._info()[2]

    def has_gradients(self):
        return self._info()[3]

    def level_shape(self, level):
        H, W = C.c_int(), C.c_int()
        _ck(lib().slamklt_pyr_level_dims(self._h, level, C.byref(H), C.byref(W)))
        return H.value, W.value

    def plane(self, level: int, name: str) -> np.ndarray:
        """lk.layers[level+1], lk.Iy[level+1], ... as Float64 (H_l, W_l); the device's own row-prefix planes "Ryy" / "Rxx" /
        "Ryx" (exclusive prefix sums along x of the smoothed products, what the tracking kernel reads) as (H_l, W_l + 1)."""
        H, W = self.level_shape(level)
        out = np.empty((H, W + (1 if name in ("Ryy", "Rxx", "Ryx") else 0)), dtype=np.float64, order="F")
        _ck(lib().slamklt_pyr_download(self.ctx._h, self._h, level, PLANES[name], _dp(out)))
        return out

    def copy_from(self, src: "LKPyramid"):
        """copy!(dst, src) pyramid.jl:28-38"""
        _ck(lib().slamklt_pyr_copy(self.ctx._h, self._h, src._h))
        return self

    def deepcopy(self) -> "LKPyramid":
        h = C.c_void_p()
        _ck(lib().slamklt_pyr_clone(self.ctx._h, self._h, C.byref(h)))
        return LKPyramid(self.ctx, _handle=h)

    def swap(self, other: "LKPyramid"):
        _ck(lib().slamklt_pyr_swap(self.ctx._h, self._h, other._h))

    def __del__(self):
        try:
            if self._owner is None and getattr(self, "_h", None) and self.ctx._h:
                lib().slamklt_pyr_destroy(self.ctx._h, self._h)
                self._h = None
        except Exception:
            pass


def optflow(displacement, first: LKPyramid, second: LKPyramid, points, algorithm: LucasKanade):
    """optflow! (lucas_kanade.jl:9-100).  Returns (displacement, status, n_good); raises like the reference's
    throw("Not enough layers in pyramids.") when a pyramid is too shallow."""
    pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 2)
    disp = np.array(displacement, dtype=np.float64, copy=True).reshape(-1, 2)
    st = np.zeros(len(pts), dtype=np.uint8)
    ng = C.c_int()
    p = algorithm._c()
    _ck(lib().slamklt_optflow(first.ctx._h, first._h, second._h, _dp(pts), _dp(disp), len(pts), C.byref(p),
                              st.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(ng)))
    return disp, st.astype(bool), ng.value


def fb_tracking(previous: LKPyramid, current: LKPyramid, keypoints, displacement=None, iterations=30, window_size=11,
                pyramid_levels=3, max_distance=0.5, eigenvalue_threshold=1e-4, eps=1e-2):
    """fb_tracking! (tracker.jl:17-82).  Returns None for empty input (tracker.jl:24), else
    (new_keypoints, status, forward_status); new_keypoints rows are NaN where the forward pass failed
    (undefined in the reference)."""
    pts = np.ascontiguousarray(keypoints, dtype=np.float64).reshape(-1, 2)
    n = len(pts)
    if n == 0:
        return None
    out = np.full((n, 2), np.nan)
    st = np.zeros(n, dtype=np.uint8)
    d = None if displacement is None else np.ascontiguousarray(displacement, dtype=np.float64).reshape(-1, 2)
    p = LKParams(iterations, window_size, pyramid_levels, 0, eigenvalue_threshold, eps, float(max_distance))
    _ck(lib().slamklt_fb_track(previous.ctx._h, previous._h, current._h, _dp(pts), None if d is None else _dp(d), n,
                               C.byref(p), _dp(out), st.ctypes.data_as(C.POINTER(C.c_uint8))))
    return out, (st & 1).astype(bool), ((st >> 1) & 1).astype(bool)


def optical_flow_matching(from_pyramid: LKPyramid, to_pyramid: LKPyramid, pixels, prior_displacement, is_3d, window_size=9,
                          pyramid_levels=3, pyramid_levels_3d=1, max_distance=1.0, iterations=30, eigenvalue_threshold=1e-4,
                          eps=1e-2):
    """Tracking part of optical_flow_matching! (map_manager.jl:451-564) in one launch: 3-D keypoints first with their prior
    on pyramid_levels_3d levels, failures and 2-D keypoints on pyramid_levels levels from a zero displacement.
    Returns (new_keypoints, status, tracked_by_prior_pass)."""
    pts = np.ascontiguousarray(pixels, dtype=np.float64).reshape(-1, 2)
    n = len(pts)
    pri = np.ascontiguousarray(prior_displacement, dtype=np.float64).reshape(-1, 2)
    flg = np.ascontiguousarray(is_3d, dtype=np.uint8).reshape(-1)
    assert len(pri) == n and len(flg) == n
    out = np.full((n, 2), np.nan)
    st = np.zeros(n, dtype=np.uint8)
    if n == 0:
        return out, st.astype(bool), st.astype(bool)
    p = LKParams(iterations, window_size, pyramid_levels, 0, eigenvalue_threshold, eps, float(max_distance))
    _ck(lib().slamklt_flow_matching(from_pyramid.ctx._h, from_pyramid._h, to_pyramid._h, _dp(pts), _dp(pri),
                                    flg.ctypes.data_as(C.POINTER(C.c_uint8)), n, C.byref(p), int(pyramid_levels_3d), _dp(out),
                                    st.ctypes.data_as(C.POINTER(C.c_uint8))))
    return out, (st & 1).astype(bool), ((st >> 2) & 1).astype(bool)


class Camera:
    """Camera (camera.jl:1-46): intrinsics, distortion, image size, Ti0 (4 x 4, this camera <- camera 0)."""

    def __init__(self, fx, fy, cx, cy, k1=0.0, k2=0.0, p1=0.0, p2=0.0, height=0, width=0, Ti0=None):
        self.fx, self.fy, self.cx, self.cy = float(fx), float(fy), float(cx), float(cy)
        self.k1, self.k2, self.p1, self.p2 = float(k1), float(k2), float(p1), float(p2)
        self.height, self.width = int(height), int(width)
        self.Ti0 = np.eye(4) if Ti0 is None else np.array(Ti0, dtype=np.float64).reshape(4, 4)

    def _c(self):
        t = np.asfortranarray(self.Ti0).ravel(order="F")  # SMatrix storage order
        return CameraC(self.fx, self.fy, self.cx, self.cy, self.k1, self.k2, self.p1, self.p2, self.height, self.width,
                       (C.c_double * 16)(*t))


def optical_flow_matching_frame(from_pyramid: LKPyramid, to_pyramid: LKPyramid, pixels, is_3d, world, cw, camera: Camera,
                                right_camera: Camera = None, undistorted=None, stereo=False, window_size=9, pyramid_levels=3,
                                pyramid_levels_3d=1, max_distance=1.0, epipolar_error=2.0, iterations=30,
                                eigenvalue_threshold=1e-4, eps=1e-2):
    """optical_flow_matching!(map_manager, frame, from, to, stereo) (map_manager.jl:451-564) with the per-keypoint geometry on
    the device: projection of the 3-D keypoints' map points + prior, both tracking passes, and update_keypoint! /
    maybe_stereo_update! of the tracked ones.  Arrays replace the dictionaries: pixels[i] = kp.pixel, is_3d[i], world[i] =
    map-point position, undistorted[i] = kp.undistorted_pixel (stereo), cw = frame.cw (4 x 4).
    Returns (pixel, undistorted_pixel, position, status); status bits as in include/slamklt.h."""
    pix = np.ascontiguousarray(pixels, dtype=np.float64).reshape(-1, 2)
    n = len(pix)
    flg = np.ascontiguousarray(is_3d, dtype=np.uint8).reshape(-1)
    wor = np.ascontiguousarray(world, dtype=np.float64).reshape(-1, 3)
    assert len(flg) == n and len(wor) == n
    und = None if undistorted is None else np.ascontiguousarray(undistorted, dtype=np.float64).reshape(-1, 2)
    T = np.ascontiguousarray(np.asarray(cw, dtype=np.float64).reshape(4, 4).ravel(order="F"))
    out_pix = np.full((n, 2), np.nan); out_und = np.full((n, 2), np.nan); out_pos = np.full((n, 3), np.nan)
    st = np.zeros(n, dtype=np.uint8)
    mp = MatchingParams(LKParams(iterations, window_size, pyramid_levels, 0, eigenvalue_threshold, eps, float(max_distance)),
                        1 if stereo else 0, int(pyramid_levels_3d), float(epipolar_error))
    cam = camera._c()
    rcam = None if right_camera is None else right_camera._c()
    _ck(lib().slamklt_optical_flow_matching(
        from_pyramid.ctx._h, from_pyramid._h, to_pyramid._h, _dp(pix), flg.ctypes.data_as(C.POINTER(C.c_uint8)), _dp(wor),
        None if und is None else _dp(und), n, _dp(T), C.byref(cam), None if rcam is None else C.byref(rcam), C.byref(mp),
        _dp(out_pix), _dp(out_und), _dp(out_pos), st.ctypes.data_as(C.POINTER(C.c_uint8))))
    return out_pix, out_und, out_pos, st


def triangulate_stereo(ctx: Context, undistorted, right_undistorted, camera: Camera, right_camera: Camera, wc, max_error=3.0):
    """triangulate_stereo! (mapper.jl:142-183) for the stereo keypoints the caller selected: DLT triangulation from the left /
    right undistorted pixels, depth and reprojection checks, world point = wc * point.  Returns (world (n, 3), status (n,)):
    status 1 = update_mappoint!, 2..5 = remove_stereo_keypoint! (include/slamklt.h)."""
    und = np.ascontiguousarray(undistorted, dtype=np.float64).reshape(-1, 2)
    rund = np.ascontiguousarray(right_undistorted, dtype=np.float64).reshape(-1, 2)
    n = len(und)
    assert len(rund) == n
    T = np.ascontiguousarray(np.asarray(wc, dtype=np.float64).reshape(4, 4).ravel(order="F"))
    world = np.full((n, 3), np.nan)
    st = np.zeros(n, dtype=np.uint8)
    cam, rcam = camera._c(), right_camera._c()
    _ck(lib().slamklt_triangulate_stereo(ctx._h, _dp(und), _dp(rund), n, C.byref(cam), C.byref(rcam), _dp(T), float(max_error), _dp(world),
                                         st.ctypes.data_as(C.POINTER(C.c_uint8))))
    return world, st


def brief_pairs_stand_in(n_bits=256, window=9, seed=123):
    """A sampling pattern with the distribution of ImageFeatures' `gaussian(size, window)` (normal with std window^2 / 25, offsets
    floored and kept inside +-ceil(window / 2)).  It is NOT the reference's pattern: that one is drawn from Julia's RNG
    (Random.seed!(123)) and has to be exported from Julia once (julia/SlamKLT.jl does).  Tests and benchmarks use this stand-in."""
    rng = np.random.default_rng(seed)
    lim = -(-window // 2)
    out = []
    while len(out) < 4 * n_bits:
        v = rng.normal(0.0, window * window / 25.0)
        if -lim <= v <= lim:
            out.append(int(np.floor(v)))
    return np.asarray(out, dtype=np.int32).reshape(n_bits, 4)


def describe(ctx: Context, image, keypoints, pairs, window=9, sigma=2 ** 0.5):
    """describe(e, image, keypoints) (extractor.jl:103-105 -> create_descriptor with BRIEF(size = len(pairs))).  keypoints: (n, 2)
    int64 1-based (y, x); pairs: (n_bits, 4) int32 (dy1, dx1, dy2, dx2).  Returns (descriptors (m, n_bits / 32) uint32, keypoints
    (m, 2)) of the keypoints far enough from the border, like the reference."""
    a, code = _image(image)
    H, W = a.shape
    kps = np.ascontiguousarray(keypoints, dtype=np.int64).reshape(-1, 2)
    pr = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 4)
    n, n_bits = len(kps), len(pr)
    desc = np.zeros((n, n_bits // 32), dtype=np.uint32)
    valid = np.zeros(n, dtype=np.uint8)
    _ck(lib().slamklt_describe(ctx._h, a.ctypes.data_as(C.c_void_p), code, H, W, H, kps.ctypes.data_as(C.POINTER(C.c_int64)), n,
                               pr.ctypes.data_as(C.POINTER(C.c_int32)), n_bits, int(window), float(sigma),
                               desc.ctypes.data_as(C.POINTER(C.c_uint32)), valid.ctypes.data_as(C.POINTER(C.c_uint8))))
    keep = valid.astype(bool)
    return desc[keep], kps[keep]


def find_best_match(ctx: Context, descriptors, set_offsets, target_sets, cand_offsets, candidates, max_distance):
    """Descriptor side of find_best_match (mapper.jl:392-462) for many targets: see include/slamklt.h.  Returns (best_pos,
    best_dist, second_dist) int32 arrays, distances in bits."""
    d = np.ascontiguousarray(descriptors, dtype=np.uint32)
    d = d.reshape(len(d), -1) if d.ndim == 2 else d.reshape(0, 8)
    so = np.ascontiguousarray(set_offsets, dtype=np.int32)
    ts = np.ascontiguousarray(target_sets, dtype=np.int32)
    co = np.ascontiguousarray(cand_offsets, dtype=np.int32)
    ca = np.ascontiguousarray(candidates, dtype=np.int32)
    nt = len(ts)
    bp, bd, sd = np.zeros(nt, np.int32), np.zeros(nt, np.int32), np.zeros(nt, np.int32)
    ip = C.POINTER(C.c_int32)
    _ck(lib().slamklt_find_best_match(ctx._h, d.ctypes.data_as(C.POINTER(C.c_uint32)), len(d), d.shape[1], so.ctypes.data_as(ip), len(so) - 1,
                                      ts.ctypes.data_as(ip), co.ctypes.data_as(ip), ca.ctypes.data_as(ip), nt, int(max_distance),
                                      bp.ctypes.data_as(ip), bd.ctypes.data_as(ip), sd.ctypes.data_as(ip)))
    return bp, bd, sd


class Extractor:
    """Extractor (extractor.jl:7-22); `describe` (BRIEF) is the free function above (the sampling pattern is an argument)."""

    def __init__(self, max_points, radius, grid_resolution, cell_size):
        self.max_points, self.radius = int(max_points), int(radius)
        self.grid_resolution, self.cell_size = (int(grid_resolution[0]), int(grid_resolution[1])), int(cell_size)

    def _c(self, sigma_mask=3.0, min_response=1e-4):
        return DetectParams(self.max_points, self.radius, self.grid_resolution[0], self.grid_resolution[1],
                            self.cell_size, 0, float(sigma_mask), float(min_response))


def detect(ctx: Context, e: Extractor, image, current_points, sigma_mask=3.0, min_response=1e-4) -> np.ndarray:
    """detect (extractor.jl:63-95).  Returns (n, 2) int64 1-based (y, x) keypoints."""
    a, code = _image(image)
    H, W = a.shape
    cur = np.ascontiguousarray(current_points, dtype=np.float64).reshape(-1, 2)
    k_cell = -(-max(e.max_points - len(cur), 0) // (e.grid_resolution[0] * e.grid_resolution[1]))
    cap = max(1, min(k_cell, e.cell_size * e.cell_size) * e.grid_resolution[0] * e.grid_resolution[1])
    out = np.empty((cap, 2), dtype=np.int64)
    n = C.c_int()
    p = e._c(sigma_mask, min_response)
    _ck(lib().slamklt_detect(ctx._h, a.ctypes.data_as(C.c_void_p), code, H, W, H, _dp(cur) if len(cur) else None, len(cur),
                             C.byref(p), out.ctypes.data_as(C.POINTER(C.c_int64)), cap, C.byref(n)))
    return out[:n.value].copy()


class PinnedArray:
    """numpy view over cudaMallocHost memory."""

    def __init__(self, shape, dtype):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        _ck(lib().slamklt_host_alloc(max(self.nbytes, 1), C.byref(p)))
        self._p = p
        buf = (C.c_uint8 * max(self.nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._p:
            self.array = None
            lib().slamklt_host_free(self._p)
            self._p = None


class StreamBatch:
    """n_frames consecutive frames of one stream processed in one go (preprocess! + klt_tracking! of
    front_end.jl:454-481, batched).  Frames are (n_frames, W, H)-shaped C arrays, i.e. each frame column-major."""

    def __init__(self, ctx: Context, H, W, levels, n_frames, max_points):
        self.ctx, self.H, self.W, self.levels, self.n_frames, self.max_points = ctx, H, W, levels, n_frames, max_points
        h = C.c_void_p()
        _ck(lib().slamklt_batch_create(ctx._h, H, W, levels, n_frames, max_points, C.byref(h)))
        self._h = h

    @staticmethod
    def pack_frames(frames) -> np.ndarray:
        """(n, H, W) array -> (n, W, H) C-contiguous (each frame column-major), dtype preserved."""
        f = np.asarray(frames)
        return np.ascontiguousarray(np.transpose(f, (0, 2, 1)))

    @staticmethod
    def _code(a):
        return {np.dtype(np.float64): F64, np.dtype(np.float32): F32, np.dtype(np.uint8): U8}[a.dtype]

    def prime(self, image, sigma=1.0, mode=MODE_CTOR):
        a, code = _image(image)
        _ck(lib().slamklt_batch_prime(self.ctx._h, self._h, a.ctypes.data_as(C.c_void_p), code, self.H, float(sigma), mode))

    def upload(self, packed_frames: np.ndarray, points: np.ndarray):
        assert packed_frames.shape == (self.n_frames, self.W, self.H) and packed_frames.flags.c_contiguous
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(self.n_frames, -1, 2)
        _ck(lib().slamklt_batch_upload(self.ctx._h, self._h, packed_frames.ctypes.data_as(C.c_void_p), self._code(packed_frames),
                                       self.H, packed_frames.strides[0], _dp(pts), pts.shape[1]))
        self._npts = pts.shape[1]

    def build(self, sigma=1.0, mode=MODE_UPDATE):
        _ck(lib().slamklt_batch_build(self.ctx._h, self._h, float(sigma), mode))

    def track(self, alg: LucasKanade, max_distance=1.0):
        p = alg._c(max_distance)
        _ck(lib().slamklt_batch_track(self.ctx._h, self._h, C.byref(p)))

    def track_cross(self, to: "StreamBatch", alg: LucasKanade, max_distance=1.0):
        """Stereo matching of two batches (mapper.jl:51-60 for n_frames keyframes at once): pair i tracks the points uploaded to
        `to` from frame i of this batch (left) to frame i of `to` (right); fetch the results with to.download()."""
        p = alg._c(max_distance)
        _ck(lib().slamklt_batch_track_cross(self.ctx._h, self._h, to._h, C.byref(p)))

    def process(self, alg: LucasKanade, max_distance=1.0, sigma=1.0, mode=MODE_UPDATE):
        """build + track of the uploaded frames in one asynchronous, internally pipelined call."""
        p = alg._c(max_distance)
        _ck(lib().slamklt_batch_process(self.ctx._h, self._h, float(sigma), mode, C.byref(p)))

    def download(self, out_pts=None, status=None):
        if out_pts is None:
            out_pts = np.empty((self.n_frames, self._npts, 2))
        if status is None:
            status = np.empty((self.n_frames, self._npts), dtype=np.uint8)
        _ck(lib().slamklt_batch_download(self.ctx._h, self._h, _dp(out_pts), status.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out_pts, status

    def rotate(self):
        _ck(lib().slamklt_batch_rotate(self.ctx._h, self._h))

    def step(self, packed_frames, points, alg: LucasKanade, max_distance=1.0, sigma=1.0, mode=MODE_UPDATE, out_pts=None,
             status=None):
        """Whole step through host buffers (upload, build, track, download, rotate) in one C call."""
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(self.n_frames, -1, 2)
        n = pts.shape[1]
        if out_pts is None:
            out_pts = np.empty((self.n_frames, n, 2))
        if status is None:
            status = np.empty((self.n_frames, n), dtype=np.uint8)
        p = alg._c(max_distance)
        _ck(lib().slamklt_batch_step(self.ctx._h, self._h, packed_frames.ctypes.data_as(C.c_void_p), self._code(packed_frames),
                                     self.H, packed_frames.strides[0], _dp(pts), n, float(sigma), mode, C.byref(p), _dp(out_pts),
                                     status.ctypes.data_as(C.POINTER(C.c_uint8))))
        self._npts = n
        return out_pts, status

    def step_begin(self, packed_frames, points, alg: LucasKanade, max_distance=1.0, sigma=1.0, mode=MODE_UPDATE, out_pts=None,
                   status=None):
        """First half of step(): queue everything and return without waiting.  Returns (out_pts, status), valid after step_end().
        The frames, points and result arrays are kept alive by the batch until then; frames and points must not be modified."""
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(self.n_frames, -1, 2)
        n = pts.shape[1]
        if out_pts is None:
            out_pts = np.empty((self.n_frames, n, 2))
        if status is None:
            status = np.empty((self.n_frames, n), dtype=np.uint8)
        p = alg._c(max_distance)
        _ck(lib().slamklt_batch_step_begin(self.ctx._h, self._h, packed_frames.ctypes.data_as(C.c_void_p), self._code(packed_frames),
                                           self.H, packed_frames.strides[0], _dp(pts), n, float(sigma), mode, C.byref(p), _dp(out_pts),
                                           status.ctypes.data_as(C.POINTER(C.c_uint8))))
        self._npts = n
        self._inflight = (packed_frames, pts, out_pts, status)
        return out_pts, status

    def step_end(self):
        """Second half of step(): wait for the results of step_begin(), rotate, return (out_pts, status)."""
        _ck(lib().slamklt_batch_step_end(self.ctx._h, self._h))
        _, _, out_pts, status = self._inflight
        self._inflight = None
        return out_pts, status

    def slot(self, i) -> LKPyramid:
        h = C.c_void_p()
        _ck(lib().slamklt_batch_slot(self._h, i, C.byref(h)))
        return LKPyramid(self.ctx, _handle=h, _owner=self)

    def detect(self, e: Extractor, current_points=None, sigma_mask=3.0, min_response=1e-4):
        """Batched detect on the frames last uploaded.  Returns a list of (n_f, 2) int64 arrays."""
        n_cur = 0 if current_points is None else np.asarray(current_points).reshape(self.n_frames, -1, 2).shape[1]
        cur = None if n_cur == 0 else np.ascontiguousarray(current_points, dtype=np.float64).reshape(self.n_frames, -1, 2)
        cells = e.grid_resolution[0] * e.grid_resolution[1]
        k_cell = -(-max(e.max_points - n_cur, 0) // cells)
        cap = max(1, min(k_cell, e.cell_size * e.cell_size) * cells)
        out = np.empty((self.n_frames, cap, 2), dtype=np.int64)
        n = np.zeros(self.n_frames, dtype=np.int32)
        p = e._c(sigma_mask, min_response)
        _ck(lib().slamklt_batch_detect(self.ctx._h, self._h, None if cur is None else _dp(cur), n_cur, C.byref(p),
                                       out.ctypes.data_as(C.POINTER(C.c_int64)), cap, n.ctypes.data_as(C.POINTER(C.c_int))))
        return [out[f, :n[f]].copy() for f in range(self.n_frames)]

    def close(self):
        if getattr(self, "_h", None) and self.ctx._h:
            lib().slamklt_batch_destroy(self.ctx._h, self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
