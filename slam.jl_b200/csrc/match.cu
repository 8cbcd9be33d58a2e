// Per-keypoint geometry around the tracking kernel of optical_flow_matching! (first half):
//   prologue  map_manager.jl:472-507  project the map point of every 3-D keypoint (frame.jl:458-484, camera.jl:60-85,112-128),
//                                     in-image test (camera.jl:87-95), prior displacement scale .* (projection .- pixel)
//   epilogue  frame.jl:252-270        update_keypoint!: undistort_point + backproject (camera.jl:97-128,130-143)
//             map_manager.jl:579-590  maybe_stereo_update!: epipolar gate, y taken from the left pixel,
//             frame.jl:272-287        update_stereo_keypoint!
// Float64 like the reference; this file is compiled with -fmad=false and every expression keeps the reference's
// evaluation order, so for identical tracked pixels the outputs are bit-identical to the CPU arithmetic.
#include "common.cuh"

namespace sk {
namespace {

// undistort_pdn_point, camera.jl:106-128; (py, px) is the pre-divided, normalised point in (y, x) order
__device__ __forceinline__ void undistort_pdn(const MatchCam& c, double py, double px, double& oy, double& ox) {
    const double sy = py * py, sx = px * px;
    const double r2 = sy + sx;
    const double rd = 1.0 + c.k1 * r2 + c.k2 * (r2 * r2);
    const double p = py * px;
    const double dtx = 2.0 * c.p1 * p + c.p2 * (r2 + 2.0 * sy);
    const double dty = c.p1 * (r2 + 2.0 * sx) + 2.0 * c.p2 * p;
    oy = (rd * py + dty) * c.fy + c.cy;
    ox = (rd * px + dtx) * c.fx + c.cx;
}

// undistort_point, camera.jl:97-104
__device__ __forceinline__ void undistort_point(const MatchCam& c, double y, double x, double& oy, double& ox) {
    undistort_pdn(c, (y - c.cy) / c.fy, (x - c.cx) / c.fx, oy, ox);
}

__global__ void k_match_prior(const MatchArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    uint8_t flag = 0;
    double dy = 0.0, dx = 0.0;
    if (a.is_3d[i]) {
        const double X = a.world[3 * (size_t)i], Y = a.world[3 * (size_t)i + 1], Z = a.world[3 * (size_t)i + 2];
        // (T * [X, Y, Z, 1])[1:3], T column-major (frame.jl:458-468); T = cw, or Ti0 * cw for the right camera
        const double* T = a.T;
        const double xc = T[0] * X + T[4] * Y + T[8] * Z + T[12];
        const double yc = T[1] * X + T[5] * Y + T[9] * Z + T[13];
        const double zc = T[2] * X + T[6] * Y + T[10] * Z + T[14];
        double py, px;
        undistort_pdn(a.cam, yc / zc, xc / zc, py, px);  // project_undistort, camera.jl:73-85 (always the left intrinsics, frame.jl:482-484)
        if (1.0 <= py && py <= a.bound_h && 1.0 <= px && px <= a.bound_w) {  // in_image / in_right_image
            flag = 1;
            dy = a.scale * (py - a.pix[2 * (size_t)i]);
            dx = a.scale * (px - a.pix[2 * (size_t)i + 1]);
        } else {
            flag = 2;  // neither tracked with the prior nor re-queued (map_manager.jl:489-506)
        }
    }
    a.flag[i] = flag;
    a.disp[2 * (size_t)i] = dy;
    a.disp[2 * (size_t)i + 1] = dx;
}

__global__ void k_match_update(const MatchArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    uint8_t st = a.status[i];
    double oy = nan, ox = nan, uy = nan, ux = nan, p0 = nan, p1 = nan, p2 = nan;
    if (st & 1) {
        const double ny = a.tracked[2 * (size_t)i], nx = a.tracked[2 * (size_t)i + 1];
        if (!a.stereo) {
            oy = ny; ox = nx;
            undistort_point(a.cam, oy, ox, uy, ux);
        } else {
            double ry, rx;
            undistort_point(a.rcam, ny, nx, ry, rx);
            if (fabs(a.undist[2 * (size_t)i] - ry) > a.epipolar) {
                st = (uint8_t)((st & ~1) | 16);
            } else {
                oy = a.pix[2 * (size_t)i]; ox = nx;  // same row as the left keypoint
                undistort_point(a.rcam, oy, ox, uy, ux);
            }
        }
        if (st & 1) {
            const MatchCam& c = a.stereo ? a.rcam : a.cam;
            p0 = (ux - c.cx) / c.fx; p1 = (uy - c.cy) / c.fy; p2 = 1.0;  // backproject, camera.jl:130-143: (x, y, 1)
        }
    }
    a.status[i] = st;
    a.out_pix[2 * (size_t)i] = oy; a.out_pix[2 * (size_t)i + 1] = ox;
    a.out_und[2 * (size_t)i] = uy; a.out_und[2 * (size_t)i + 1] = ux;
    a.out_pos[3 * (size_t)i] = p0; a.out_pos[3 * (size_t)i + 1] = p1; a.out_pos[3 * (size_t)i + 2] = p2;
}

}  // namespace

void launch_match_prior(cudaStream_t s, const MatchArgs& a) { k_match_prior<<<(a.n + 127) / 128, 128, 0, s>>>(a); }
void launch_match_update(cudaStream_t s, const MatchArgs& a) { k_match_update<<<(a.n + 127) / 128, 128, 0, s>>>(a); }

}  // namespace sk
