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

// triangulate_stereo! (mapper.jl:142-183), one thread per stereo keypoint.  DLT system of RecoverPose.triangulate [3P]:
// rows x * P[3,:] - P[1,:], y * P[3,:] - P[2,:] for both views with P1 = to_4x4(K) and P2 = to_4x4(K_right) * Ti0 (pixel units,
// points in (x, y) order, mapper.jl:151-152,162-164); the homogeneous point is the eigenvector of A'A for the smallest
// eigenvalue.  The reference gets it from LAPACK's geev; here a cyclic Jacobi iteration on the symmetric 4 x 4 matrix (the
// eigenvector is defined up to scale, and the point is normalised by its 4th component right away, mapper.jl:165).
__device__ void jacobi4_smallest(double M[4][4], double v[4]) {
    double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    for (int sweep = 0; sweep < 24; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int p = 0; p < 4; ++p) {
            diag += M[p][p] * M[p][p];
            for (int q = p + 1; q < 4; ++q) off += M[p][q] * M[p][q];
        }
        if (off <= 1e-40 * diag || off == 0.0) break;
        for (int p = 0; p < 3; ++p)
            for (int q = p + 1; q < 4; ++q) {
                const double apq = M[p][q];
                if (apq == 0.0) continue;
                const double theta = (M[q][q] - M[p][p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 4; ++k) {  // columns p, q
                    const double mkp = M[k][p], mkq = M[k][q];
                    M[k][p] = c * mkp - sn * mkq; M[k][q] = sn * mkp + c * mkq;
                }
                for (int k = 0; k < 4; ++k) {  // rows p, q
                    const double mpk = M[p][k], mqk = M[q][k];
                    M[p][k] = c * mpk - sn * mqk; M[q][k] = sn * mpk + c * mqk;
                }
                for (int k = 0; k < 4; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - sn * vkq; V[k][q] = sn * vkp + c * vkq;
                }
            }
    }
    int best = 0;
    for (int k = 1; k < 4; ++k)
        if (M[k][k] < M[best][best]) best = k;
    for (int k = 0; k < 4; ++k) v[k] = V[k][best];
}

__global__ void k_triangulate_stereo(const TriArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    // P1 = to_4x4(K), P2 = to_4x4(K_right) * Ti0, row-major here
    double P1[3][4] = {{a.cam.fx, 0, a.cam.cx, 0}, {0, a.cam.fy, a.cam.cy, 0}, {0, 0, 1, 0}};
    double P2[3][4];
    for (int c = 0; c < 4; ++c) {
        const double t0 = a.Ti0[0 + 4 * c], t1 = a.Ti0[1 + 4 * c], t2 = a.Ti0[2 + 4 * c];
        P2[0][c] = a.rcam.fx * t0 + a.rcam.cx * t2;
        P2[1][c] = a.rcam.fy * t1 + a.rcam.cy * t2;
        P2[2][c] = t2;
    }
    const double y1 = a.und[2 * (size_t)i], x1 = a.und[2 * (size_t)i + 1];
    const double y2 = a.rund[2 * (size_t)i], x2 = a.rund[2 * (size_t)i + 1];
    double A[4][4];
    for (int c = 0; c < 4; ++c) {
        A[0][c] = x1 * P1[2][c] - P1[0][c];
        A[1][c] = y1 * P1[2][c] - P1[1][c];
        A[2][c] = x2 * P2[2][c] - P2[0][c];
        A[3][c] = y2 * P2[2][c] - P2[1][c];
    }
    double M[4][4];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) M[r][c] = A[0][r] * A[0][c] + A[1][r] * A[1][c] + A[2][r] * A[2][c] + A[3][r] * A[3][c];
    double v[4];
    jacobi4_smallest(M, v);
    const double iw = 1.0 / v[3];
    const double X = v[0] * iw, Y = v[1] * iw, Z = v[2] * iw;  // left camera coordinates (w = 1)
    uint8_t st = 1;
    double w0 = nan, w1 = nan, w2 = nan;
    if (!(Z >= 0.1)) st = 2;  // left_point[3] < 0.1 (a NaN point is dropped as well)
    else {
        const double rx = a.Ti0[0] * X + a.Ti0[4] * Y + a.Ti0[8] * Z + a.Ti0[12];
        const double ry = a.Ti0[1] * X + a.Ti0[5] * Y + a.Ti0[9] * Z + a.Ti0[13];
        const double rz = a.Ti0[2] * X + a.Ti0[6] * Y + a.Ti0[10] * Z + a.Ti0[14];
        if (!(rz >= 0.1)) st = 3;
        else {
            // project (camera.jl:62-67): (fy * y / z + cy, fx * x / z + cx)
            const double iz = 1.0 / Z;
            const double ly = a.cam.fy * Y * iz + a.cam.cy, lx = a.cam.fx * X * iz + a.cam.cx;
            const double ey = y1 - ly, ex = x1 - lx;
            if (sqrt(ey * ey + ex * ex) > a.max_error) st = 4;
            else {
                const double irz = 1.0 / rz;
                const double qy = a.rcam.fy * ry * irz + a.rcam.cy, qx = a.rcam.fx * rx * irz + a.rcam.cx;
                const double fy_ = y2 - qy, fx_ = x2 - qx;
                if (sqrt(fy_ * fy_ + fx_ * fx_) > a.max_error) st = 5;
                else {
                    // project_camera_to_world (frame.jl:452-456): wc * (X, Y, Z, 1)
                    w0 = a.wc[0] * X + a.wc[4] * Y + a.wc[8] * Z + a.wc[12];
                    w1 = a.wc[1] * X + a.wc[5] * Y + a.wc[9] * Z + a.wc[13];
                    w2 = a.wc[2] * X + a.wc[6] * Y + a.wc[10] * Z + a.wc[14];
                }
            }
        }
    }
    a.status[i] = st;
    a.world[3 * (size_t)i] = w0; a.world[3 * (size_t)i + 1] = w1; a.world[3 * (size_t)i + 2] = w2;
}

}  // namespace

void launch_triangulate_stereo(cudaStream_t s, const TriArgs& a) { k_triangulate_stereo<<<(a.n + 127) / 128, 128, 0, s>>>(a); }
void launch_match_prior(cudaStream_t s, const MatchArgs& a) { k_match_prior<<<(a.n + 127) / 128, 128, 0, s>>>(a); }
void launch_match_update(cudaStream_t s, const MatchArgs& a) { k_match_update<<<(a.n + 127) / 128, 128, 0, s>>>(a); }

}  // namespace sk
