// Pyramidal Lucas-Kanade with forward-backward check (sm_100a).  One warp per keypoint runs every level of the
// forward pass and then the backward pass, so no host-side compaction (tracker.jl:30-48) is needed.
//
// Reference behaviour: lucas_kanade.jl:9-100 (optflow!), :140-212 (helpers), utils.jl:5-45 (2x2 pinv),
// tracker.jl:17-68 (fb_tracking!).
//
// Mapping: lane i owns window row i (<= 2w+1 <= 31 rows, plus one extra row for the bilinear tap), the template
// row (I, Iy, Ix) lives in registers, the target is read column by column (y is contiguous => every load is
// one coalesced segment), the vertical lerp partner comes from lane i+1 by shuffle and the horizontal one from
// the previous column.  The 2x2 solve, the position arithmetic and every decision (bounds, eigenvalue gate,
// epsilon stop) are in Float64 and warp-uniform.
#include "common.cuh"

namespace sk {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

struct Offs { int up, down, left, right; };

// get_offsets, lucas_kanade.jl:199-208
__device__ __forceinline__ Offs get_offsets(int py, int px, double ny, double nx, int window, int H, int W) {
    Offs o;
    o.up = (int)floor(fmin((double)window, fmin((double)py, ny) - 1.0));
    o.down = (int)floor(fmin((double)window, (double)H - fmax((double)py, ny)));
    o.left = (int)floor(fmin((double)window, fmin((double)px, nx) - 1.0));
    o.right = (int)floor(fmin((double)window, (double)W - fmax((double)px, nx)));
    return o;
}

// compute_spatial_gradient, lucas_kanade.jl:140-157 with the integral-image boxdiff replaced by a direct window
// sum of the smoothed planes.  Returns false when the window is degenerate.
__device__ __forceinline__ bool spatial_gradient(const float* __restrict__ fb, const LKLevel& L, int py, int px, Offs o, int lane,
                                                 double* Ginv, double& min_eig) {
    const int r0 = py - o.up, r1 = py + o.down, c0 = px - o.left, c1 = px + o.right;
    if (r1 < r0 || c1 < c0 || r0 < 1 || c0 < 1 || r1 > L.H || c1 > L.W) return false;
    const int nrows = r1 - r0 + 1, ncols = c1 - c0 + 1;
    float syy = 0.f, sxx = 0.f, syx = 0.f;
    if (lane < nrows) {
        const size_t base = (size_t)(r0 - 1 + lane) + (size_t)(c0 - 1) * L.pitch;
        const float* pyy = fb + L.oSyy + base;
        const float* pxx = fb + L.oSxx + base;
        const float* pyx = fb + L.oSyx + base;
        for (int k = 0; k < ncols; ++k) {
            syy += __ldg(pyy + (size_t)k * L.pitch);
            sxx += __ldg(pxx + (size_t)k * L.pitch);
            syx += __ldg(pyx + (size_t)k * L.pitch);
        }
    }
    const double a = warp_sum((double)syy), c = warp_sum((double)sxx), b = warp_sum((double)syx);
    // singular values of the symmetric G = [a b; b c] (utils.jl:5-27 with H = 0): Q +- R
    const double E = 0.5 * (a + c), F = 0.5 * (a - c);
    const double R = sqrt(F * F + b * b), Q = fabs(E);
    const double s1 = Q + R, s2 = fabs(Q - R);
    min_eig = fmin(s1, s2) / ((double)nrows * (double)ncols);
    const double tol = 1.4901161193847656e-08;  // sqrt(eps(Float64)), utils.jl:37
    if (s1 > tol && s2 > tol) {
        const double det = a * c - b * b;
        const double id = 1.0 / det;
        Ginv[0] = c * id; Ginv[1] = -b * id; Ginv[2] = -b * id; Ginv[3] = a * id;
    } else {
        // rank-deficient: Moore-Penrose through the eigen-decomposition (only reachable with eigenvalue_threshold ~ 0)
        Ginv[0] = Ginv[1] = Ginv[2] = Ginv[3] = 0.0;
        const double l1 = E + (E >= 0 ? R : -R);  // eigenvalue of largest magnitude
        if (fabs(l1) > tol) {
            double vx = b, vy = l1 - a;
            if (fabs(vx) + fabs(vy) < 1e-300) { vx = l1 - c; vy = b; }
            if (fabs(vx) + fabs(vy) < 1e-300) { vx = fabs(a) >= fabs(c) ? 1.0 : 0.0; vy = 1.0 - vx; }
            const double nn = 1.0 / (vx * vx + vy * vy);
            Ginv[0] = vx * vx * nn / l1; Ginv[1] = vx * vy * nn / l1; Ginv[2] = Ginv[1]; Ginv[3] = vy * vy * nn / l1;
        }
    }
    return true;
}

template <int W2>
__device__ __forceinline__ void load_template(const float* __restrict__ fb, const LKLevel& L, int py, int px, Offs o, int lane,
                                              float (&tI)[W2], float (&tIy)[W2], float (&tIx)[W2]) {
    const int nrows = o.up + o.down + 1, ncols = o.left + o.right + 1;
    const bool rowact = lane < nrows;
    const size_t base = (size_t)(py - o.up - 1 + (rowact ? lane : 0)) + (size_t)(px - o.left - 1) * L.pitch;
    const float* pI = fb + L.oI + base;
    const float* pIy = fb + L.oIy + base;
    const float* pIx = fb + L.oIx + base;
#pragma unroll
    for (int k = 0; k < W2; ++k) {
        const bool ok = rowact && k < ncols;
        tI[k] = ok ? __ldg(pI + (size_t)k * L.pitch) : 0.f;
        tIy[k] = ok ? __ldg(pIy + (size_t)k * L.pitch) : 0.f;
        tIx[k] = ok ? __ldg(pIx + (size_t)k * L.pitch) : 0.f;
    }
}

// optflow! for one point (lucas_kanade.jl:24-97).  (dy, dx) in/out at the scale of level `levels`.
template <int W2>
__device__ bool lk_track(const float* __restrict__ fbA, const float* __restrict__ fbB, const LKArgs& a, double pty, double ptx,
                         double& dy, double& dx, int levels, double eps, int lane, unsigned long long& wpx, unsigned int& nit) {
    for (int lvl = levels; lvl >= 0; --lvl) {
        const LKLevel& L = a.lv[lvl];
        const double inv = 1.0 / (double)(1 << lvl);
        const int py = (int)floor(pty * inv), px = (int)floor(ptx * inv);  // get_pyramid_coordinate :197
        Offs o = get_offsets(py, px, (double)py, (double)px, a.window, L.H, L.W);
        double Ginv[4], min_eig;
        if (!spatial_gradient(fbA, L, py, px, o, lane, Ginv, min_eig)) return false;
        if (min_eig < a.eig_thr) return false;
        float tI[W2], tIy[W2], tIx[W2];
        load_template<W2>(fbA, L, py, px, o, lane, tI, tIy, tIx);
        const float* __restrict__ T = fbB + L.oI;
        double cy = 0.0, cx = 0.0;
        for (int it = 0; it < a.iterations; ++it) {
            const double pcy = (double)py + (dy + cy), pcx = (double)px + (dx + cx);
            if (!(1.0 <= pcy && pcy <= (double)L.H && 1.0 <= pcx && pcx <= (double)L.W)) return false;
            const Offs no = get_offsets(py, px, pcy, pcx, a.window, L.H, L.W);
            if (no.up != o.up || no.down != o.down || no.left != o.left || no.right != o.right) {
                o = no;
                if (!spatial_gradient(fbA, L, py, px, o, lane, Ginv, min_eig)) return false;
                if (min_eig < a.eig_thr) return false;
                load_template<W2>(fbA, L, py, px, o, lane, tI, tIy, tIx);
            }
            const int nrows = o.up + o.down + 1, ncols = o.left + o.right + 1;
            // prepare_linear_system (lucas_kanade.jl:159-173): bilinear weights are identical for the whole window
            const double fy = floor(pcy), fx = floor(pcx);
            const float wy = (float)(pcy - fy), wx = (float)(pcx - fx);
            const int iy0 = (int)fy - o.up, ix0 = (int)fx - o.left;  // 1-based top-left tap
            const int ry = min(iy0 + min(lane, nrows), L.H);          // clamped tap row carries weight 0
            const float* trow = T + (ry - 1);
            float by = 0.f, bx = 0.f;
            float t = __ldg(trow + (size_t)(min(ix0, L.W) - 1) * L.pitch);
            float tn = __shfl_down_sync(FULL, t, 1);
            float prev = fmaf(wy, tn - t, t);
#pragma unroll
            for (int k = 0; k < W2; ++k) {
                if (k < ncols) {
                    t = __ldg(trow + (size_t)(min(ix0 + k + 1, L.W) - 1) * L.pitch);
                    tn = __shfl_down_sync(FULL, t, 1);
                    const float cur = fmaf(wy, tn - t, t);
                    const float val = fmaf(wx, cur - prev, prev);
                    const float dI = tI[k] - val;
                    by = fmaf(dI, tIy[k], by);
                    bx = fmaf(dI, tIx[k], bx);
                    prev = cur;
                }
            }
            const double sby = warp_sum((double)by), sbx = warp_sum((double)bx);
            wpx += (unsigned long long)(nrows * ncols);
            nit += 1;
            const double ffy = Ginv[0] * sby + Ginv[1] * sbx, ffx = Ginv[2] * sby + Ginv[3] * sbx;
            if (fabs(ffy) < eps && fabs(ffx) < eps) break;
            cy += ffy; cx += ffx;
            const double qy = pcy + ffy, qx = pcx + ffx;
            if (!(1.0 <= qy && qy <= (double)L.H && 1.0 <= qx && qx <= (double)L.W)) return false;
        }
        dy += cy; dx += cx;
        if (lvl > 0) { dy *= 2.0; dx *= 2.0; }
    }
    return true;
}

template <int W2>
__global__ void __launch_bounds__(128) k_lk(LKArgs a) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int total = a.n_frames * a.n_per_frame;
    if (gw >= total) return;
    const int f = gw / a.n_per_frame;
    const float* fbA = a.A.frame(a.offA + f);
    const float* fbB = a.B.frame(a.offB + f);
    const double pty = a.pts[2 * (size_t)gw], ptx = a.pts[2 * (size_t)gw + 1];
    double dy = 0.0, dx = 0.0;
    if (a.disp_in) { dy = a.disp_in[2 * (size_t)gw]; dx = a.disp_in[2 * (size_t)gw + 1]; }
    unsigned long long wpx = 0; unsigned int nit = 0;
    bool ok = lk_track<W2>(fbA, fbB, a, pty, ptx, dy, dx, a.levels, a.eps, lane, wpx, nit);
    uint8_t st = ok ? 1 : 0;
    if (a.mode == 0) {
        // optflow!: a failed point keeps the displacement it had when it failed (stale value); we write it back as is
        if (lane == 0) {
            if (a.disp_out) { a.disp_out[2 * (size_t)gw] = dy; a.disp_out[2 * (size_t)gw + 1] = dx; }
            a.status[gw] = st;
        }
    } else {
        uint8_t out = 0;
        if (ok) {
            const double ny = pty + dy, nx = ptx + dx;  // tracker.jl:41-43
            if (lane == 0 && a.out_pts) { a.out_pts[2 * (size_t)gw] = ny; a.out_pts[2 * (size_t)gw + 1] = nx; }
            out = 2;
            double bdy = -dy, bdx = -dx;  // back_pyramid_levels = 0 => scale 1 (tracker.jl:34-46)
            // backward pass: LucasKanade default epsilon (tracker.jl:51-54 does not forward it)
            bool okb = lk_track<W2>(fbB, fbA, a, ny, nx, bdy, bdx, 0, 1e-2, lane, wpx, nit);
            if (okb) {
                const double by = ny + bdy, bx = nx + bdx;
                const double ey = pty - by, ex = ptx - bx;
                if (!(sqrt(ey * ey + ex * ex) >= a.max_dist)) out = 3;
            }
        }
        else if (lane == 0 && a.out_pts) {  // the reference leaves these entries undefined; make them recognisable
            a.out_pts[2 * (size_t)gw] = __longlong_as_double(0x7ff8000000000000LL);
            a.out_pts[2 * (size_t)gw + 1] = __longlong_as_double(0x7ff8000000000000LL);
        }
        if (lane == 0) a.status[gw] = out;
    }
    if (lane == 0 && a.counters) {
        atomicAdd(a.counters, wpx);
        atomicAdd(a.counters + 1, (unsigned long long)nit);
    }
}

int launch_lk(cudaStream_t s, const LKArgs& a, const Hook* hk) {
    const int total = a.n_frames * a.n_per_frame;
    if (total <= 0) return 0;
    mark(hk, a.mode ? "k_lk_fb" : "k_lk_optflow");
    const int wpb = 4;
    const int blocks = (total + wpb - 1) / wpb;
    const int w2 = 2 * a.window + 1;
    if (w2 <= 19) k_lk<19><<<blocks, wpb * 32, 0, s>>>(a);
    else if (w2 <= 23) k_lk<23><<<blocks, wpb * 32, 0, s>>>(a);
    else k_lk<31><<<blocks, wpb * 32, 0, s>>>(a);
    return 1;
}

}  // namespace sk
