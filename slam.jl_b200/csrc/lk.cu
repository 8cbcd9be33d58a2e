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
//
// Code-size discipline: the forward levels and the backward pass run through ONE copy of the level body (a stage
// loop), otherwise the fully unrolled window loops overflow the instruction cache (measured: 8.7k SASS
// instructions and 20% "no instruction" stalls in the first version).
// Planes carry one zeroed guard row and guard column (pitch >= H+1, W+1 columns allocated), so the bilinear tap
// that falls on H+1 / W+1 with weight exactly 0 needs no clamp.
#include <cstdlib>

#include "common.cuh"

namespace sk {

template <typename T>
__device__ __forceinline__ const T* col_ptr(const T* base, unsigned stride_bytes, unsigned k) {
    // base + k columns: one IMAD.WIDE.U32 (u32 x u32 + u64) instead of a 64-bit multiply-shift-add chain
    return reinterpret_cast<const T*>(reinterpret_cast<const char*>(base) + (unsigned long long)stride_bytes * k);
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

#ifndef LK_MIN_BLOCKS
#define LK_MIN_BLOCKS 4
#endif
template <int W2>
__global__ void __launch_bounds__(128, (W2 <= 19 ? LK_MIN_BLOCKS : (W2 <= 23 ? 3 : 2))) k_lk(const LKArgs a) {
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

    const int w = a.window;
    const int nstage = a.levels + 1 + (a.mode ? 1 : 0);
    unsigned int wpx = 0, nit = 0;
    double qy = pty, qx = ptx;  // point tracked by the current pass
    bool ok = true;             // status of the current pass
    uint8_t result = 0;

    float tI[W2], tIy[W2], tIx[W2];

    for (int s = 0; s < nstage; ++s) {
        const bool back = s > a.levels;
        const int lvl = back ? 0 : a.levels - s;
        if (back) {
            // forward pass succeeded (otherwise we left the loop): tracker.jl:37-46
            qy = pty + dy; qx = ptx + dx;
            if (lane == 0 && a.out_pts) { a.out_pts[2 * (size_t)gw] = qy; a.out_pts[2 * (size_t)gw + 1] = qx; }
            result = 2;
            dy = -dy; dx = -dx;  // back_pyramid_levels = 0 => scale 1
            const float* t = fbA; fbA = fbB; fbB = t;
        }
        const double eps = back ? 1e-2 : a.eps;  // tracker.jl:51-54 does not forward epsilon to the backward pass
        const LKLevel& L = a.lv[lvl];
        const int H = L.H, W = L.W, pitch = L.pitch;
        const unsigned pitch4 = (unsigned)L.pitch * 4u;
        const double inv = 1.0 / (double)(1 << lvl);
        const int py = (int)floor(qy * inv), px = (int)floor(qx * inv);  // get_pyramid_coordinate, lucas_kanade.jl:197
        // get_offsets(point, point), lucas_kanade.jl:199-208, in integer arithmetic
        int up = min(w, py - 1), down = min(w, H - py), left = min(w, px - 1), right = min(w, W - px);
        bool setup = true;
        double g00 = 0, g01 = 0, g11 = 0;  // G^-1 (symmetric)
        double cy = 0.0, cx = 0.0;
        int it = 0;
        while (true) {
            const int nrows = up + down + 1, ncols = left + right + 1;
            if (setup) {
                // ---- compute_spatial_gradient (lucas_kanade.jl:140-157): window sums of the smoothed planes
                const int r0 = py - up, c0 = px - left;
                if (nrows < 1 || ncols < 1 || r0 < 1 || c0 < 1 || py + down > H || px + right > W) { ok = false; break; }
                const bool rowact = lane < nrows;
                const int row = r0 - 1 + min(lane, nrows - 1);  // 0-based
                const float* colA = fbA + (size_t)row;
                // window sums from the exclusive row prefix planes: R[., c1] - R[., c0-1]  (0-based prefix columns)
                float syy, sxx, syx;
                {
                    const size_t lo = (size_t)(c0 - 1) * pitch, hi = (size_t)(px + right) * pitch;
                    syy = __ldg(colA + L.oRyy + hi) - __ldg(colA + L.oRyy + lo);
                    sxx = __ldg(colA + L.oRxx + hi) - __ldg(colA + L.oRxx + lo);
                    syx = __ldg(colA + L.oRyx + hi) - __ldg(colA + L.oRyx + lo);
                    if (!rowact) { syy = 0.f; sxx = 0.f; syx = 0.f; }
                }
                // ---- template rows into registers (issued before the reduction so the loads overlap it)
                {
                    const float* pI = colA + L.oI + (size_t)(c0 - 1) * pitch;
                    const float2* pG = reinterpret_cast<const float2*>(fbA + L.oG) + (size_t)row + (size_t)(c0 - 1) * pitch;
#pragma unroll
                    for (int k = 0; k < W2; ++k) {
                        const bool okk = rowact && k < ncols;
                        float2 g2 = make_float2(0.f, 0.f);
                        float iv = 0.f;
                        if (okk) {
                            iv = __ldg(col_ptr(pI, pitch4, k));
                            g2 = __ldg(col_ptr(pG, 2u * pitch4, k));
                        }
                        tI[k] = iv; tIy[k] = g2.x; tIx[k] = g2.y;
                    }
                }
                const double ga = (double)warp_sum_f(syy), gc = (double)warp_sum_f(sxx), gb = (double)warp_sum_f(syx);
                // singular values of the symmetric G = [a b; b c] (utils.jl:5-27 with H = 0): Q +- R
                const double E = 0.5 * (ga + gc), F = 0.5 * (ga - gc);
                const double R = sqrt(F * F + gb * gb), Q = fabs(E);
                const double s1 = Q + R, s2 = fabs(Q - R);
                const double min_eig = fmin(s1, s2) / (double)(nrows * ncols);
                if (min_eig < a.eig_thr) { ok = false; break; }
                const double tol = 1.4901161193847656e-08;  // sqrt(eps(Float64)), utils.jl:37
                if (s2 > tol) {
                    const double id = 1.0 / (ga * gc - gb * gb);
                    g00 = gc * id; g01 = -gb * id; g11 = ga * id;
                } else {
                    // rank-deficient (only reachable with eigenvalue_threshold ~ 0): Moore-Penrose via the eigenvectors
                    g00 = g01 = g11 = 0.0;
                    const double l1 = E + (E >= 0 ? R : -R);
                    if (fabs(l1) > tol) {
                        double vx = gb, vy = l1 - ga;
                        if (fabs(vx) + fabs(vy) < 1e-300) { vx = l1 - gc; vy = gb; }
                        if (fabs(vx) + fabs(vy) < 1e-300) { vx = fabs(ga) >= fabs(gc) ? 1.0 : 0.0; vy = 1.0 - vx; }
                        const double nn = 1.0 / ((vx * vx + vy * vy) * l1);
                        g00 = vx * vx * nn; g01 = vx * vy * nn; g11 = vy * vy * nn;
                    }
                }
                setup = false;
            }
            if (it >= a.iterations) break;
            const double pcy = (double)py + (dy + cy), pcx = (double)px + (dx + cx);
            // floor / ceil as integers serve both lies_in (1 <= pc <= size  <=>  floor >= 1 && ceil <= size) and get_offsets:
            // floor(min(w, min(p, pc) - 1)) = min(w, min(p, floor pc) - 1), floor(min(w, H - max(p, pc))) = min(w, H - max(p, ceil pc))
            const int fy = __double2int_rd(pcy), fx = __double2int_rd(pcx);
            const int cyi = __double2int_ru(pcy), cxi = __double2int_ru(pcx);
            if (!(fy >= 1 && cyi <= H && fx >= 1 && cxi <= W)) { ok = false; break; }
            const int nup = min(w, min(py, fy) - 1), ndown = min(w, H - max(py, cyi));
            const int nleft = min(w, min(px, fx) - 1), nright = min(w, W - max(px, cxi));
            if (nup != up || ndown != down || nleft != left || nright != right) {
                up = nup; down = ndown; left = nleft; right = nright;
                setup = true;  // recompute G and reload the template for the new grid (lucas_kanade.jl:55-66)
                continue;
            }
            // ---- prepare_linear_system (lucas_kanade.jl:159-173); the bilinear weights are the same for the whole window
            const float wy = (float)(pcy - (double)fy), wx = (float)(pcx - (double)fx);
            const float* tp = fbB + L.oI + (size_t)(fy - up - 1 + min(lane, nrows)) + (size_t)(fx - left - 1) * pitch;
            float tv[W2 + 1];
            // unpredicated: columns past the window are finite junk (guard column, next plane or allocation slack) and
            // meet zero template gradients below
#pragma unroll
            for (int k = 0; k <= W2; ++k) tv[k] = __ldg(col_ptr(tp, pitch4, k));
            float by = 0.f, bx = 0.f;
            float tn = __shfl_down_sync(FULL, tv[0], 1);
            float prev = fmaf(wy, tn - tv[0], tv[0]);
#pragma unroll
            for (int k = 0; k < W2; ++k) {
                tn = __shfl_down_sync(FULL, tv[k + 1], 1);
                const float cur = fmaf(wy, tn - tv[k + 1], tv[k + 1]);
                const float val = fmaf(wx, cur - prev, prev);
                const float dI = tI[k] - val;
                by = fmaf(dI, tIy[k], by);  // template gradients are 0 outside the window => no predicate needed
                bx = fmaf(dI, tIx[k], bx);
                prev = cur;
            }
            const double sby = (double)warp_sum_f(by), sbx = (double)warp_sum_f(bx);
            wpx += (unsigned)(nrows * ncols);
            nit += 1;
            ++it;
            const double ffy = g00 * sby + g01 * sbx, ffx = g01 * sby + g11 * sbx;
            if (fabs(ffy) < eps && fabs(ffx) < eps) break;
            cy += ffy; cx += ffx;
            const double ny = pcy + ffy, nx = pcx + ffx;
            if (!(__double2int_rd(ny) >= 1 && __double2int_ru(ny) <= H && __double2int_rd(nx) >= 1 && __double2int_ru(nx) <= W)) { ok = false; break; }
        }
        if (!ok) break;
        dy += cy; dx += cx;
        if (lvl > 0) { dy *= 2.0; dx *= 2.0; }
    }

    if (a.mode == 0) {
        // optflow!: a failed point keeps the displacement it had when it failed (stale value, lucas_kanade.jl:94-95)
        if (lane == 0) {
            if (a.disp_out) { a.disp_out[2 * (size_t)gw] = dy; a.disp_out[2 * (size_t)gw + 1] = dx; }
            a.status[gw] = ok ? 1 : 0;
        }
    } else if (lane == 0) {
        if (result == 0) {
            // forward pass failed; the reference leaves new_keypoints[i] undefined, make it recognisable
            if (a.out_pts) {
                a.out_pts[2 * (size_t)gw] = __longlong_as_double(0x7ff8000000000000LL);
                a.out_pts[2 * (size_t)gw + 1] = __longlong_as_double(0x7ff8000000000000LL);
            }
        } else if (ok) {
            // tracker.jl:59-66: back-tracked point must land within max_distance of the original keypoint
            const double by = qy + dy, bx = qx + dx;
            const double ey = pty - by, ex = ptx - bx;
            if (!(sqrt(ey * ey + ex * ex) >= a.max_dist)) result = 3;
        }
        a.status[gw] = result;
    }
    if (lane == 0 && a.counters) {
        atomicAdd(a.counters, (unsigned long long)wpx);
        atomicAdd(a.counters + 1, (unsigned long long)nit);
    }
}

// Any window size (2w + 1 > 31 in the product; every size with SLAMKLT_LK_VARIANT=a).  Same control flow, same Float64 decisions
// and the same fp32 expressions as k_lk, without the register-resident template: lane l owns window rows l, l + 32, ... and reads
// template, gradients and the four bilinear taps of every pixel from global memory (L1 / L2 resident: a window is a few KB).  For
// windows of at most 32 rows every lane owns one row and sums it in the same order as k_lk, so both kernels give identical bits.
__global__ void __launch_bounds__(128) k_lk_any(const LKArgs a) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int total = a.n_frames * a.n_per_frame;
    if (gw >= total) return;
    const int f = gw / a.n_per_frame;
    const float* fbA = a.A.frame(a.offA + f);
    const float* fbB = a.B.frame(a.offB + f);
    const double pty = a.pts[2 * (size_t)gw], ptx = a.pts[2 * (size_t)gw + 1];
    // mode 2 = optical_flow_matching! (map_manager.jl:451-564): keypoints with a prior (3-D keypoints) are first tracked with that
    // prior on levels3d levels; those that fail, and all others, are tracked from a zero displacement on all levels
    const uint8_t prior_flag = (a.mode == 2 && a.has_prior) ? a.has_prior[gw] : (uint8_t)0;
    if (prior_flag == 2) {  // projection outside the image: not tracked at all (map_manager.jl:489-506)
        if (lane == 0) {
            a.status[gw] = 8;
            if (a.out_pts) {
                a.out_pts[2 * (size_t)gw] = __longlong_as_double(0x7ff8000000000000LL);
                a.out_pts[2 * (size_t)gw + 1] = __longlong_as_double(0x7ff8000000000000LL);
            }
        }
        return;
    }
    const bool prior_first = a.mode == 2 && prior_flag != 0;
    bool second_try = false;
    int levels_cur = prior_first ? a.levels3d : a.levels;
    const float* const fbA0 = fbA;
    const float* const fbB0 = fbB;
    double dy = 0.0, dx = 0.0;
    if (a.disp_in && !(a.mode == 2 && !prior_first)) { dy = a.disp_in[2 * (size_t)gw]; dx = a.disp_in[2 * (size_t)gw + 1]; }

    const int w = a.window;
    unsigned int wpx = 0, nit = 0;
    double qy = pty, qx = ptx;
    bool ok = true;
    uint8_t result = 0;

retry:
    const int nstage = levels_cur + 1 + (a.mode ? 1 : 0);
    for (int s = 0; s < nstage; ++s) {
        const bool back = s > levels_cur;
        const int lvl = back ? 0 : levels_cur - s;
        if (back) {  // tracker.jl:37-46
            qy = pty + dy; qx = ptx + dx;
            if (lane == 0 && a.out_pts) { a.out_pts[2 * (size_t)gw] = qy; a.out_pts[2 * (size_t)gw + 1] = qx; }
            result = 2;
            dy = -dy; dx = -dx;
            const float* t = fbA; fbA = fbB; fbB = t;
        }
        const double eps = back ? 1e-2 : a.eps;
        const LKLevel& L = a.lv[lvl];
        const int H = L.H, W = L.W;
        const size_t pitch = (size_t)L.pitch;
        const double inv = 1.0 / (double)(1 << lvl);
        const int py = (int)floor(qy * inv), px = (int)floor(qx * inv);
        int up = min(w, py - 1), down = min(w, H - py), left = min(w, px - 1), right = min(w, W - px);
        bool setup = true;
        double g00 = 0, g01 = 0, g11 = 0;
        double cy = 0.0, cx = 0.0;
        int it = 0;
        while (true) {
            const int nrows = up + down + 1, ncols = left + right + 1;
            const int r0 = py - up, c0 = px - left;
            if (setup) {
                if (nrows < 1 || ncols < 1 || r0 < 1 || c0 < 1 || py + down > H || px + right > W) { ok = false; break; }
                float syy = 0.f, sxx = 0.f, syx = 0.f;
                const size_t lo = (size_t)(c0 - 1) * pitch, hi = (size_t)(px + right) * pitch;
                for (int r = lane; r < nrows; r += 32) {
                    const float* colA = fbA + (size_t)(r0 - 1 + r);
                    syy += __ldg(colA + L.oRyy + hi) - __ldg(colA + L.oRyy + lo);
                    sxx += __ldg(colA + L.oRxx + hi) - __ldg(colA + L.oRxx + lo);
                    syx += __ldg(colA + L.oRyx + hi) - __ldg(colA + L.oRyx + lo);
                }
                const double ga = (double)warp_sum_f(syy), gc = (double)warp_sum_f(sxx), gb = (double)warp_sum_f(syx);
                const double E = 0.5 * (ga + gc), F = 0.5 * (ga - gc);
                const double R = sqrt(F * F + gb * gb), Q = fabs(E);
                const double s1 = Q + R, s2 = fabs(Q - R);
                const double min_eig = fmin(s1, s2) / (double)(nrows * ncols);
                if (min_eig < a.eig_thr) { ok = false; break; }
                const double tol = 1.4901161193847656e-08;
                if (s2 > tol) {
                    const double id = 1.0 / (ga * gc - gb * gb);
                    g00 = gc * id; g01 = -gb * id; g11 = ga * id;
                } else {
                    g00 = g01 = g11 = 0.0;
                    const double l1 = E + (E >= 0 ? R : -R);
                    if (fabs(l1) > tol) {
                        double vx = gb, vy = l1 - ga;
                        if (fabs(vx) + fabs(vy) < 1e-300) { vx = l1 - gc; vy = gb; }
                        if (fabs(vx) + fabs(vy) < 1e-300) { vx = fabs(ga) >= fabs(gc) ? 1.0 : 0.0; vy = 1.0 - vx; }
                        const double nn = 1.0 / ((vx * vx + vy * vy) * l1);
                        g00 = vx * vx * nn; g01 = vx * vy * nn; g11 = vy * vy * nn;
                    }
                }
                setup = false;
            }
            if (it >= a.iterations) break;
            const double pcy = (double)py + (dy + cy), pcx = (double)px + (dx + cx);
            const int fy = __double2int_rd(pcy), fx = __double2int_rd(pcx);
            const int cyi = __double2int_ru(pcy), cxi = __double2int_ru(pcx);
            if (!(fy >= 1 && cyi <= H && fx >= 1 && cxi <= W)) { ok = false; break; }
            const int nup = min(w, min(py, fy) - 1), ndown = min(w, H - max(py, cyi));
            const int nleft = min(w, min(px, fx) - 1), nright = min(w, W - max(px, cxi));
            if (nup != up || ndown != down || nleft != left || nright != right) {
                up = nup; down = ndown; left = nleft; right = nright;
                setup = true;
                continue;
            }
            // prepare_linear_system (lucas_kanade.jl:159-173): the same fmaf forms as k_lk (vertical lerp, horizontal lerp, difference)
            const float wy = (float)(pcy - (double)fy), wx = (float)(pcx - (double)fx);
            float by = 0.f, bx = 0.f;
            for (int r = lane; r < nrows; r += 32) {
                const float* pI = fbA + L.oI + (size_t)(r0 - 1 + r) + (size_t)(c0 - 1) * pitch;
                const float2* pG = reinterpret_cast<const float2*>(fbA + L.oG) + (size_t)(r0 - 1 + r) + (size_t)(c0 - 1) * pitch;
                const float* tp = fbB + L.oI + (size_t)(fy - up - 1 + r) + (size_t)(fx - left - 1) * pitch;
                float t0 = __ldg(tp), t1 = __ldg(tp + 1);
                float prev = fmaf(wy, t1 - t0, t0);
                for (int k = 0; k < ncols; ++k) {
                    const float* tq = tp + (size_t)(k + 1) * pitch;
                    t0 = __ldg(tq); t1 = __ldg(tq + 1);
                    const float cur = fmaf(wy, t1 - t0, t0);
                    const float val = fmaf(wx, cur - prev, prev);
                    const float dI = __ldg(pI + (size_t)k * pitch) - val;
                    const float2 g2 = __ldg(pG + (size_t)k * pitch);
                    by = fmaf(dI, g2.x, by);
                    bx = fmaf(dI, g2.y, bx);
                    prev = cur;
                }
            }
            const double sby = (double)warp_sum_f(by), sbx = (double)warp_sum_f(bx);
            wpx += (unsigned)(nrows * ncols);
            nit += 1;
            ++it;
            const double ffy = g00 * sby + g01 * sbx, ffx = g01 * sby + g11 * sbx;
            if (fabs(ffy) < eps && fabs(ffx) < eps) break;
            cy += ffy; cx += ffx;
            const double ny = pcy + ffy, nx = pcx + ffx;
            if (!(__double2int_rd(ny) >= 1 && __double2int_ru(ny) <= H && __double2int_rd(nx) >= 1 && __double2int_ru(nx) <= W)) { ok = false; break; }
        }
        if (!ok) break;
        dy += cy; dx += cx;
        if (lvl > 0) { dy *= 2.0; dx *= 2.0; }
    }

    if (a.mode != 0 && result != 0 && ok) {  // tracker.jl:59-66
        const double by = qy + dy, bx = qx + dx;
        const double ey = pty - by, ex = ptx - bx;
        if (!(sqrt(ey * ey + ex * ex) >= a.max_dist)) result = 3;
    }
    if (prior_first && !second_try && result != 3) {
        // the prior pass failed: map_manager.jl:531-536 re-queues the keypoint with the 2-D ones (no prior, all levels)
        second_try = true;
        levels_cur = a.levels;
        dy = 0.0; dx = 0.0; qy = pty; qx = ptx;
        ok = true; result = 0;
        fbA = fbA0; fbB = fbB0;
        goto retry;
    }
    if (a.mode == 0) {
        if (lane == 0) {
            if (a.disp_out) { a.disp_out[2 * (size_t)gw] = dy; a.disp_out[2 * (size_t)gw + 1] = dx; }
            a.status[gw] = ok ? 1 : 0;
        }
    } else if (lane == 0) {
        if (result == 0 && a.out_pts) {
            a.out_pts[2 * (size_t)gw] = __longlong_as_double(0x7ff8000000000000LL);
            a.out_pts[2 * (size_t)gw + 1] = __longlong_as_double(0x7ff8000000000000LL);
        }
        a.status[gw] = result | ((prior_first && !second_try && result == 3) ? 4 : 0);  // bit2: tracked by the prior pass
    }
    if (lane == 0 && a.counters) {
        atomicAdd(a.counters, (unsigned long long)wpx);
        atomicAdd(a.counters + 1, (unsigned long long)nit);
    }
}

int launch_lk(cudaStream_t s, const LKArgs& a, const Hook* hk) {
    const int total = a.n_frames * a.n_per_frame;
    if (total <= 0) return 0;
    mark(hk, a.mode ? "k_lk_fb" : "k_lk_optflow");
    static const int wpb_env = getenv("SLAMKLT_LK_WPB") ? atoi(getenv("SLAMKLT_LK_WPB")) : 0;  // experiment knob
    const int wpb = (wpb_env >= 1 && wpb_env <= 4) ? wpb_env : 4;
    const int blocks = (total + wpb - 1) / wpb;
    const int w2 = 2 * a.window + 1;
    // A/B knob: 'r' = row-per-lane kernel (lk.cu), 'p' = cp.async patch kernel (lk_patch.cu), 'a' = any-window kernel; default = TMA-staged patch kernel
    // (lk_tma.cu) where it covers the window, then the cp.async patch kernel, then the row kernel
    const char* venv = getenv("SLAMKLT_LK_VARIANT");
    const char variant = venv ? venv[0] : 't';
    if (variant == 'a' || w2 > 31 || (a.mode == 2 && w2 > 23)) { k_lk_any<<<blocks, wpb * 32, 0, s>>>(a); return 1; }
    if (variant == 't' && launch_lk_tma(s, a)) return a.gtab ? 2 : 1;
    if (variant != 'r' && launch_lk_patch(s, a)) return 1;
    if (w2 <= 19) k_lk<19><<<blocks, wpb * 32, 0, s>>>(a);
    else if (w2 <= 23) k_lk<23><<<blocks, wpb * 32, 0, s>>>(a);
    else k_lk<31><<<blocks, wpb * 32, 0, s>>>(a);
    return 1;
}

}  // namespace sk
