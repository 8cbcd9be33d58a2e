// Shi-Tomasi extraction with avoidance mask, strict 3x3 NMS, per-cell top-k and grid bucketing (sm_100a).
//
// Reference behaviour: extractor.jl:24-42 (_shi_tomasi), :63-95 (detect), :116-122 (get_mask); third-party
// semantics (Images.shi_tomasi / findlocalmaxima, Kernel.gaussian, ImageDraw circle) as listed in SURVEY.md A.8-A.10.
//
// One CTA per 35x35 cell keeps the whole cell in shared memory.  The path is Float64 end to end and this file
// is compiled with -fmad=false so that responses are bit-identical to the reference arithmetic order: keypoint
// sets must be identical except for exact score ties, and near-ties must not flip either.
#include <cstdlib>

#include "common.cuh"

namespace sk {

constexpr int DET_THREADS = 256;
// ImageDraw's filled CirclePointRadius is drawn as an ellipse: pixel (i, j) belongs to it iff
// ((i - cy) / r)^2 + ((j - cx) / r)^2 < 1 evaluated in Float64 [3P ImageDraw 0.2 ellipse2d.jl] -- strict, so the four axis
// extremes and the other lattice points exactly on the circle stay outside (for a few radii, e.g. 41, the rounded sum of an
// on-circle point is 0.9999999999999999 and it is inside: the expression is evaluated exactly as written, no FMA).
__device__ __forceinline__ bool in_disc(int dy, int dx, int radius) {
    const double vy = (double)dy / (double)radius, vx = (double)dx / (double)radius;
    return vy * vy + vx * vx < 1.0;
}

// the Float64 test is only ever needed for a lattice point exactly ON the circle: kept out of line so that the common path
// does not carry its divisions
__device__ __noinline__ bool in_disc_slow(int dy, int dx, int radius) { return in_disc(dy, dx, radius); }
// largest sy >= 0 such that (sy, dx) lies inside the disc, or -1
__device__ __forceinline__ int disc_half_height(int dx, int radius) {
    const int rem = radius * radius - dx * dx;
    if (rem < 0) return -1;
    int sy = (int)sqrtf((float)rem);
    while (sy * sy > rem) --sy;
    while ((sy + 1) * (sy + 1) <= rem) ++sy;
    if (sy * sy == rem && !in_disc_slow(sy, dx, radius)) --sy;
    return sy;
}

constexpr int MAX_NEAR = 512;  // current points kept in shared memory per cell (more => slow path over the global list)

// Shared-memory plan of one cell (doubles unless noted), identical on host and device:
//   A: padded plane, image * mask first, response R later            (P*P)
//   B: three padded product planes gyy, gyx, gxx; during the mask phase the same bytes hold the y-filtered mask
//      (cs*(cs+2hw) doubles) followed by the binary mask ((cs+2hw)^2 bytes)
//   C: candidates (response double + index int, at most ceil(cs/2)^2 strict maxima); during the mask phase the list of
//      nearby current points (2*MAX_NEAR ints)
//   D: 64 ints of scan scratch
struct DetSmem { size_t oA, oB, oTmp, oM0, oCandR, oCandI, oNear, oMisc, oBits, total; };
__host__ __device__ inline DetSmem det_smem_plan(int cs, int hw) {
    const size_t P = cs + 2, pad = P * P, rw = cs + 2 * hw;
    const size_t nc = (size_t)((cs + 1) / 2) * ((cs + 1) / 2);
    DetSmem m;
    m.oA = 0;
    m.oB = pad * 8;
    m.oTmp = m.oB;
    m.oM0 = m.oTmp + (size_t)cs * rw * 8;
    size_t endB = m.oB + 3 * pad * 8;
    const size_t endMask = m.oM0 + rw * rw;
    if (endMask > endB) endB = endMask;
    endB = (endB + 15) & ~(size_t)15;
    m.oCandR = endB;
    m.oCandI = m.oCandR + nc * 8;
    m.oNear = m.oCandR;
    size_t endC = m.oCandI + nc * 4;
    const size_t endNear = m.oNear + 2 * (size_t)MAX_NEAR * 4;
    if (endNear > endC) endC = endNear;
    m.oMisc = (endC + 15) & ~(size_t)15;
    m.oBits = m.oMisc + 64 * 4;          // bit planes of the fast mask path: 64 column words + 64 "zero" + 64 "full" row words
    m.total = m.oBits + 3 * 64 * 8;
    return m;
}

size_t detect_smem_bytes(int cs, int hw) { return det_smem_plan(cs, hw).total; }

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// block-wide exclusive scan of one int per thread (DET_THREADS threads)
__device__ int block_excl_scan(int v, int* s_warp, int& total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    int base = 0, tot = 0;
    for (int i = 0; i < DET_THREADS / 32; ++i) {
        int t = s_warp[i];
        if (i < w) base += t;
        tot += t;
    }
    total = tot;
    return base + inc - v;
}

// i -> (y, x) with i = y + x*h, exact for i*h < 2^20 (h <= 64): one multiply and a shift instead of an integer division
struct DivH {
    unsigned magic; int h;
    __device__ constexpr DivH(int h_) : magic((1u << 20) / (unsigned)h_ + 1u), h(h_) {}
    __device__ __forceinline__ void split(int i, int& y, int& x) const { x = (int)(((unsigned)i * magic) >> 20); y = i - x * h; }
};

__global__ void __launch_bounds__(DET_THREADS) k_detect_cells(DetArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int cs = a.cs, hw = a.hw;
    const int cell_id = blockIdx.x, f = blockIdx.y;
    const int gy = cell_id / a.grid_w, gx = cell_id % a.grid_w;
    const int H = a.H, W = a.W;
    const int y0 = gy * cs, x0 = gx * cs;
    const int y1 = min((gy + 1) * cs, H), x1 = min((gx + 1) * cs, W);
    const int h = y1 - y0, w = x1 - x0;
    int* cnt_out = a.cell_cnt + (size_t)f * gridDim.x + cell_id;
    if (h <= 0 || w <= 0) { if (threadIdx.x == 0) *cnt_out = 0; return; }

    // padded planes: element (y, x) of the cell lives at (y+1) + (x+1)*P; the halo carries the replicate border of the
    // CELL (Images.shi_tomasi runs on the cell sub-image) so that the stencils below need no clamps
    const int P = cs + 2;
    const size_t pad = (size_t)P * P;
    const DetSmem sm = det_smem_plan(cs, hw);
    double* s_img = (double*)(smem_raw + sm.oA);   // later: the response plane
    double* s_R = s_img;
    double* s_gyy = (double*)(smem_raw + sm.oB);
    double* s_gyx = s_gyy + pad;
    double* s_gxx = s_gyx + pad;
    double* s_tmp = (double*)(smem_raw + sm.oTmp);  // mask phase only (aliases the product planes)
    unsigned char* s_m0 = (unsigned char*)(smem_raw + sm.oM0);  // mask phase only: binary mask, 1 byte per pixel
    double* s_candr = (double*)(smem_raw + sm.oCandR);
    int* s_candi = (int*)(smem_raw + sm.oCandI);
    int* s_near = (int*)(smem_raw + sm.oNear);      // mask phase only (aliases the candidates)
    int* s_misc = (int*)(smem_raw + sm.oMisc);      // [0] near count, [1..8] warp scan scratch

    const double* img = a.img + (size_t)f * H * W;
    const int tid = threadIdx.x;
    const DivH dh(h);
    const int npx = h * w;

    // ---- image * mask (extractor.jl:66-71) into the padded tile ---------------------------------
    const bool masked = a.n_cur > 0;
    if (masked) {
        const double* cur = a.cur + (size_t)f * a.n_cur * 2;
        if (tid == 0) s_misc[0] = 0;
        __syncthreads();
        const int reach = a.radius + hw;
        for (int k = tid; k < a.n_cur; k += DET_THREADS) {
            // Julia round(): ties to even == rint in the default rounding mode
            const int cy = (int)rint(cur[2 * k]), cx = (int)rint(cur[2 * k + 1]);  // 1-based
            if (cy >= y0 + 1 - reach && cy <= y1 + reach && cx >= x0 + 1 - reach && cx <= x1 + reach) {
                int slot = atomicAdd(&s_misc[0], 1);
                if (slot < MAX_NEAR) { s_near[2 * slot] = cy; s_near[2 * slot + 1] = cx; }
            }
        }
        __syncthreads();
        const int n_near = s_misc[0];
        const int rh = h + 2 * hw, rw = w + 2 * hw;
        const int r2 = a.radius * a.radius;
        const DivH drh(rh);
        // ---- fast path: the binary mask of the halo region as one 64-bit word per region column (bit = region row).  Discs are
        // rasterised column by column with one atomicAnd each; the replicated image border is a bit fill; the 13-tap y pass reads a
        // 13-bit pattern per output and only patterns that are neither all ones nor all zeros run the tap loop (same taps, same
        // order => same Float64 sums); the x pass does the same on per-row "all zero" / "all full" words.
        const bool fastmask = hw > 0 && rh <= 64 && rw <= 64 && n_near <= MAX_NEAR;
        if (fastmask) {
            unsigned long long* s_col = (unsigned long long*)(smem_raw + sm.oBits);
            unsigned long long* s_zero = s_col + 64;
            unsigned long long* s_full = s_zero + 64;
            const unsigned long long allrows = rh == 64 ? ~0ull : ((1ull << rh) - 1ull);
            for (int i = tid; i < 64; i += DET_THREADS) { s_col[i] = allrows; s_zero[i] = 0ull; s_full[i] = 0ull; }
            __syncthreads();
            const int ry0 = y0 - hw, rx0 = x0 - hw;  // image row / column (0-based) of region row / column 0
            const int ncolsd = 2 * a.radius + 1;
            for (int it = tid; it < n_near * ncolsd; it += DET_THREADS) {
                const int k = it / ncolsd, dx = it - k * ncolsd - a.radius;
                const int X = s_near[2 * k + 1] - 1 + dx;  // 0-based image column
                const int xx = X - rx0;
                if (X < 0 || X >= W || xx < 0 || xx >= rw) continue;
                const int rem = r2 - dx * dx;
                int sy = (int)sqrtf((float)rem);
                while (sy * sy > rem) --sy;
                while ((sy + 1) * (sy + 1) <= rem) ++sy;
                // strictly inside the circle the Float64 ellipse test always holds; only a lattice point ON the circle needs it
                if (sy * sy == rem && !in_disc(sy, dx, a.radius)) --sy;
                if (sy < 0) continue;
                const int Yc = s_near[2 * k] - 1;
                const int ra = max(max(Yc - sy, 0) - ry0, 0), rb = min(min(Yc + sy, H - 1) - ry0, rh - 1);
                if (ra > rb) continue;
                const int len = rb - ra + 1;
                const unsigned long long bits = (len >= 64 ? ~0ull : ((1ull << len) - 1ull)) << ra;
                atomicAnd(&s_col[xx], ~bits);
            }
            __syncthreads();
            // replicate border of the full-image mask: rows above / below the image repeat the first / last image row ...
            const int top = max(0, -ry0), bot = min(rh, H - ry0);  // region rows [top, bot) lie in the image
            const int left = max(0, -rx0), right = min(rw, W - rx0);
            for (int xx = tid + left; xx < right; xx += DET_THREADS) {
                unsigned long long wv = s_col[xx];
                if (top > 0) {
                    const unsigned long long lowm = (1ull << top) - 1ull;
                    wv = ((wv >> top) & 1ull) ? (wv | lowm) : (wv & ~lowm);
                }
                if (bot < rh) {
                    const unsigned long long him = allrows & ~((1ull << bot) - 1ull);
                    wv = ((wv >> (bot - 1)) & 1ull) ? (wv | him) : (wv & ~him);
                }
                s_col[xx] = wv;
            }
            __syncthreads();
            // ... and columns left / right of the image repeat the first / last image column
            for (int xx = tid; xx < rw; xx += DET_THREADS) {
                if (xx < left) s_col[xx] = s_col[left];
                else if (xx >= right) s_col[xx] = s_col[right - 1];
            }
            __syncthreads();
            const int nt = 2 * hw + 1;
            const unsigned pmask = (1u << nt) - 1u;
            double sall = 0.0, call = 0.0;
            for (int t = 0; t < nt; ++t) sall += a.kw[t];
            for (int t = 0; t < nt; ++t) call += a.kw[t] * sall;
            // y pass -> s_tmp[h][rw]
            for (int i = tid; i < h * rw; i += DET_THREADS) {
                int y, xx;
                dh.split(i, y, xx);
                const unsigned pat = (unsigned)(s_col[xx] >> y) & pmask;
                double acc;
                if (pat == pmask) { acc = sall; atomicOr(&s_full[y], 1ull << xx); }
                else if (pat == 0u) { acc = 0.0; atomicOr(&s_zero[y], 1ull << xx); }
                else {
                    acc = 0.0;
                    for (int t = 0; t < nt; ++t) acc += ((pat >> t) & 1u) ? a.kw[t] : 0.0;
                }
                s_tmp[i] = acc;
            }
            __syncthreads();
            for (int i = tid; i < npx; i += DET_THREADS) {
                int y, x;
                dh.split(i, y, x);
                const unsigned zb = (unsigned)(s_zero[y] >> x) & pmask, fb = (unsigned)(s_full[y] >> x) & pmask;
                double acc;
                if (zb == pmask) acc = 0.0;
                else if (fb == pmask) acc = call;
                else {
                    const double* tp = s_tmp + y + x * h;
                    acc = 0.0;
                    for (int t = 0; t < nt; ++t) acc += a.kw[t] * tp[t * h];
                }
                s_img[(y + 1) + (x + 1) * P] = img[(size_t)(y0 + y) + (size_t)(x0 + x) * H] * acc;
            }
        } else {
        // binary mask on the halo region; coordinates clamped to the image (replicate border of the blur)
        const bool inner = y0 - hw >= 0 && y1 + hw <= H && x0 - hw >= 0 && x1 + hw <= W && n_near <= MAX_NEAR;
        if (inner) {
            // no clamping inside the image: rasterise each nearby disc row by row instead of testing every pixel against
            // every point (get_mask + ImageDraw circle, extractor.jl:116-122)
            for (int i = tid; i < rh * rw; i += DET_THREADS) s_m0[i] = 1;
            __syncthreads();
            const int rows = 2 * a.radius + 1;
            for (int it = tid; it < n_near * rows; it += DET_THREADS) {
                const int k = it / rows, dy = it - k * rows - a.radius;
                const int yy = s_near[2 * k] + dy - 1 - (y0 - hw);  // region row of image row cy + dy (1-based -> 0-based)
                if (yy < 0 || yy >= rh) continue;
                const int rem = r2 - dy * dy;
                int sx = (int)sqrtf((float)rem);
                while (sx * sx > rem) --sx;
                while ((sx + 1) * (sx + 1) <= rem) ++sx;
                if (sx * sx == rem && !in_disc(dy, sx, a.radius)) --sx;  // only a lattice point ON the circle needs the Float64 test
                if (sx < 0) continue;
                const int xc = s_near[2 * k + 1] - 1 - (x0 - hw);
                const int xa = max(xc - sx, 0), xb = min(xc + sx, rw - 1);
                for (int xx = xa; xx <= xb; ++xx) s_m0[yy + xx * rh] = 0;
            }
        } else {
            for (int i = tid; i < rh * rw; i += DET_THREADS) {
                int yy, xx;
                drh.split(i, yy, xx);
                const int Y = clampi(y0 - hw + yy, 0, H - 1) + 1, X = clampi(x0 - hw + xx, 0, W - 1) + 1;  // 1-based
                unsigned char m = 1;
                if (n_near <= MAX_NEAR) {
                    for (int k = 0; k < n_near; ++k) {
                        const int dy = Y - s_near[2 * k], dx = X - s_near[2 * k + 1];
                        const int d2 = dy * dy + dx * dx;
                        if (d2 < r2 || (d2 == r2 && in_disc(dy, dx, a.radius))) { m = 0; break; }
                    }
                } else {
                    for (int k = 0; k < a.n_cur; ++k) {
                        const int dy = Y - (int)rint(cur[2 * k]), dx = X - (int)rint(cur[2 * k + 1]);
                        const int d2 = dy * dy + dx * dx;
                        if (d2 < r2 || (d2 == r2 && in_disc(dy, dx, a.radius))) { m = 0; break; }
                    }
                }
                s_m0[i] = m;
            }
        }
        __syncthreads();
        if (hw > 0) {
            // y pass -> s_tmp[h][rw]
            for (int i = tid; i < h * rw; i += DET_THREADS) {
                int y, xx;
                dh.split(i, y, xx);
                const unsigned char* mp = s_m0 + y + xx * rh;
                double acc = 0.0;
                for (int t = 0; t <= 2 * hw; ++t) acc += mp[t] ? a.kw[t] : 0.0;  // == kw[t] * mask bit for bit (mask is exactly 0 or 1)
                s_tmp[i] = acc;
            }
            __syncthreads();
            for (int i = tid; i < npx; i += DET_THREADS) {
                int y, x;
                dh.split(i, y, x);
                const double* tp = s_tmp + y + x * h;
                double acc = 0.0;
                for (int t = 0; t <= 2 * hw; ++t) acc += a.kw[t] * tp[t * h];
                s_img[(y + 1) + (x + 1) * P] = img[(size_t)(y0 + y) + (size_t)(x0 + x) * H] * acc;
            }
        } else {
            for (int i = tid; i < npx; i += DET_THREADS) {
                int y, x;
                dh.split(i, y, x);
                s_img[(y + 1) + (x + 1) * P] = s_m0[i] ? img[(size_t)(y0 + y) + (size_t)(x0 + x) * H] : img[(size_t)(y0 + y) + (size_t)(x0 + x) * H] * 0.0;
            }
        }
        }  // generic mask path
    } else {
        for (int i = tid; i < npx; i += DET_THREADS) {
            int y, x;
            dh.split(i, y, x);
            s_img[(y + 1) + (x + 1) * P] = img[(size_t)(y0 + y) + (size_t)(x0 + x) * H];
        }
    }
    __syncthreads();
    // replicate halo of a padded plane: rows first (x in 1..w), then full columns including the corners
    auto fill_halo = [&](double* pl) {
        for (int x = tid; x < w; x += DET_THREADS) {
            pl[0 + (x + 1) * P] = pl[1 + (x + 1) * P];
            pl[(h + 1) + (x + 1) * P] = pl[h + (x + 1) * P];
        }
        __syncthreads();
        for (int y = tid; y < h + 2; y += DET_THREADS) {
            pl[y] = pl[y + P];
            pl[y + (w + 1) * P] = pl[y + w * P];
        }
        __syncthreads();
    };
    fill_halo(s_img);

    // ---- Shi-Tomasi response on the cell sub-image (replicate border at the cell edge) --------
    // NPT independent pixels per thread and round are unrolled together: the Float64 chains of one pixel are short on ILP
    constexpr int NPT = 5;
    const double k9 = 1.0 / 9.0;  // the 3x3 mean is imfilter with a (1/9)-valued kernel: products accumulated tap by tap
    for (int base = 0; base < npx; base += DET_THREADS * NPT) {
#pragma unroll
        for (int k = 0; k < NPT; ++k) {
            const int i = base + tid + k * DET_THREADS;
            if (i < npx) {
                int y, x;
                dh.split(i, y, x);
                const double* c = s_img + (y + 1) + (x + 1) * P;  // centre
                const double mm = c[-1 - P], m0 = c[-P], mp = c[1 - P];   // column x-1: rows y-1, y, y+1
                const double zm = c[-1], zp = c[1];                         // column x
                const double pm = c[-1 + P], p0 = c[P], pp = c[1 + P];     // column x+1
                const double g_y = ((mp - mm) + 2.0 * (zp - zm) + (pp - pm)) / 8.0;
                const double g_x = ((pm - mm) + 2.0 * (p0 - m0) + (pp - mp)) / 8.0;
                const int o = (y + 1) + (x + 1) * P;
                // the box filter multiplies every tap by 1/9 (imfilter with a (1/9)-valued kernel): the rounded product is the
                // same for each of the nine windows a pixel belongs to, so it is formed once here
                s_gyy[o] = k9 * (g_y * g_y); s_gyx[o] = k9 * (g_y * g_x); s_gxx[o] = k9 * (g_x * g_x);
            }
        }
    }
    __syncthreads();
    // the three product planes share the halo fill loops
    for (int x = tid; x < w; x += DET_THREADS) {
        const int t = (x + 1) * P;
        s_gyy[t] = s_gyy[1 + t]; s_gyy[h + 1 + t] = s_gyy[h + t];
        s_gyx[t] = s_gyx[1 + t]; s_gyx[h + 1 + t] = s_gyx[h + t];
        s_gxx[t] = s_gxx[1 + t]; s_gxx[h + 1 + t] = s_gxx[h + t];
    }
    __syncthreads();
    for (int y = tid; y < h + 2; y += DET_THREADS) {
        s_gyy[y] = s_gyy[y + P]; s_gyy[y + (w + 1) * P] = s_gyy[y + w * P];
        s_gyx[y] = s_gyx[y + P]; s_gyx[y + (w + 1) * P] = s_gyx[y + w * P];
        s_gxx[y] = s_gxx[y + P]; s_gxx[y + (w + 1) * P] = s_gxx[y + w * P];
    }
    // R halo = -inf so that out-of-cell neighbours never block a maximum
    for (int t = tid; t < 2 * (w + 2) + 2 * h; t += DET_THREADS) {
        int yy, xx;
        if (t < w + 2) { yy = 0; xx = t; }
        else if (t < 2 * (w + 2)) { yy = h + 1; xx = t - (w + 2); }
        else if (t < 2 * (w + 2) + h) { yy = t - 2 * (w + 2) + 1; xx = 0; }
        else { yy = t - 2 * (w + 2) - h + 1; xx = w + 1; }
        s_R[yy + xx * P] = -__longlong_as_double(0x7ff0000000000000LL);
    }
    __syncthreads();
    for (int base = 0; base < npx; base += DET_THREADS * NPT) {
#pragma unroll
        for (int k = 0; k < NPT; ++k) {
            const int i = base + tid + k * DET_THREADS;
            if (i < npx) {
                int y, x;
                dh.split(i, y, x);
                const int o = (y + 1) + (x + 1) * P;
                double sa = 0.0, sb = 0.0, sc = 0.0;
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx)
#pragma unroll
                    for (int dy = -1; dy <= 1; ++dy) {
                        const int j = o + dy + dx * P;
                        sa += s_gyy[j]; sb += s_gyx[j]; sc += s_gxx[j];
                    }
                s_R[o] = ((sa + sc) - sqrt((sa - sc) * (sa - sc) + 4.0 * sb * sb)) / 2.0;
            }
        }
    }
    __syncthreads();

    // ---- strict 3x3 local maxima, enumerated column-major (findlocalmaxima) -------------------
    int n_cand = 0;
    {
        const int PT = (npx + DET_THREADS - 1) / DET_THREADS;  // contiguous pixels per thread => one block scan keeps the order
        const int i0 = tid * PT, i1 = min(npx, i0 + PT);
        unsigned mask = 0;  // PT <= 16 for cells up to 64 x 64
        for (int i = i0; i < i1; ++i) {
            int y, x;
            dh.split(i, y, x);
            const double* c = s_R + (y + 1) + (x + 1) * P;
            const double r = c[0];
            const int ismax = (c[-1 - P] < r) & (c[-P] < r) & (c[1 - P] < r) & (c[-1] < r) & (c[1] < r) & (c[-1 + P] < r) & (c[P] < r) & (c[1 + P] < r);
            mask |= (unsigned)ismax << (i - i0);
        }
        int tot;
        int pos = block_excl_scan(__popc(mask), s_misc + 1, tot);
        for (int i = i0; i < i1; ++i)
            if (mask >> (i - i0) & 1u) {
                int y, x;
                dh.split(i, y, x);
                s_candr[pos] = s_R[(y + 1) + (x + 1) * P];
                s_candi[pos] = i;
                ++pos;
            }
        n_cand = tot;
    }
    __syncthreads();

    // ---- top-k by response, stable (sortperm with lt = >), threshold, emit in column-major order ----
    int64_t* out = a.cell_out + ((size_t)f * gridDim.x + cell_id) * a.slots * 2;
    int n_sel = 0;
    for (int base = 0; base < n_cand; base += DET_THREADS) {
        const int i = base + tid;
        int sel = 0;
        if (i < n_cand) {
            const double r = s_candr[i];
            int rank = 0;
            for (int j = 0; j < n_cand; ++j) {
                const double rj = s_candr[j];
                rank += (rj > r) || (rj == r && j < i);
            }
            sel = (rank < a.k_cell) && !(r < a.min_resp);
        }
        int tot;
        const int pos = block_excl_scan(sel, s_misc + 1, tot);
        if (sel && n_sel + pos < a.slots) {
            int y, x;
            dh.split(s_candi[i], y, x);
            out[2 * (n_sel + pos)] = (int64_t)y + 1 + y0;
            out[2 * (n_sel + pos) + 1] = (int64_t)x + 1 + x0;
        }
        n_sel += tot;
    }
    if (tid == 0) *cnt_out = n_sel;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Register-tiled variant (round 2).  Same arithmetic, same operation order, same outputs as k_detect_cells; what changes is how
// the work is cut.  Every stencil phase walks STRIPS of DET_RS pixels with a sliding window in registers (a new row costs three
// shared-memory loads per plane instead of nine), the replicate border of the cell is a clamp on the neighbour column / row
// instead of halo-fill passes with their barriers, the 13-tap y pass of the binary mask is one load from a table of the 2^13
// possible tap-subset sums (built on the host with the same sequential Float64 additions), the "all zero / all ones" row words
// come from 13 shifts-and-ANDs per column plus a ballot transpose instead of one shared-memory atomic per pixel, and the
// current points are rasterised in chunks of MAX_NEAR so that no cell ever needs a slow path.
// Covers: no mask, or a 13-tap mask blur (sigma_mask = 3, the reference's default) with cell_size + 12 <= 64.
// ---------------------------------------------------------------------------------------------------------------------------
// one pixel of the staged frames as the Float64 the reference computes on: Gray{Float64}.(img) of 8-bit data is k / 255
template <bool F64SRC>
__device__ __forceinline__ double det_px(const DetArgs& a, size_t i) {
    if (F64SRC || a.src_dtype == SLAMKLT_F64) return __ldg(reinterpret_cast<const double*>(a.src) + i);
    if (a.src_dtype == SLAMKLT_U8) return (double)__ldg(reinterpret_cast<const uint8_t*>(a.src) + i) / 255.0;
    return (double)__ldg(reinterpret_cast<const float*>(a.src) + i);
}

constexpr int DET_RS = 5;
constexpr int DET2_HW = 6, DET2_NT = 2 * DET2_HW + 1;

struct Det2Smem { size_t oImg, oG, oTmp, oNear, oCandR, oCandI, oMaskW, oMisc, oBits, total; };
__host__ __device__ inline Det2Smem det2_smem_plan(int cs, int hw) {
    const size_t P = cs + 2, pad = P * P, rw = cs + 2 * hw;
    const size_t nc = (size_t)((cs + 1) / 2) * ((cs + 1) / 2);
    const size_t nstrips = (size_t)cs * ((cs + DET_RS - 1) / DET_RS);
    Det2Smem m;
    m.oImg = 0;                                   // padded plane: image * mask, later the response (its halo is -inf)
    m.oG = pad * 8;                               // three padded product planes (halo unused: neighbours are clamped)
    const size_t endG = m.oG + 3 * pad * 8;
    m.oTmp = m.oG;                                // mask phase: y-filtered mask, cs x rw doubles ...
    m.oNear = m.oTmp + (size_t)cs * rw * 8;       // ... and the chunk of nearby current points
    const size_t endMask = m.oNear + 2 * (size_t)MAX_NEAR * 4;
    m.oCandR = m.oG;                              // after the response pass the product planes are dead: candidates ...
    m.oCandI = m.oCandR + nc * 8;
    m.oMaskW = m.oCandI + nc * 4;                 // ... and the per-strip maxima masks
    const size_t endC = m.oMaskW + nstrips * 4;
    size_t end = endG > endMask ? endG : endMask;
    if (endC > end) end = endC;
    m.oMisc = (end + 15) & ~(size_t)15;
    m.oBits = m.oMisc + 64 * 4;                   // five bit planes of 64 words: column mask, F, Z (per column), full, zero (per row)
    m.total = m.oBits + 5 * 64 * 8;
    return m;
}
size_t detect2_smem_bytes(int cs, int hw) { return det2_smem_plan(cs, hw).total; }

// CS: the cell size as a compile-time constant (plane pitch and offsets become immediates), or 0 = taken from the arguments
// Current points of a frame that can reach a row of cells (rounded row within radius + hw of the row's pixels), as int2 (y, x)
// 1-based: every cell of the row then scans this short list instead of all n_cur points.  The order inside a list is arbitrary
// (the discs are cleared with atomic ANDs).  a.bin_pts: [frame][cell row][n_cur] int2, a.bin_cnt: [frame][cell row].
__global__ void __launch_bounds__(DET_THREADS) k_detect_bin(DetArgs a) {
    __shared__ int s_n;
    const int gy = blockIdx.x, f = blockIdx.y;
    const int y0 = gy * a.cs, y1 = min((gy + 1) * a.cs, a.H);
    const int reach = a.radius + a.hw;
    const double* cur = a.cur + (size_t)f * a.n_cur * 2;
    int2* out = a.bin_pts + ((size_t)f * gridDim.x + gy) * a.n_cur;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int k = threadIdx.x; k < a.n_cur; k += DET_THREADS) {
        const int cy = (int)rint(cur[2 * k]), cx = (int)rint(cur[2 * k + 1]);  // Julia round(): ties to even; 1-based
        if (cy >= y0 + 1 - reach && cy <= y1 + reach) out[atomicAdd(&s_n, 1)] = make_int2(cy, cx);
    }
    __syncthreads();
    if (threadIdx.x == 0) a.bin_cnt[(size_t)f * gridDim.x + gy] = s_n;
}

#ifndef DET2_MINB
#define DET2_MINB 4
#endif
// F64SRC: the staged frames are Float64 (no run-time switch on the pixel type in the load path)
template <bool MASKED, int CS, bool F64SRC>
__global__ void __launch_bounds__(DET_THREADS, DET2_MINB) k_detect_cells2(DetArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int RS = DET_RS, hw = MASKED ? DET2_HW : 0, nt = DET2_NT;
    const int cs = CS ? CS : a.cs;
    const int cell_id = blockIdx.x, f = blockIdx.y;
    const int gy = cell_id / a.grid_w, gx = cell_id % a.grid_w;
    const int H = a.H, W = a.W;
    const int y0 = gy * cs, x0 = gx * cs;
    const int y1 = min((gy + 1) * cs, H), x1 = min((gx + 1) * cs, W);
    const int h = y1 - y0, w = x1 - x0;
    int* cnt_out = a.cell_cnt + (size_t)f * gridDim.x + cell_id;
    if (h <= 0 || w <= 0) { if (threadIdx.x == 0) *cnt_out = 0; return; }

    const int P = cs + 2;
    const size_t pad = (size_t)P * P;
    const Det2Smem sm = det2_smem_plan(cs, hw);
    double* s_img = (double*)(smem_raw + sm.oImg);
    double* s_R = s_img;
    double* s_gyy = (double*)(smem_raw + sm.oG);
    double* s_gyx = s_gyy + pad;
    double* s_gxx = s_gyx + pad;
    double* s_tmp = (double*)(smem_raw + sm.oTmp);
    int* s_near = (int*)(smem_raw + sm.oNear);
    double* s_candr = (double*)(smem_raw + sm.oCandR);
    int* s_candi = (int*)(smem_raw + sm.oCandI);
    unsigned* s_maskw = (unsigned*)(smem_raw + sm.oMaskW);
    int* s_misc = (int*)(smem_raw + sm.oMisc);

    const size_t img0 = (size_t)f * H * W;   // first element of this frame in a.src
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // (the magic numbers of the reference's full cells are compile-time constants; edge cells compute theirs)
    const DivH dh = (CS && h == CS) ? DivH(CS) : DivH(h), dw = (CS && w == CS) ? DivH(CS) : DivH(w);
    const int npx = h * w;

    // halo of the response plane = -inf, so that out-of-cell neighbours never block a maximum.  The plane shares its bytes with
    // the image tile, whose halo nobody reads (neighbours are clamped), so the ring can be written right away.
    for (int t = tid; t < 2 * (w + 2) + 2 * h; t += DET_THREADS) {
        int yy, xx;
        if (t < w + 2) { yy = 0; xx = t; }
        else if (t < 2 * (w + 2)) { yy = h + 1; xx = t - (w + 2); }
        else if (t < 2 * (w + 2) + h) { yy = t - 2 * (w + 2) + 1; xx = 0; }
        else { yy = t - 2 * (w + 2) - h + 1; xx = w + 1; }
        s_R[yy + xx * P] = -__longlong_as_double(0x7ff0000000000000LL);
    }

    // ---- image * mask (extractor.jl:66-71, 116-122) into the padded tile ---------------------------------------------------
    if (MASKED) {
        unsigned long long* s_col = (unsigned long long*)(smem_raw + sm.oBits);
        unsigned long long* s_F = s_col + 64;
        unsigned long long* s_Z = s_F + 64;
        unsigned long long* s_full = s_Z + 64;
        unsigned long long* s_zero = s_full + 64;
        const double* cur = a.cur + (size_t)f * a.n_cur * 2;
        const int rh = h + 2 * hw, rw = w + 2 * hw;
        const int reach = a.radius + hw;
        const unsigned long long allrows = rh == 64 ? ~0ull : ((1ull << rh) - 1ull);
        for (int i = tid; i < 64; i += DET_THREADS) { s_col[i] = allrows; s_F[i] = 0ull; s_Z[i] = 0ull; }
        const int ry0 = y0 - hw, rx0 = x0 - hw;  // image row / column (0-based) of region row / column 0
        const int ncolsd = 2 * a.radius + 1;
        // half height of the disc at every column offset dx, the same for every point: the largest sy with (sy, dx) inside
        // ImageDraw's ellipse test (-1: the column is empty).  A disc wider than the table falls back to the per-item computation.
        const bool sy_tab = a.sy_valid != 0;   // (host-built: the table is the same for every cell of every frame)
        const DivH dcol(ncolsd);
        const int2* const bin = a.bin_pts ? a.bin_pts + ((size_t)f * a.grid_h + gy) * a.n_cur : nullptr;
        const int n_scan = bin ? a.bin_cnt[(size_t)f * a.grid_h + gy] : a.n_cur;
        for (int k0 = 0; k0 < n_scan; k0 += MAX_NEAR) {
            if (tid == 0) s_misc[0] = 0;
            __syncthreads();
            const int k1 = min(n_scan, k0 + MAX_NEAR);
            for (int k = k0 + tid; k < k1; k += DET_THREADS) {
                int cy, cx;
                if (bin) { const int2 p = __ldg(bin + k); cy = p.x; cx = p.y; }
                else { cy = (int)rint(cur[2 * k]); cx = (int)rint(cur[2 * k + 1]); }  // Julia round(): ties to even; 1-based
                if (cy >= y0 + 1 - reach && cy <= y1 + reach && cx >= x0 + 1 - reach && cx <= x1 + reach) {
                    const int slot = atomicAdd(&s_misc[0], 1);
                    s_near[2 * slot] = cy; s_near[2 * slot + 1] = cx;
                }
            }
            __syncthreads();
            const int n_near = s_misc[0];
            // discs column by column: one atomicAnd clears a column's rows (get_mask + ImageDraw circle)
            for (int it = tid; it < n_near * ncolsd; it += DET_THREADS) {
                int k, di;
                if (n_near * ncolsd * ncolsd < (1 << 20)) dcol.split(it, di, k);   // it = di + k * ncolsd (DivH is exact while it * ncolsd < 2^20)
                else { k = it / ncolsd; di = it - k * ncolsd; }
                const int dx = di - a.radius;
                const int X = s_near[2 * k + 1] - 1 + dx;  // 0-based image column
                const int xx = X - rx0;
                if (X < 0 || X >= W || xx < 0 || xx >= rw) continue;
                const int sy = sy_tab ? (int)a.sy[di] : disc_half_height(dx, a.radius);
                if (sy < 0) continue;
                const int Yc = s_near[2 * k] - 1;
                const int ra = max(max(Yc - sy, 0) - ry0, 0), rb = min(min(Yc + sy, H - 1) - ry0, rh - 1);
                if (ra > rb) continue;
                // rows [ra, rb] of the column word, cleared as two 32-bit halves (native shared-memory atomics)
                unsigned* const wp = reinterpret_cast<unsigned*>(&s_col[xx]);
                const unsigned lo = ra < 32 ? ((rb >= 31 ? ~0u : ((2u << rb) - 1u)) & ~((1u << ra) - 1u)) : 0u;
                const unsigned hi = rb >= 32 ? ((rb >= 63 ? ~0u : ((2u << (rb - 32)) - 1u)) & (ra > 32 ? ~((1u << (ra - 32)) - 1u) : ~0u)) : 0u;
                if (lo) atomicAnd(wp, ~lo);
                if (hi) atomicAnd(wp + 1, ~hi);
            }
            __syncthreads();
        }
        if (n_scan <= 0) __syncthreads();
        // replicate border of the full-image mask: rows above / below the image repeat the first / last image row ...
        const int top = max(0, -ry0), bot = min(rh, H - ry0);  // region rows [top, bot) lie in the image
        const int left = max(0, -rx0), right = min(rw, W - rx0);
        for (int xx = tid + left; xx < right; xx += DET_THREADS) {
            unsigned long long wv = s_col[xx];
            if (top > 0) {
                const unsigned long long lowm = (1ull << top) - 1ull;
                wv = ((wv >> top) & 1ull) ? (wv | lowm) : (wv & ~lowm);
            }
            if (bot < rh) {
                const unsigned long long him = allrows & ~((1ull << bot) - 1ull);
                wv = ((wv >> (bot - 1)) & 1ull) ? (wv | him) : (wv & ~him);
            }
            s_col[xx] = wv;
        }
        __syncthreads();
        // ... and columns left / right of the image repeat the first / last image column
        if (left > 0 || right < rw) {
            for (int xx = tid; xx < rw; xx += DET_THREADS) {
                if (xx < left) s_col[xx] = s_col[left];
                else if (xx >= right) s_col[xx] = s_col[right - 1];
            }
            __syncthreads();
        }
        // per column: F bit y = rows y .. y+12 all ones, Z bit y = all zeros; then transposed into row words by ballots
        for (int xx = tid; xx < rw; xx += DET_THREADS) {
            const unsigned long long c = s_col[xx], nc = ~c;
            unsigned long long fw = c, zw = nc;
#pragma unroll
            for (int t = 1; t < nt; ++t) { fw &= c >> t; zw &= nc >> t; }
            s_F[xx] = fw; s_Z[xx] = zw;
        }
        __syncthreads();
        for (int y = wid; y < h; y += DET_THREADS / 32) {
            const unsigned f0 = __ballot_sync(FULL, (s_F[lane] >> y) & 1ull), f1 = __ballot_sync(FULL, (s_F[lane + 32] >> y) & 1ull);
            const unsigned z0 = __ballot_sync(FULL, (s_Z[lane] >> y) & 1ull), z1 = __ballot_sync(FULL, (s_Z[lane + 32] >> y) & 1ull);
            if (lane == 0) { s_full[y] = (unsigned long long)f0 | ((unsigned long long)f1 << 32); s_zero[y] = (unsigned long long)z0 | ((unsigned long long)z1 << 32); }
        }
        // y pass -> s_tmp[rw][h]: the sum of the taps whose mask bit is set, taken from the table of all 2^13 subsets
        const unsigned pmask = (1u << nt) - 1u;
        for (int i = tid; i < h * rw; i += DET_THREADS) {
            int y, xx;
            dh.split(i, y, xx);
            s_tmp[y + xx * cs] = __ldg(a.ytab + ((unsigned)(s_col[xx] >> y) & pmask));   // (pitch cs: immediates when CS is fixed)
        }
        const double sall = __ldg(a.ytab + pmask);
        double call = 0.0;
#pragma unroll
        for (int t = 0; t < nt; ++t) call += a.kw[t] * sall;
        __syncthreads();
        // x pass on strips of RS pixels along x (lanes = consecutive rows); the 13-tap windows of a strip share their loads
        const int nxs = (w + RS - 1) / RS;
        for (int s = tid; s < h * nxs; s += DET_THREADS) {
            int y, xs;
            dh.split(s, y, xs);
            const int xa = xs * RS, nx = min(RS, w - xa);
            const unsigned wmask = (1u << (nx + nt - 1)) - 1u;
            const bool allz = ((unsigned)(s_zero[y] >> xa) & wmask) == wmask;   // (bits of Z / F beyond rw are zero)
            const bool allf = ((unsigned)(s_full[y] >> xa) & wmask) == wmask;
            const size_t ip = img0 + (size_t)(y0 + y) + (size_t)(x0 + xa) * H;
            double* op = s_img + (y + 1) + (xa + 1) * P;
            if (allz || allf) {  // every 13 x 13 window of the strip is all zeros / all ones
                const double acc = allz ? 0.0 : call;
#pragma unroll
                for (int j = 0; j < RS; ++j)
                    if (j < nx) op[j * P] = det_px<F64SRC>(a, ip + (size_t)j * H) * acc;
            } else {
                // (a short last strip reads up to RS - 1 columns past the filtered mask: still inside this block's scratch, never used)
                const double* tp = s_tmp + y + xa * cs;
                double v[RS + nt - 1];
#pragma unroll
                for (int t = 0; t < RS + nt - 1; ++t) v[t] = tp[t * cs];
#pragma unroll
                for (int j = 0; j < RS; ++j) {
                    if (j < nx) {
                        double acc = 0.0;
#pragma unroll
                        for (int t = 0; t < nt; ++t) acc += a.kw[t] * v[j + t];
                        op[j * P] = det_px<F64SRC>(a, ip + (size_t)j * H) * acc;
                    }
                }
            }
        }
    } else {
        for (int i = tid; i < npx; i += DET_THREADS) {
            int y, x;
            dh.split(i, y, x);
            s_img[(y + 1) + (x + 1) * P] = det_px<F64SRC>(a, img0 + (size_t)(y0 + y) + (size_t)(x0 + x) * H);
        }
    }
    __syncthreads();

    // ---- strips along y: strip s = seg * w + x covers rows [seg * RS, seg * RS + RS) of column x (lanes = consecutive columns) ----
    const int nseg = (h + RS - 1) / RS;
    const int nstrips = w * nseg;
    const double k9 = 1.0 / 9.0;  // the 3x3 mean is imfilter with a (1/9)-valued kernel: products accumulated tap by tap
    // Sobel / 8 and the three products (Images.shi_tomasi on the cell sub-image: replicate border at the cell edge = clamped neighbours)
    for (int s = tid; s < nstrips; s += DET_THREADS) {
        int x, seg;
        dw.split(s, x, seg);
        const int ys = seg * RS, ye = min(ys + RS, h);
        const double* cz = s_img + 1 + (x + 1) * P;
        const double* cm = s_img + 1 + (max(x - 1, 0) + 1) * P;
        const double* cp = s_img + 1 + (min(x + 1, w - 1) + 1) * P;
        const int yp = max(ys - 1, 0);
        double a_m = cm[yp], a_z = cz[yp], a_p = cp[yp];   // row y-1 of columns x-1, x, x+1
        double b_m = cm[ys], b_z = cz[ys], b_p = cp[ys];   // row y
        // rows past the end of a short last strip are computed on clamped (valid) inputs and simply not stored: the body has no
        // branch, so the sliding window is renamed by the unroller instead of moved
        double* const og = s_gyy + (x + 1) * P + 1 + ys;
#pragma unroll
        for (int r = 0; r < RS; ++r) {
            const int y = ys + r;
            const int yn = min(y + 1, h - 1);
            const double c_m = cm[yn], c_z = cz[yn], c_p = cp[yn];  // row y+1
            const double g_y = ((c_m - a_m) + 2.0 * (c_z - a_z) + (c_p - a_p)) / 8.0;
            const double g_x = ((a_p - a_m) + 2.0 * (b_p - b_m) + (c_p - c_m)) / 8.0;
            if (y < ye) { og[r] = k9 * (g_y * g_y); og[r + pad] = k9 * (g_y * g_x); og[r + 2 * pad] = k9 * (g_x * g_x); }
            a_m = b_m; a_z = b_z; a_p = b_p;
            b_m = c_m; b_z = c_z; b_p = c_p;
        }
    }
    __syncthreads();
    // 3x3 sums (column outer, row inner, like the reference's tap order) and the smaller eigenvalue
    for (int s = tid; s < nstrips; s += DET_THREADS) {
        int x, seg;
        dw.split(s, x, seg);
        const int ys = seg * RS, ye = min(ys + RS, h);
        const double* const pc[3] = {s_gyy + 1 + (max(x - 1, 0) + 1) * P, s_gyy + 1 + (x + 1) * P, s_gyy + 1 + (min(x + 1, w - 1) + 1) * P};
        double A[3][3], B[3][3], C[3][3];  // [column][row slot]: rows y-1, y, y+1 of gyy, gyx, gxx
        const int yp = max(ys - 1, 0);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            A[c][0] = pc[c][yp]; B[c][0] = pc[c][yp + pad]; C[c][0] = pc[c][yp + 2 * pad];
            A[c][1] = pc[c][ys]; B[c][1] = pc[c][ys + pad]; C[c][1] = pc[c][ys + 2 * pad];
        }
        double* const orow = s_R + (x + 1) * P + 1 + ys;
#pragma unroll
        for (int r = 0; r < RS; ++r) {
            const int y = ys + r;
            const int yn = min(y + 1, h - 1);
#pragma unroll
            for (int c = 0; c < 3; ++c) { A[c][2] = pc[c][yn]; B[c][2] = pc[c][yn + pad]; C[c][2] = pc[c][yn + 2 * pad]; }
            double sa = 0.0, sb = 0.0, sc = 0.0;
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int q = 0; q < 3; ++q) { sa += A[c][q]; sb += B[c][q]; sc += C[c][q]; }
            // sqrt(0) = 0: a warp whose rows all lie in a flat (masked-out) region skips the Float64 square root, same value
            const double disc = (sa - sc) * (sa - sc) + 4.0 * sb * sb;
            double root = 0.0;
            if (__any_sync(__activemask(), disc != 0.0)) root = sqrt(disc);   // (the last warp's idle lanes are outside this loop)
            const double resp = ((sa + sc) - root) / 2.0;
            if (y < ye) orow[r] = resp;
#pragma unroll
            for (int c = 0; c < 3; ++c) { A[c][0] = A[c][1]; B[c][0] = B[c][1]; C[c][0] = C[c][1]; A[c][1] = A[c][2]; B[c][1] = B[c][2]; C[c][1] = C[c][2]; }
        }
    }
    __syncthreads();
    // ---- strict 3x3 local maxima (findlocalmaxima): one 5-bit word per strip, stored in column-major strip order ----------------
    for (int s = tid; s < nstrips; s += DET_THREADS) {
        int x, seg;
        dw.split(s, x, seg);
        const int ys = seg * RS, ye = min(ys + RS, h);
        const double* cz = s_R + 1 + (x + 1) * P;
        const double* cm = cz - P;
        const double* cp = cz + P;
        double a_m = cm[ys - 1], a_z = cz[ys - 1], a_p = cp[ys - 1];
        double b_m = cm[ys], b_z = cz[ys], b_p = cp[ys];
        unsigned bits = 0;
#pragma unroll
        for (int r = 0; r < RS; ++r) {
            const int yn = min(ys + r + 1, h);  // (row h is the halo)
            const double c_m = cm[yn], c_z = cz[yn], c_p = cp[yn];
            const double v = b_z;
            const int ismax = (a_m < v) & (b_m < v) & (c_m < v) & (a_z < v) & (c_z < v) & (a_p < v) & (b_p < v) & (c_p < v);
            bits |= (unsigned)ismax << r;
            a_m = b_m; a_z = b_z; a_p = b_p; b_m = c_m; b_z = c_z; b_p = c_p;
        }
        s_maskw[x * nseg + seg] = bits & ((1u << (ye - ys)) - 1u);
    }
    __syncthreads();
    int n_cand = 0;
    for (int base = 0; base < nstrips; base += DET_THREADS) {
        const int t = base + tid;
        const unsigned m = t < nstrips ? s_maskw[t] : 0u;
        int tot;
        int pos = n_cand + block_excl_scan(__popc(m), s_misc + 1, tot);
        if (m) {
            const int x = t / nseg, seg = t - x * nseg;
#pragma unroll
            for (int r = 0; r < RS; ++r)
                if (m >> r & 1u) {
                    const int y = seg * RS + r;
                    s_candr[pos] = s_R[(y + 1) + (x + 1) * P];
                    s_candi[pos] = y + x * h;
                    ++pos;
                }
        }
        n_cand += tot;
    }
    __syncthreads();

    // ---- top-k by response, stable (sortperm with lt = >), threshold, emit in column-major order ----
    int64_t* out = a.cell_out + ((size_t)f * gridDim.x + cell_id) * a.slots * 2;
    int n_sel = 0;
    for (int base = 0; base < n_cand; base += DET_THREADS) {
        const int i = base + tid;
        int sel = 0;
        if (i < n_cand) {
            const double r = s_candr[i];
            int rank = 0;
            for (int j = 0; j < n_cand; ++j) {
                const double rj = s_candr[j];
                rank += (rj > r) || (rj == r && j < i);
            }
            sel = (rank < a.k_cell) && !(r < a.min_resp);
        }
        int tot;
        const int pos = block_excl_scan(sel, s_misc + 1, tot);
        if (sel && n_sel + pos < a.slots) {
            int y, x;
            dh.split(s_candi[i], y, x);
            out[2 * (n_sel + pos)] = (int64_t)y + 1 + y0;
            out[2 * (n_sel + pos) + 1] = (int64_t)x + 1 + x0;
        }
        n_sel += tot;
    }
    if (tid == 0) *cnt_out = n_sel;
}

// cells are visited y-outer, x-inner (extractor.jl:81); one CTA per frame concatenates the cell lists
__global__ void __launch_bounds__(DET_THREADS) k_detect_compact(DetArgs a, int n_cells) {
    __shared__ int s_warp[DET_THREADS / 32 + 1];
    __shared__ int s_base;
    const int f = blockIdx.x;
    const int* cnt = a.cell_cnt + (size_t)f * n_cells;
    int64_t* out = a.out + (size_t)f * a.cap * 2;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int base = 0; base < n_cells; base += DET_THREADS) {
        const int c = base + threadIdx.x;
        const int n = c < n_cells ? cnt[c] : 0;
        int tot;
        const int pos = s_base + block_excl_scan(n, s_warp, tot);
        if (n > 0) {
            const int64_t* src = a.cell_out + ((size_t)f * n_cells + c) * a.slots * 2;
            for (int k = 0; k < n && k < a.slots; ++k)
                if (pos + k < a.cap) { out[2 * (pos + k)] = src[2 * k]; out[2 * (pos + k) + 1] = src[2 * k + 1]; }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) a.n_out[f] = s_base;
}

// shapes the register-tiled kernel covers (SLAMKLT_DETECT_V1=1 forces the first kernel: A/B timing)
bool detect2_supported(const DetArgs& a) {
    static const bool v1 = getenv("SLAMKLT_DETECT_V1") != nullptr;
    if (v1) return false;
    if (a.n_cur <= 0) return true;
    return a.hw == DET2_HW && a.cs + 2 * DET2_HW <= 64 && a.ytab != nullptr;
}

int launch_detect(cudaStream_t s, const DetArgs& a, const Hook* hk) {
    const int n_cells = a.grid_h * a.grid_w;
    dim3 grid(n_cells, a.n_frames);
    if (detect2_supported(a)) {
        const bool masked = a.n_cur > 0;
        const size_t smem = detect2_smem_bytes(a.cs, masked ? DET2_HW : 0);
        auto go = [&](auto kern) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            kern<<<grid, DET_THREADS, smem, s>>>(a);
        };
        // the reference's cell size (35) gets the instantiation with compile-time plane offsets
        const bool f64 = a.src_dtype == SLAMKLT_F64;
        if (masked) {
            if (a.bin_pts) {   // current points within reach of each row of cells, once per frame instead of once per cell
                mark(hk, "k_detect_bin");
                k_detect_bin<<<dim3(a.grid_h, a.n_frames), DET_THREADS, 0, s>>>(a);
            }
            mark(hk, "k_detect_cells");
            if (a.cs == 35) { if (f64) go(k_detect_cells2<true, 35, true>); else go(k_detect_cells2<true, 35, false>); }
            else { if (f64) go(k_detect_cells2<true, 0, true>); else go(k_detect_cells2<true, 0, false>); }
        } else {
            mark(hk, "k_detect_cells");
            if (a.cs == 35) { if (f64) go(k_detect_cells2<false, 35, true>); else go(k_detect_cells2<false, 35, false>); }
            else { if (f64) go(k_detect_cells2<false, 0, true>); else go(k_detect_cells2<false, 0, false>); }
        }
        mark(hk, "k_detect_compact");
        k_detect_compact<<<a.n_frames, DET_THREADS, 0, s>>>(a, n_cells);
        return (masked && a.bin_pts) ? 3 : 2;
    }
    const size_t smem = detect_smem_bytes(a.cs, a.hw);
    cudaFuncSetAttribute(k_detect_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    mark(hk, "k_detect_cells");
    k_detect_cells<<<grid, DET_THREADS, smem, s>>>(a);
    mark(hk, "k_detect_compact");
    k_detect_compact<<<a.n_frames, DET_THREADS, 0, s>>>(a, n_cells);
    return 2;
}

}  // namespace sk
