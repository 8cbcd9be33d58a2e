/*
 * slam_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Float64 CPU restatement of the SLAM.jl KLT front-end hot path (pyramid build,
 * pyramidal Lucas-Kanade, forward-backward tracking, Shi-Tomasi extraction).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (libslamklt.so) never links or calls it.
 *
 * PARITY UNPINNED: the reference repository holds no tests, fixtures or golden
 * vectors, Julia is not installed in this image, and most
 * of the numerics live in third-party Julia packages that are not vendored under
 * /root/reference (Images 0.24, ImageFiltering 0.6/0.7, ImageTransformations,
 * Interpolations 0.13, ImageFeatures 0.4, ImageDraw 0.2; ranges from
 * Project.toml:28-45, no Manifest).  Their published algorithms are restated here
 * and marked [3P]; everything marked [REF] follows the cited reference lines.
 * Independent pins live in tests/test_oracle.py (analytic invariants of the
 * recursive filter, brute-force infinite-extension check of the Triggs-Sdika
 * boundary matrix, OpenCV cross-checks, a pure-numpy second restatement).
 *
 * Conventions (same as the reference): arrays are column-major, img[y + x*H],
 * points are 1-based (y, x) Float64 pairs, all arithmetic is Float64.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* A.1  Recursive Gaussian [3P KernelFactors.IIRGaussian / TriggsSdika]       */
/* used at pyramid.jl:105-108 (get_kernel), lucas_kanade.jl:112               */
/* ------------------------------------------------------------------------- */

typedef struct {
    double a[3];
    double scale; /* B */
    double M[9];  /* row-major 3x3 Triggs-Sdika matrix */
    double asum;
} orc_iir;

ORC_API void orc_iir_design(double sigma, orc_iir* k) {
    /* Young & van Vliet 1995 coefficients as used by ImageFiltering's IIRGaussian. */
    const double m0 = 1.16680, m1 = 1.10783, m2 = 1.40586;
    double q = 1.31564 * (sqrt(1.0 + 0.490811 * sigma * sigma) - 1.0);
    double ascale = (m0 + q) * (m1 * m1 + m2 * m2 + 2 * m1 * q + q * q);
    double B = m0 * (m1 * m1 + m2 * m2) / ascale;
    B = B * B;
    double a1 = q * (2 * m0 * m1 + m1 * m1 + m2 * m2 + (2 * m0 + 4 * m1) * q + 3 * q * q) / ascale;
    double a2 = -q * q * (m0 + 2 * m1 + 3 * q) / ascale;
    double a3 = q * q * q / ascale;
    k->a[0] = a1; k->a[1] = a2; k->a[2] = a3;
    k->scale = B;
    k->asum = a1 + a2 + a3;
    /* Triggs & Sdika 2006, eq. 15 */
    double den = (1 + a1 - a2 + a3) * (1 - a1 - a2 - a3) * (1 + a2 + (a1 - a3) * a3);
    double* M = k->M;
    M[0] = (-a3 * a1 + 1 - a3 * a3 - a2) / den;
    M[1] = ((a3 + a1) * (a2 + a3 * a1)) / den;
    M[2] = (a3 * (a1 + a3 * a2)) / den;
    M[3] = (a1 + a3 * a2) / den;
    M[4] = (-(a2 - 1) * (a2 + a3 * a1)) / den;
    M[5] = (-(a3 * a1 + a3 * a3 + a2 - 1) * a3) / den;
    M[6] = (a3 * a1 + a2 + a1 * a1 - a2 * a2) / den;
    M[7] = (a1 * a2 + a3 * a2 * a2 - a1 * a3 * a3 - a3 * a3 * a3 - a3 * a2 + a3) / den;
    M[8] = (a3 * (a1 + a3 * a2)) / den;
}

/* One line, in place, stride s, length n (n > 3 required by ImageFiltering).
 * iminus / iplus are the virtual constant inputs left / right of the line:
 * replicate border => x[first], x[last]; Fill(0) => 0. */
static void iir_line(double* x, int n, int s, const orc_iir* k, double iminus, double iplus) {
    const double a1 = k->a[0], a2 = k->a[1], a3 = k->a[2];
    double uminus = iminus / (1.0 - k->asum);
    /* forward */
    double u1 = uminus, u2 = uminus, u3 = uminus;
    for (int i = 0; i < n; ++i) {
        double u = x[(size_t)i * s] + a1 * u1 + a2 * u2 + a3 * u3;
        x[(size_t)i * s] = u;
        u3 = u2; u2 = u1; u1 = u;
    }
    /* right boundary */
    double uplus = iplus / (1.0 - k->asum);
    double vplus = uplus / (1.0 - k->asum);
    double d0 = x[(size_t)(n - 1) * s] - uplus, d1 = x[(size_t)(n - 2) * s] - uplus, d2 = x[(size_t)(n - 3) * s] - uplus;
    const double* M = k->M;
    double v1 = M[0] * d0 + M[1] * d1 + M[2] * d2 + vplus; /* v[n]   */
    double v2 = M[3] * d0 + M[4] * d1 + M[5] * d2 + vplus; /* v[n+1] */
    double v3 = M[6] * d0 + M[7] * d1 + M[8] * d2 + vplus; /* v[n+2] */
    x[(size_t)(n - 1) * s] = v1;
    for (int i = n - 2; i >= 0; --i) {
        double v = x[(size_t)i * s] + a1 * v1 + a2 * v2 + a3 * v3;
        x[(size_t)i * s] = v;
        v3 = v2; v2 = v1; v1 = v;
    }
    for (int i = 0; i < n; ++i) x[(size_t)i * s] *= k->scale;
}

/* border: 0 = replicate (imfilter! default, update! path and structure planes),
 *         1 = NA (Images.gaussian_pyramid, ctor path): zero borders, then divide by the
 *             same filter applied to ones, per dimension. */
ORC_API void orc_iir2d(const double* in, double* out, int H, int W, double sigma, int border) {
    orc_iir k;
    orc_iir_design(sigma, &k);
    if (out != in) memcpy(out, in, sizeof(double) * (size_t)H * W);
    double* ny = NULL;
    double* nx = NULL;
    if (border == 1) {
        ny = (double*)malloc(sizeof(double) * H);
        nx = (double*)malloc(sizeof(double) * W);
        for (int i = 0; i < H; ++i) ny[i] = 1.0;
        for (int i = 0; i < W; ++i) nx[i] = 1.0;
        iir_line(ny, H, 1, &k, 0.0, 0.0);
        iir_line(nx, W, 1, &k, 0.0, 0.0);
    }
    for (int x = 0; x < W; ++x) { /* dim 1 (y) */
        double* col = out + (size_t)x * H;
        if (border == 0) iir_line(col, H, 1, &k, col[0], col[H - 1]);
        else             iir_line(col, H, 1, &k, 0.0, 0.0);
    }
    for (int y = 0; y < H; ++y) { /* dim 2 (x) */
        double* row = out + y;
        if (border == 0) iir_line(row, W, H, &k, row[0], row[(size_t)(W - 1) * H]);
        else             iir_line(row, W, H, &k, 0.0, 0.0);
    }
    if (border == 1) {
        for (int x = 0; x < W; ++x)
            for (int y = 0; y < H; ++y) out[y + (size_t)x * H] /= (ny[y] * nx[x]);
        free(ny); free(nx);
    }
}

/* 1-D helper exported for the tests */
ORC_API void orc_iir1d(double* x, int n, double sigma, double iminus, double iplus) {
    orc_iir k;
    orc_iir_design(sigma, &k);
    iir_line(x, n, 1, &k, iminus, iplus);
}

/* ------------------------------------------------------------------------- */
/* A.2  Bilinear resize [3P ImageTransformations.imresize! over BSpline(Linear)] */
/* pyramid.jl:120-121,132-133                                                */
/* ------------------------------------------------------------------------- */
static inline double bilinear(const double* img, int H, int W, double r, double c) {
    /* r, c 1-based, inside [1,H]x[1,W]  [3P Interpolations.jl linear B-spline] */
    int iy = (int)floor(r), ix = (int)floor(c);
    if (iy > H - 1) iy = H - 1;
    if (ix > W - 1) ix = W - 1;
    if (iy < 1) iy = 1;
    if (ix < 1) ix = 1;
    double wy = r - iy, wx = c - ix;
    const double* p = img + (iy - 1) + (size_t)(ix - 1) * H;
    double c0 = (1.0 - wy) * p[0] + wy * p[1];
    double c1 = (1.0 - wy) * p[H] + wy * p[H + 1];
    return (1.0 - wx) * c0 + wx * c1;
}

ORC_API void orc_resize(const double* in, int H, int W, double* out, int Ho, int Wo) {
    double sy = (double)H / Ho, sx = (double)W / Wo;
    for (int j = 1; j <= Wo; ++j)
        for (int i = 1; i <= Ho; ++i) {
            double r = sy * (i - 0.5) + 0.5, c = sx * (j - 0.5) + 0.5;
            out[(i - 1) + (size_t)(j - 1) * Ho] = bilinear(in, H, W, r, c);
        }
}

/* ------------------------------------------------------------------------- */
/* A.3  Scharr gradients [3P KernelFactors.scharr]; pyramid.jl:59,75,98-103  */
/* border: 0 replicate (update!), 1 Fill(0) (ctor)                           */
/* ------------------------------------------------------------------------- */
static inline double px(const double* img, int H, int W, int y, int x, int border) {
    if (border == 1) {
        if (y < 0 || y >= H || x < 0 || x >= W) return 0.0;
    } else {
        if (y < 0) y = 0; if (y >= H) y = H - 1;
        if (x < 0) x = 0; if (x >= W) x = W - 1;
    }
    return img[y + (size_t)x * H];
}

ORC_API void orc_scharr(const double* img, int H, int W, double* Iy, double* Ix, int border) {
    for (int x = 0; x < W; ++x)
        for (int y = 0; y < H; ++y) {
            double gy = 0.0, gx = 0.0;
            static const double s[3] = {3.0 / 16.0, 10.0 / 16.0, 3.0 / 16.0};
            for (int t = -1; t <= 1; ++t) {
                gy += s[t + 1] * 0.5 * (px(img, H, W, y + 1, x + t, border) - px(img, H, W, y - 1, x + t, border));
                gx += s[t + 1] * 0.5 * (px(img, H, W, y + t, x + 1, border) - px(img, H, W, y + t, x - 1, border));
            }
            Iy[y + (size_t)x * H] = gy;
            Ix[y + (size_t)x * H] = gx;
        }
}

/* ------------------------------------------------------------------------- */
/* A.4  integral image (lucas_kanade.jl:131-138) and boxdiff [3P Images]     */
/* ------------------------------------------------------------------------- */
ORC_API void orc_integral(const double* in, double* out, int H, int W) {
    for (int x = 0; x < W; ++x) { /* cumsum dims=1 */
        double acc = 0.0;
        for (int y = 0; y < H; ++y) { acc += in[y + (size_t)x * H]; out[y + (size_t)x * H] = acc; }
    }
    for (int x = 1; x < W; ++x) /* cumsum dims=2 */
        for (int y = 0; y < H; ++y) out[y + (size_t)x * H] += out[y + (size_t)(x - 1) * H];
}

/* inclusive 1-based box rows r0..r1, cols c0..c1 */
static inline double boxdiff(const double* S, int H, int r0, int r1, int c0, int c1) {
    double sum = S[(r1 - 1) + (size_t)(c1 - 1) * H];
    if (c0 > 1) sum -= S[(r1 - 1) + (size_t)(c0 - 2) * H];
    if (r0 > 1) sum -= S[(r0 - 2) + (size_t)(c1 - 1) * H];
    if (r0 > 1 && c0 > 1) sum += S[(r0 - 2) + (size_t)(c0 - 2) * H];
    return sum;
}

/* ------------------------------------------------------------------------- */
/* LKPyramid (pyramid.jl:16-96)                                              */
/* ------------------------------------------------------------------------- */
typedef struct {
    int H0, W0, nl; /* nl = levels + 1 layers */
    int* H; int* W;
    double** layer; double** Iy; double** Ix;
    double** Iyy; double** Ixx; double** Iyx;   /* integral images of smoothed products */
    double** Syy; double** Sxx; double** Syx;   /* smoothed products before integration (`filtered`) */
    double** blur;                               /* cache.gaussian_filtered */
} orc_pyr;

static double** alloc_planes(const int* H, const int* W, int nl) {
    double** p = (double**)malloc(sizeof(double*) * nl);
    for (int i = 0; i < nl; ++i) p[i] = (double*)calloc((size_t)H[i] * W[i], sizeof(double));
    return p;
}
static void free_planes(double** p, int nl) { for (int i = 0; i < nl; ++i) free(p[i]); free(p); }

ORC_API orc_pyr* orc_pyr_create(int H, int W, int levels) {
    orc_pyr* p = (orc_pyr*)calloc(1, sizeof(orc_pyr));
    p->H0 = H; p->W0 = W; p->nl = levels + 1;
    p->H = (int*)malloc(sizeof(int) * p->nl);
    p->W = (int*)malloc(sizeof(int) * p->nl);
    p->H[0] = H; p->W[0] = W;
    for (int i = 1; i < p->nl; ++i) { p->H[i] = (p->H[i - 1] + 1) / 2; p->W[i] = (p->W[i - 1] + 1) / 2; } /* ceil(s/2) [3P pyramid_scale] */
    p->layer = alloc_planes(p->H, p->W, p->nl);
    p->Iy = alloc_planes(p->H, p->W, p->nl); p->Ix = alloc_planes(p->H, p->W, p->nl);
    p->Iyy = alloc_planes(p->H, p->W, p->nl); p->Ixx = alloc_planes(p->H, p->W, p->nl); p->Iyx = alloc_planes(p->H, p->W, p->nl);
    p->Syy = alloc_planes(p->H, p->W, p->nl); p->Sxx = alloc_planes(p->H, p->W, p->nl); p->Syx = alloc_planes(p->H, p->W, p->nl);
    p->blur = alloc_planes(p->H, p->W, p->nl);
    return p;
}

ORC_API void orc_pyr_destroy(orc_pyr* p) {
    if (!p) return;
    free_planes(p->layer, p->nl); free_planes(p->Iy, p->nl); free_planes(p->Ix, p->nl);
    free_planes(p->Iyy, p->nl); free_planes(p->Ixx, p->nl); free_planes(p->Iyx, p->nl);
    free_planes(p->Syy, p->nl); free_planes(p->Sxx, p->nl); free_planes(p->Syx, p->nl);
    free_planes(p->blur, p->nl);
    free(p->H); free(p->W); free(p);
}

/* compute_partial_derivatives! (lucas_kanade.jl:109-129) */
static void partial_derivatives(orc_pyr* p, int l) {
    int H = p->H[l], W = p->W[l];
    size_t n = (size_t)H * W;
    double* sq = (double*)malloc(sizeof(double) * n);
    const double* Iy = p->Iy[l]; const double* Ix = p->Ix[l];
    for (size_t i = 0; i < n; ++i) sq[i] = Iy[i] * Iy[i];
    orc_iir2d(sq, p->Syy[l], H, W, 4.0, 0); orc_integral(p->Syy[l], p->Iyy[l], H, W);
    for (size_t i = 0; i < n; ++i) sq[i] = Ix[i] * Ix[i];
    orc_iir2d(sq, p->Sxx[l], H, W, 4.0, 0); orc_integral(p->Sxx[l], p->Ixx[l], H, W);
    for (size_t i = 0; i < n; ++i) sq[i] = Iy[i] * Ix[i];
    orc_iir2d(sq, p->Syx[l], H, W, 4.0, 0); orc_integral(p->Syx[l], p->Iyx[l], H, W);
    free(sq);
}

/* mode 0 = update! (pyramid.jl:81-96, replicate borders everywhere)
 * mode 1 = constructor (pyramid.jl:40-79: NA-border blur via Images.gaussian_pyramid, Fill(0) Scharr) */
ORC_API void orc_pyr_build(orc_pyr* p, const double* img, double sigma, int mode) {
    memcpy(p->layer[0], img, sizeof(double) * (size_t)p->H0 * p->W0);
    for (int l = 0; l + 1 < p->nl; ++l) {
        orc_iir2d(p->layer[l], p->blur[l], p->H[l], p->W[l], sigma, mode == 1 ? 1 : 0);
        orc_resize(p->blur[l], p->H[l], p->W[l], p->layer[l + 1], p->H[l + 1], p->W[l + 1]);
    }
    for (int l = 0; l < p->nl; ++l) {
        orc_scharr(p->layer[l], p->H[l], p->W[l], p->Iy[l], p->Ix[l], mode == 1 ? 1 : 0);
        partial_derivatives(p, l);
    }
}

/* plane ids: 0 layer, 1 Iy, 2 Ix, 3 Iyy(SAT), 4 Ixx(SAT), 5 Iyx(SAT), 6 Syy, 7 Sxx, 8 Syx, 9 blur */
ORC_API const double* orc_pyr_plane(const orc_pyr* p, int level, int plane, int* H, int* W) {
    if (level < 0 || level >= p->nl) return NULL;
    *H = p->H[level]; *W = p->W[level];
    switch (plane) {
        case 0: return p->layer[level]; case 1: return p->Iy[level]; case 2: return p->Ix[level];
        case 3: return p->Iyy[level]; case 4: return p->Ixx[level]; case 5: return p->Iyx[level];
        case 6: return p->Syy[level]; case 7: return p->Sxx[level]; case 8: return p->Syx[level];
        case 9: return p->blur[level];
    }
    return NULL;
}

/* ------------------------------------------------------------------------- */
/* utils.jl:5-45  Blinn 2x2 SVD and pseudo-inverse                           */
/* ------------------------------------------------------------------------- */
static double sgn(double x) { return (x > 0) - (x < 0); }

/* M = [m11 m12; m21 m22]; returns pinv in out[4] (row-major) and singular values */
static void pinv2x2(double m11, double m12, double m21, double m22, double* out, double* s1, double* s2) {
    double E = (m11 + m22) / 2, F = (m11 - m22) / 2, G = (m21 + m12) / 2, Hh = (m21 - m12) / 2;
    double Q = sqrt(E * E + Hh * Hh), R = sqrt(F * F + G * G);
    double sx = Q + R, sy = Q - R;
    double a1 = atan2(G, F), a2 = atan2(Hh, E);
    double th = (a2 - a1) / 2, ph = (a2 + a1) / 2;
    double s = sgn(sy);
    double sp = sin(ph), cp = cos(ph), st = sin(th), ct = cos(th);
    /* U = [cp -s*sp; sp s*cp]  (column-major ctor in the reference), S = diag(sx,|sy|), V = [ct st; -st ct] */
    double U11 = cp, U21 = sp, U12 = -s * sp, U22 = s * cp;
    double V11 = ct, V21 = -st, V12 = st, V22 = ct;
    double S1 = sx, S2 = fabs(sy);
    double tol = sqrt(2.220446049250313e-16);
    double D1 = S1 > tol ? 1.0 / S1 : 0.0, D2 = S2 > tol ? 1.0 / S2 : 0.0;
    /* U * D * V' */
    out[0] = U11 * D1 * V11 + U12 * D2 * V12;
    out[1] = U11 * D1 * V21 + U12 * D2 * V22;
    out[2] = U21 * D1 * V11 + U22 * D2 * V12;
    out[3] = U21 * D1 * V21 + U22 * D2 * V22;
    *s1 = S1; *s2 = S2;
}

ORC_API void orc_pinv2x2(const double* m, double* out, double* sv) { pinv2x2(m[0], m[1], m[2], m[3], out, &sv[0], &sv[1]); }

/* ------------------------------------------------------------------------- */
/* lucas_kanade.jl:9-212  optflow!                                           */
/* ------------------------------------------------------------------------- */
typedef struct { int up, down, left, right; } offs_t;

static inline int ifloor(double v) { return (int)floor(v); }
static inline double dmin(double a, double b) { return a < b ? a : b; }
static inline double dmax(double a, double b) { return a > b ? a : b; }

/* get_offsets lucas_kanade.jl:199-208; point is the integer level coordinate, np the float correspondence */
static offs_t get_offsets(int py, int pxx, double ny, double nx, int window, int H, int W) {
    offs_t o;
    o.up    = ifloor(dmin(window, dmin(py, ny) - 1));
    o.down  = ifloor(dmin(window, H - dmax(py, ny)));
    o.left  = ifloor(dmin(window, dmin(pxx, nx) - 1));
    o.right = ifloor(dmin(window, W - dmax(pxx, nx)));
    return o;
}
static inline int offs_eq(offs_t a, offs_t b) { return a.up == b.up && a.down == b.down && a.left == b.left && a.right == b.right; }

/* compute_spatial_gradient lucas_kanade.jl:140-157; returns 0 if the window is degenerate */
static int spatial_gradient(const orc_pyr* p, int l, int py, int pxx, offs_t o, double* Ginv, double* min_eig) {
    int r0 = py - o.up, r1 = py + o.down, c0 = pxx - o.left, c1 = pxx + o.right;
    int H = p->H[l], W = p->W[l];
    if (r1 < r0 || c1 < c0 || r0 < 1 || c0 < 1 || r1 > H || c1 > W) return 0; /* Julia would throw; treated as failure */
    double syy = boxdiff(p->Iyy[l], H, r0, r1, c0, c1);
    double sxx = boxdiff(p->Ixx[l], H, r0, r1, c0, c1);
    double syx = boxdiff(p->Iyx[l], H, r0, r1, c0, c1);
    double s1, s2;
    pinv2x2(syy, syx, syx, sxx, Ginv, &s1, &s2);
    double area = (double)(r1 - r0 + 1) * (double)(c1 - c0 + 1);
    *min_eig = dmin(s1, s2) / area;
    return 1;
}

static inline int lies_in(int H, int W, double y, double x) { return 1 <= y && y <= H && 1 <= x && x <= W; }

/* disp: n x 2 (y,x) in/out at the coarsest level's scale; status: n bytes out.  Returns n_good. */
ORC_API int orc_optflow(const orc_pyr* p1, const orc_pyr* p2, const double* pts, double* disp, int n,
                        int iterations, int window, int levels, double eig_thr, double eps, uint8_t* status) {
    if (p1->nl <= levels || p2->nl <= levels) return -1; /* "Not enough layers in pyramids." lucas_kanade.jl:12-15 */
    for (int i = 0; i < n; ++i) status[i] = 1;
    for (int level = levels; level >= 0; --level) { /* 0-based level == reference level-1 */
        int H = p1->H[level], W = p1->W[level];
        const double* A = p1->layer[level]; const double* Iy = p1->Iy[level]; const double* Ix = p1->Ix[level];
        const double* B = p2->layer[level];
        double inv_scale = 1.0 / (double)(1 << level);
#pragma omp parallel for schedule(dynamic, 16)
        for (int k = 0; k < n; ++k) { /* Threads.@threads lucas_kanade.jl:33 */
            if (!status[k]) continue;
            int py = ifloor(pts[2 * k] * inv_scale), pxx = ifloor(pts[2 * k + 1] * inv_scale); /* get_pyramid_coordinate :197 */
            offs_t o = get_offsets(py, pxx, py, pxx, window, H, W);
            double Ginv[4], me;
            if (!spatial_gradient(p1, level, py, pxx, o, Ginv, &me) || me < eig_thr) { status[k] = 0; continue; }
            double cy = 0.0, cx = 0.0; /* pyramid_contribution */
            for (int it = 0; it < iterations; ++it) {
                double pcy = py + (disp[2 * k] + cy), pcx = pxx + (disp[2 * k + 1] + cx);
                if (!lies_in(H, W, pcy, pcx)) { status[k] = 0; break; }
                offs_t no = get_offsets(py, pxx, pcy, pcx, window, H, W);
                if (!offs_eq(no, o)) {
                    o = no;
                    if (!spatial_gradient(p1, level, py, pxx, o, Ginv, &me) || me < eig_thr) { status[k] = 0; break; }
                }
                /* prepare_linear_system lucas_kanade.jl:159-173 (q outer = columns, p inner = rows) */
                double by = 0.0, bx = 0.0;
                for (int q = -o.left; q <= o.right; ++q)
                    for (int pp = -o.up; pp <= o.down; ++pp) {
                        size_t idx = (size_t)(py + pp - 1) + (size_t)(pxx + q - 1) * H;
                        double dI = A[idx] - bilinear(B, H, W, pcy + pp, pcx + q);
                        by += dI * Iy[idx];
                        bx += dI * Ix[idx];
                    }
                double fy = Ginv[0] * by + Ginv[1] * bx, fx = Ginv[2] * by + Ginv[3] * bx;
                if (fabs(fy) < eps && fabs(fx) < eps) break;
                cy += fy; cx += fx;
                if (!lies_in(H, W, pcy + fy, pcx + fx)) { status[k] = 0; break; }
            }
            if (!status[k]) continue;
            disp[2 * k] += cy; disp[2 * k + 1] += cx;
            if (level > 0) { disp[2 * k] *= 2.0; disp[2 * k + 1] *= 2.0; }
        }
    }
    int good = 0;
    for (int i = 0; i < n; ++i) good += status[i];
    return good;
}

/* ------------------------------------------------------------------------- */
/* tracker.jl:17-68  fb_tracking!                                            */
/* disp_in may be NULL (zeros).  out_pts[i] written only where forward ok.   */
/* status bit0 = final status, bit1 = forward status                         */
/* ------------------------------------------------------------------------- */
ORC_API int orc_fb_track(const orc_pyr* prev, const orc_pyr* cur, const double* pts, const double* disp_in, int n,
                         int iterations, int window, int levels, double eig_thr, double eps, double max_distance,
                         double* out_pts, uint8_t* status) {
    if (n == 0) return 0;
    double* disp = (double*)calloc((size_t)2 * n, sizeof(double));
    if (disp_in) memcpy(disp, disp_in, sizeof(double) * 2 * n);
    uint8_t* st = (uint8_t*)malloc(n);
    int ngood = orc_optflow(prev, cur, pts, disp, n, iterations, window, levels, eig_thr, eps, st);
    if (ngood < 0) { free(disp); free(st); return -1; }
    int* ids = (int*)malloc(sizeof(int) * (ngood + 1));
    double* vc = (double*)malloc(sizeof(double) * 2 * (ngood + 1));
    double* bd = (double*)malloc(sizeof(double) * 2 * (ngood + 1));
    uint8_t* bst = (uint8_t*)malloc(ngood + 1);
    int c = 0;
    for (int i = 0; i < n; ++i) {
        status[i] = 0;
        if (!st[i]) continue;
        double ny = pts[2 * i] + disp[2 * i], nx = pts[2 * i + 1] + disp[2 * i + 1];
        out_pts[2 * i] = ny; out_pts[2 * i + 1] = nx;
        vc[2 * c] = ny; vc[2 * c + 1] = nx;
        bd[2 * c] = -disp[2 * i]; bd[2 * c + 1] = -disp[2 * i + 1]; /* scale = 1 since back levels = 0 (tracker.jl:34-35) */
        ids[c++] = i;
        status[i] = 3;
    }
    /* backward: pyramid_levels = 0, eps = LucasKanade default 1e-2 (tracker.jl:51-57 does not forward eps) */
    orc_optflow(cur, prev, vc, bd, c, iterations, window, 0, eig_thr, 1e-2, bst);
    int final_good = 0;
    for (int j = 0; j < c; ++j) {
        int i = ids[j];
        if (!bst[j]) { status[i] = 2; continue; }
        double by = vc[2 * j] + bd[2 * j], bx = vc[2 * j + 1] + bd[2 * j + 1];
        double dy = pts[2 * i] - by, dx = pts[2 * i + 1] - bx;
        if (sqrt(dy * dy + dx * dx) >= max_distance) { status[i] = 2; continue; }
        ++final_good;
    }
    free(disp); free(st); free(ids); free(vc); free(bd); free(bst);
    return final_good;
}

/* ------------------------------------------------------------------------- */
/* extractor.jl:24-122                                                       */
/* ------------------------------------------------------------------------- */
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* [3P Images.shi_tomasi defaults]: Sobel/8 gradients, 3x3 mean of products, replicate border,
 * all evaluated on the cell sub-image (h x w view with leading dimension ld). */
static void shi_tomasi_cell(const double* cell, int ld, int h, int w, double* R, double* g /* 3*h*w scratch */) {
    double* gyy = g; double* gyx = g + (size_t)h * w; double* gxx = g + 2 * (size_t)h * w;
#define CP(y, x) cell[clampi(y, 0, h - 1) + (size_t)clampi(x, 0, w - 1) * ld]
    for (int x = 0; x < w; ++x)
        for (int y = 0; y < h; ++y) {
            double gy = ((CP(y + 1, x - 1) - CP(y - 1, x - 1)) + 2.0 * (CP(y + 1, x) - CP(y - 1, x)) + (CP(y + 1, x + 1) - CP(y - 1, x + 1))) / 8.0;
            double gx = ((CP(y - 1, x + 1) - CP(y - 1, x - 1)) + 2.0 * (CP(y, x + 1) - CP(y, x - 1)) + (CP(y + 1, x + 1) - CP(y + 1, x - 1))) / 8.0;
            gyy[y + (size_t)x * h] = gy * gy; gyx[y + (size_t)x * h] = gy * gx; gxx[y + (size_t)x * h] = gx * gx;
        }
#undef CP
    for (int x = 0; x < w; ++x)
        for (int y = 0; y < h; ++y) {
            /* meancovs: imfilter with the dense kernel (1/9) * ones(3,3): products accumulated tap by tap [3P] */
            const double k9 = 1.0 / 9.0;
            double a = 0, b = 0, c = 0;
            for (int dx = -1; dx <= 1; ++dx)
                for (int dy = -1; dy <= 1; ++dy) {
                    size_t i = clampi(y + dy, 0, h - 1) + (size_t)clampi(x + dx, 0, w - 1) * h;
                    a += k9 * gyy[i]; b += k9 * gyx[i]; c += k9 * gxx[i];
                }
            R[y + (size_t)x * h] = ((a + c) - sqrt((a - c) * (a - c) + 4.0 * b * b)) / 2.0;
        }
}

ORC_API void orc_shi_tomasi(const double* img, int H, int W, double* R) {
    double* g = (double*)malloc(sizeof(double) * 3 * (size_t)H * W);
    shi_tomasi_cell(img, H, H, W, R, g);
    free(g);
}

/* get_mask (extractor.jl:116-122) + Kernel.gaussian(sigma) blur (extractor.jl:69) [3P].
 * Disc rasterisation [3P ImageDraw 0.2]: draw!(img, CirclePointRadius) goes through the filled Ellipse, which sets pixel (i, j) iff
 * ((i - cy) / r)^2 + ((j - cx) / r)^2 < 1 in Float64 (strict: lattice points on the circle stay outside, up to the rounding of
 * that expression, which is reproduced as written). */
ORC_API void orc_mask(int H, int W, const double* pts, int n, int radius, double sigma, double* mask) {
    size_t N = (size_t)H * W;
    for (size_t i = 0; i < N; ++i) mask[i] = 1.0;
    for (int k = 0; k < n; ++k) {
        int cy = (int)nearbyint(pts[2 * k]), cx = (int)nearbyint(pts[2 * k + 1]); /* Julia round = ties-to-even */
        for (int x = cx - radius; x <= cx + radius; ++x)
            for (int y = cy - radius; y <= cy + radius; ++y) {
                if (y < 1 || y > H || x < 1 || x > W) continue;
                int dy = y - cy, dx = x - cx;
                double vy = (double)dy / (double)radius, vx = (double)dx / (double)radius;
                if (vy * vy + vx * vx < 1.0) mask[(y - 1) + (size_t)(x - 1) * H] = 0.0;
            }
    }
    if (!(sigma > 0)) return;
    int hw = 2 * (int)ceil(sigma); /* length 4*ceil(sigma)+1 */
    int len = 2 * hw + 1;
    double* k1 = (double*)malloc(sizeof(double) * len);
    double ks = 0;
    for (int i = -hw; i <= hw; ++i) { k1[i + hw] = exp(-(double)i * i / (2 * sigma * sigma)); ks += k1[i + hw]; }
    for (int i = 0; i < len; ++i) k1[i] /= ks;
    double* tmp = (double*)malloc(sizeof(double) * N);
    for (int x = 0; x < W; ++x)
        for (int y = 0; y < H; ++y) {
            double acc = 0;
            for (int t = -hw; t <= hw; ++t) acc += k1[t + hw] * mask[clampi(y + t, 0, H - 1) + (size_t)x * H];
            tmp[y + (size_t)x * H] = acc;
        }
    for (int x = 0; x < W; ++x)
        for (int y = 0; y < H; ++y) {
            double acc = 0;
            for (int t = -hw; t <= hw; ++t) acc += k1[t + hw] * tmp[y + (size_t)clampi(x + t, 0, W - 1) * H];
            mask[y + (size_t)x * H] = acc;
        }
    free(tmp); free(k1);
}

typedef struct { double r; int idx; } cand_t;

/* detect (extractor.jl:63-95).  out_yx: pairs of int64 (y,x) 1-based.  Returns count (may exceed max_points; no global cap). */
ORC_API int orc_detect(const double* image, int H, int W, const double* cur_pts, int n_cur,
                       int max_points, int radius, int grid_h, int grid_w, int cell_size, double sigma_mask,
                       double min_response, int64_t* out_yx, int cap) {
    if (n_cur >= max_points) return 0;
    size_t N = (size_t)H * W;
    double* img = (double*)malloc(sizeof(double) * N);
    memcpy(img, image, sizeof(double) * N);
    if (n_cur > 0) {
        double* mask = (double*)malloc(sizeof(double) * N);
        orc_mask(H, W, cur_pts, n_cur, radius, sigma_mask, mask);
        for (size_t i = 0; i < N; ++i) img[i] *= mask[i];
        free(mask);
    }
    int n_cells = grid_h * grid_w;
    int n_detect = max_points - n_cur;
    int k_cell = (n_detect + n_cells - 1) / n_cells; /* ceil */
    int cs = cell_size;
    double* R = (double*)malloc(sizeof(double) * cs * cs);
    double* g = (double*)malloc(sizeof(double) * 3 * cs * cs);
    cand_t* cand = (cand_t*)malloc(sizeof(cand_t) * cs * cs);
    uint8_t* corner = (uint8_t*)malloc(cs * cs);
    int count = 0;
    for (int gy = 0; gy < grid_h; ++gy)
        for (int gx = 0; gx < grid_w; ++gx) {
            int y0 = gy * cs, x0 = gx * cs;
            int y1 = (gy + 1) * cs < H ? (gy + 1) * cs : H, x1 = (gx + 1) * cs < W ? (gx + 1) * cs : W;
            int h = y1 - y0, w = x1 - x0;
            if (h <= 0 || w <= 0) continue;
            shi_tomasi_cell(img + y0 + (size_t)x0 * H, H, h, w, R, g);
            /* findlocalmaxima [3P]: strictly greater than every in-bounds 8-neighbour, column-major enumeration */
            int nc = 0;
            for (int x = 0; x < w; ++x)
                for (int y = 0; y < h; ++y) {
                    double r = R[y + x * h];
                    int ismax = 1;
                    for (int dx = -1; dx <= 1 && ismax; ++dx)
                        for (int dy = -1; dy <= 1; ++dy) {
                            if (!dx && !dy) continue;
                            int yy = y + dy, xx = x + dx;
                            if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
                            if (!(R[yy + xx * h] < r)) { ismax = 0; break; }
                        }
                    if (ismax) { cand[nc].r = r; cand[nc].idx = y + x * h; ++nc; }
                }
            /* stable sort descending by response (insertion sort keeps enumeration order on ties) */
            for (int i = 1; i < nc; ++i) {
                cand_t t = cand[i]; int j = i - 1;
                while (j >= 0 && cand[j].r < t.r) { cand[j + 1] = cand[j]; --j; }
                cand[j + 1] = t;
            }
            int keep = nc < k_cell ? nc : k_cell;
            memset(corner, 0, (size_t)h * w);
            for (int i = 0; i < keep; ++i) if (!(cand[i].r < min_response)) corner[cand[i].idx] = 1;
            for (int i = 0; i < h * w; ++i) /* Keypoints(sub_features): column-major findall */
                if (corner[i]) {
                    if (count < cap) { out_yx[2 * count] = (i % h) + 1 + y0; out_yx[2 * count + 1] = (i / h) + 1 + x0; }
                    ++count;
                }
        }
    free(img); free(R); free(g); free(cand); free(corner);
    return count;
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
ORC_API void orc_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
