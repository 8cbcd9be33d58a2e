/* C99 consumer of include/slamklt.h: proves the header is self-contained C, pins the struct layouts the Julia shim and the
 * ctypes mirror assume, and drives the library without Python: context, 64 x 64 pyramid (ctor + update!), one tracked point,
 * one detect call.  Built by tests/test_abi_c.py:
 *     gcc -std=c99 -Wall -Werror -Iinclude tests/abi_smoke.c -o <tmp>/abi_smoke -L<csrc> -lslamklt -Wl,-rpath,<csrc> -lm
 * Exit code 0 = ok, 2 = layout mismatch (compile-time), 3 = no device (the CPU-only run stops after the layout checks), 1 = failure. */
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "slamklt.h"

/* C99 has no _Static_assert: negative array size trick */
#define STATIC_CHECK(name, cond) typedef char static_check_##name[(cond) ? 1 : -1]
STATIC_CHECK(lk_params_size, sizeof(slamklt_lk_params) == 40);
STATIC_CHECK(lk_params_eig, offsetof(slamklt_lk_params, eigenvalue_threshold) == 16);
STATIC_CHECK(lk_params_maxd, offsetof(slamklt_lk_params, max_distance) == 32);
STATIC_CHECK(det_params_size, sizeof(slamklt_detect_params) == 40);
STATIC_CHECK(det_params_sigma, offsetof(slamklt_detect_params, sigma_mask) == 24);
STATIC_CHECK(camera_size, sizeof(slamklt_camera) == 208);
STATIC_CHECK(camera_height, offsetof(slamklt_camera, height) == 64);
STATIC_CHECK(camera_ti0, offsetof(slamklt_camera, Ti0) == 80);
STATIC_CHECK(matching_size, sizeof(slamklt_matching_params) == 56);
STATIC_CHECK(matching_stereo, offsetof(slamklt_matching_params, stereo) == 40);
STATIC_CHECK(stats_size, sizeof(slamklt_stats) == 40);

#define CHECK(call)                                                                            \
    do {                                                                                       \
        int rc_ = (call);                                                                      \
        if (rc_ != 0) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, slamklt_last_error()); return 1; } \
    } while (0)

int main(int argc, char** argv) {
    const int H = 64, W = 64, levels = 2;
    slamklt_ctx* ctx = NULL;
    slamklt_pyr *a = NULL, *b = NULL;
    double *img0, *img1, pts[2], out[2], disp[2] = {0.0, 0.0};
    uint8_t status = 0;
    int i, x, y, rc, n_out = 0, built = 0, hh = 0, ww = 0, ll = 0;
    int64_t kp[2 * 64];
    slamklt_lk_params lk;
    slamklt_detect_params dp;
    (void)argv;
    if (slamklt_version() != SLAMKLT_VERSION) { fprintf(stderr, "version mismatch\n"); return 1; }
    if (argc > 1 && strcmp(argv[1], "--layout-only") == 0) { printf("abi_smoke: layouts ok, version %d\n", slamklt_version()); return 0; }
    rc = slamklt_ctx_create(0, &ctx);
    if (rc == SLAMKLT_E_NODEVICE) { fprintf(stderr, "no device: %s\n", slamklt_last_error()); return 3; }
    if (rc != 0) { fprintf(stderr, "ctx_create -> %d: %s\n", rc, slamklt_last_error()); return 1; }

    /* a smooth blob pattern, column-major like Julia's Matrix; frame 1 = frame 0 shifted by (+1.5, -0.75) px */
    img0 = (double*)malloc(sizeof(double) * H * W);
    img1 = (double*)malloc(sizeof(double) * H * W);
    for (x = 0; x < W; ++x)
        for (y = 0; y < H; ++y) {
            img0[y + x * H] = 0.5 + 0.25 * sin(0.31 * y) * cos(0.23 * x) + 0.2 * exp(-((y - 30.0) * (y - 30.0) + (x - 33.0) * (x - 33.0)) / 40.0);
            img1[y + x * H] = 0.5 + 0.25 * sin(0.31 * (y - 1.5)) * cos(0.23 * (x + 0.75)) +
                              0.2 * exp(-((y - 31.5) * (y - 31.5) + (x - 32.25) * (x - 32.25)) / 40.0);
        }
    CHECK(slamklt_pyr_create(ctx, H, W, levels, &a));
    CHECK(slamklt_pyr_create(ctx, H, W, levels, &b));
    CHECK(slamklt_pyr_build(ctx, a, img0, SLAMKLT_F64, H, 1.0, SLAMKLT_MODE_CTOR));
    CHECK(slamklt_pyr_build(ctx, b, img1, SLAMKLT_F64, H, 1.0, SLAMKLT_MODE_CTOR));
    CHECK(slamklt_pyr_build(ctx, b, img1, SLAMKLT_F64, H, 1.0, SLAMKLT_MODE_UPDATE));
    CHECK(slamklt_pyr_info(b, &hh, &ww, &ll, &built));
    if (hh != H || ww != W || ll != levels || !built) { fprintf(stderr, "pyr_info mismatch\n"); return 1; }

    memset(&lk, 0, sizeof(lk));
    lk.iterations = 30; lk.window_size = 9; lk.pyramid_levels = levels;
    lk.eigenvalue_threshold = 1e-4; lk.epsilon = 1e-2; lk.max_distance = 1.0;
    pts[0] = 31.0; pts[1] = 33.0;
    out[0] = out[1] = -1.0;
    CHECK(slamklt_fb_track(ctx, a, b, pts, disp, 1, &lk, out, &status));
    printf("abi_smoke: tracked (%.3f, %.3f) -> (%.3f, %.3f), status %d\n", pts[0], pts[1], out[0], out[1], (int)status);
    if (!(status & 1) || fabs(out[0] - pts[0] - 1.5) > 0.2 || fabs(out[1] - pts[1] + 0.75) > 0.2) { fprintf(stderr, "track off\n"); return 1; }
    /* too many levels must be refused with the reference's message */
    lk.pyramid_levels = levels + 1;
    rc = slamklt_fb_track(ctx, a, b, pts, disp, 1, &lk, out, &status);
    if (rc != SLAMKLT_E_LAYERS || strstr(slamklt_last_error(), "Not enough layers") == NULL) { fprintf(stderr, "layers check\n"); return 1; }

    memset(&dp, 0, sizeof(dp));
    dp.max_points = 16; dp.radius = 5; dp.grid_h = 2; dp.grid_w = 2; dp.cell_size = 32; dp.sigma_mask = 3.0; dp.min_response = 1e-4;
    CHECK(slamklt_detect(ctx, img0, SLAMKLT_F64, H, W, H, NULL, 0, &dp, kp, 64, &n_out));
    printf("abi_smoke: detect -> %d keypoints", n_out);
    for (i = 0; i < n_out && i < 3; ++i) printf(" (%lld, %lld)", (long long)kp[2 * i], (long long)kp[2 * i + 1]);
    printf("\n");
    if (n_out < 1 || n_out > 16) { fprintf(stderr, "detect count\n"); return 1; }

    CHECK(slamklt_pyr_destroy(ctx, a));
    CHECK(slamklt_pyr_destroy(ctx, b));
    CHECK(slamklt_ctx_destroy(ctx));
    free(img0); free(img1);
    printf("abi_smoke: ok\n");
    return 0;
}
