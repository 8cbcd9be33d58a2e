/*
 * slamklt.h -- C ABI of libslamklt.so: the B200 (sm_100a) implementation of SLAM.jl's KLT
 * front-end hot path (image pyramid, pyramidal Lucas-Kanade with forward-backward check,
 * Shi-Tomasi extraction with NMS and grid bucketing).
 *
 * The reference reaches this path by ordinary Julia method dispatch (it has no FFI today);
 * each entry point below names the reference method it stands behind (file:line relative to
 * the SLAM.jl checkout).  INTEGRATION.md shows the `ccall` shim that forwards those methods.
 *
 * Conventions are the reference's own:
 *   - images are column-major, element (y, x) at img[y + x*ld] (Julia Matrix{Gray{Float64}}, ld = H);
 *   - points are dense pairs of Float64 in (y, x) order, 1-based (Vector{SVector{2,Float64}});
 *   - extracted keypoints are dense pairs of Int64 (y, x), 1-based (Vector{CartesianIndex{2}}).
 * Every function returns 0 on success or a negative SLAMKLT_E_* code; the message is available
 * from slamklt_last_error() (thread-local).  No exceptions cross the boundary.  All pointers
 * are caller-owned HOST pointers unless the name says `_dev`; the library copies.
 * A context owns one CUDA stream; calls on one context are serialised by an internal mutex, so
 * two Julia tasks (front-end and mapper, mapper.jl:51-60) may share it or use one context each.
 */
#ifndef SLAMKLT_H
#define SLAMKLT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define SLAMKLT_VERSION 100

/* error codes */
#define SLAMKLT_OK 0
#define SLAMKLT_E_INVALID (-1)    /* bad argument */
#define SLAMKLT_E_CUDA (-2)       /* CUDA runtime error (sticky errors included) */
#define SLAMKLT_E_LAYERS (-3)     /* "Not enough layers in pyramids." lucas_kanade.jl:12-15 */
#define SLAMKLT_E_NODEVICE (-4)   /* no usable sm_100 device; there is no CPU fallback */
#define SLAMKLT_E_CAPACITY (-5)   /* output buffer too small */

/* pixel types accepted at the boundary */
#define SLAMKLT_F64 0 /* Gray{Float64}: what the reference passes (SLAM.jl:22-26) */
#define SLAMKLT_F32 1
#define SLAMKLT_U8 2  /* value/255, i.e. Gray{N0f8} before the example's Gray{Float64}.() conversion */

/* pyramid build modes */
#define SLAMKLT_MODE_UPDATE 0 /* update!(pyr, img): replicate borders (pyramid.jl:81-103) */
#define SLAMKLT_MODE_CTOR 1   /* LKPyramid(img, levels): NA-border blur, Fill(0) Scharr (pyramid.jl:40-79) */

/* planes for slamklt_pyr_download (names follow LKPyramid / LKCache fields, pyramid.jl:1-24) */
#define SLAMKLT_PLANE_LAYER 0
#define SLAMKLT_PLANE_IY 1
#define SLAMKLT_PLANE_IX 2
#define SLAMKLT_PLANE_IYY 3 /* integral image of the smoothed Iy*Iy; rebuilt in Float64 on download */
#define SLAMKLT_PLANE_IXX 4
#define SLAMKLT_PLANE_IYX 5
#define SLAMKLT_PLANE_SYY 6 /* smoothed products before integration (LKCache.filtered) */
#define SLAMKLT_PLANE_SXX 7
#define SLAMKLT_PLANE_SYX 8
#define SLAMKLT_PLANE_BLUR 9 /* LKCache.gaussian_filtered (levels 0..L-1) */
/* The device keeps the smoothed products as exclusive prefix sums along x (fp32, H_l x (W_l + 1), column 0 = 0) instead of the
 * reference's 2-D integral images (lucas_kanade.jl:131-138); these three return exactly the planes the tracking kernel reads,
 * so out must hold H_l * (W_l + 1) doubles.  A window's row sum is R[y, c1 + 1] - R[y, c0] (0-based columns c0..c1). */
#define SLAMKLT_PLANE_RYY 10
#define SLAMKLT_PLANE_RXX 11
#define SLAMKLT_PLANE_RYX 12

typedef struct slamklt_ctx slamklt_ctx;
typedef struct slamklt_pyr slamklt_pyr;
typedef struct slamklt_batch slamklt_batch;

/* LucasKanade (lucas_kanade.jl:1-7) + fb_tracking!'s max_distance (tracker.jl:22) */
typedef struct slamklt_lk_params {
    int32_t iterations;          /* 30 */
    int32_t window_size;         /* half width; 9 in Params (params.jl:65); 1 .. 255 */
    int32_t pyramid_levels;      /* 3 */
    int32_t reserved;
    double eigenvalue_threshold; /* 1e-4 */
    double epsilon;              /* 1e-2 */
    double max_distance;         /* params.max_ktl_distance = 1.0; wrapper default 0.5 */
} slamklt_lk_params;

/* Extractor (extractor.jl:7-22) + detect kwargs (extractor.jl:24,63) */
typedef struct slamklt_detect_params {
    int32_t max_points;
    int32_t radius;
    int32_t grid_h, grid_w; /* grid_resolution */
    int32_t cell_size;      /* <= 64 */
    int32_t reserved;
    double sigma_mask;      /* 3.0; 0 skips the mask blur */
    double min_response;    /* 1e-4 */
} slamklt_detect_params;

/* per-context timing / counters of the last call(s), filled by slamklt_get_stats */
typedef struct slamklt_stats {
    uint64_t kernel_launches; /* kernels launched by this context since creation */
    uint64_t lk_window_iters; /* executed LK iterations x window pixels (device counter), since last reset */
    uint64_t lk_iters;        /* executed LK iterations, since last reset */
    uint64_t h2d_bytes, d2h_bytes; /* since creation */
} slamklt_stats;

const char* slamklt_last_error(void);
int slamklt_version(void);
int slamklt_device_count(void);

/* ---- context --------------------------------------------------------------------------- */
int slamklt_ctx_create(int device, slamklt_ctx** out);
int slamklt_ctx_destroy(slamklt_ctx* ctx);
int slamklt_ctx_sync(slamklt_ctx* ctx);
int slamklt_get_stats(slamklt_ctx* ctx, slamklt_stats* out, int reset_lk_counters);
/* CUDA-event stopwatch on the context's own stream (the stream every kernel here is launched on) */
int slamklt_timer_start(slamklt_ctx* ctx);
int slamklt_timer_stop(slamklt_ctx* ctx, float* elapsed_ms); /* records, synchronises, returns ms */

/* per-kernel device timing (CUDA events around every launch on the context's stream).  Off by default; never
 * enable it inside a region you time.  The report is text: one "kernel_name launches total_ms" line per kernel. */
int slamklt_profile(slamklt_ctx* ctx, int enable);
int slamklt_profile_report(slamklt_ctx* ctx, char* buf, size_t cap);

/* ---- LKPyramid ------------------------------------------------------------------------- */
/* allocation of LKPyramid(image, levels; reusable=true) -- pyramid.jl:40-72.
 * H x W from 4 x 4 to 16384 x 16384, every level at least 4 x 4 (the recursive filter needs more than 3 samples per line),
 * levels + 1 <= 8.  Levels of up to 1088 rows (level 0: 1280) and 2048 columns are built by the register-tiled kernels,
 * larger ones by general per-line kernels (same planes, same tolerances, slower). */
int slamklt_pyr_create(slamklt_ctx* ctx, int H, int W, int levels, slamklt_pyr** out);
int slamklt_pyr_destroy(slamklt_ctx* ctx, slamklt_pyr* pyr);
/* LKPyramid ctor (mode CTOR, pyramid.jl:40-79) and update!(lk, img; sigma) (mode UPDATE, pyramid.jl:81-96) */
int slamklt_pyr_build(slamklt_ctx* ctx, slamklt_pyr* pyr, const void* img, int dtype, int ld, double sigma, int mode);
/* Base.copy!(dst::LKPyramid, src) -- pyramid.jl:28-38 */
int slamklt_pyr_copy(slamklt_ctx* ctx, slamklt_pyr* dst, const slamklt_pyr* src);
/* deepcopy(front_end.current_pyramid) -- SLAM.jl:216-219 */
int slamklt_pyr_clone(slamklt_ctx* ctx, const slamklt_pyr* src, slamklt_pyr** out);
/* O(1) replacement for copy!(previous, current) followed by update!(current, img) -- front_end.jl:459-461 */
int slamklt_pyr_swap(slamklt_ctx* ctx, slamklt_pyr* a, slamklt_pyr* b);
int slamklt_pyr_info(const slamklt_pyr* pyr, int* H, int* W, int* levels, int* built /* has_gradients, pyramid.jl:26 */);
int slamklt_pyr_level_dims(const slamklt_pyr* pyr, int level, int* H, int* W);
/* field access lk.layers[l+1], lk.Iy[l+1], ... as Float64 column-major H_l x W_l (level is 0-based) */
int slamklt_pyr_download(slamklt_ctx* ctx, const slamklt_pyr* pyr, int level, int plane, double* out);

/* ---- Lucas-Kanade ---------------------------------------------------------------------- */
/* optflow!(displacement, first, second, points, algorithm) -- lucas_kanade.jl:9-100.
 * disp_inout (n x 2, coarsest-level scale) is updated in place like the reference; status[i] in {0,1};
 * *n_good receives the number of surviving points. */
int slamklt_optflow(slamklt_ctx* ctx, const slamklt_pyr* first, const slamklt_pyr* second, const double* pts_yx,
                    double* disp_inout, int n, const slamklt_lk_params* p, uint8_t* status, int* n_good);
/* fb_tracking!(new_keypoints, previous, current, keypoints, algorithm; displacement, max_distance)
 * -- tracker.jl:17-68.  disp_yx may be NULL (zeros).  out_pts_yx[i] is written only where the forward
 * pass succeeded (as in the reference).  status bit0 = returned status, bit1 = forward-pass status. */
int slamklt_fb_track(slamklt_ctx* ctx, const slamklt_pyr* previous, const slamklt_pyr* current, const double* pts_yx,
                     const double* disp_yx, int n, const slamklt_lk_params* p, double* out_pts_yx, uint8_t* status);

/* Tracking part of optical_flow_matching!(map_manager, frame, from, to, stereo) -- map_manager.jl:451-564,
 * in one launch: keypoints with has_prior[i] != 0 (3-D keypoints whose map point projects into the image) are first
 * tracked with prior_disp_yx[i] (= (projection - pixel) / 2^levels_3d, map_manager.jl:466,494) on levels_3d levels (the
 * reference hard-codes 1, map_manager.jl:458); those that fail, and all others, are tracked from a zero displacement on
 * p->pyramid_levels levels (map_manager.jl:531-551).  status bit0 = tracked, bit1 = forward pass of the deciding attempt ok,
 * bit2 = tracked by the prior pass.  out_pts_yx[i] is written only where bit1 is set. */
int slamklt_flow_matching(slamklt_ctx* ctx, const slamklt_pyr* from, const slamklt_pyr* to, const double* pts_yx,
                          const double* prior_disp_yx, const uint8_t* has_prior, int n, const slamklt_lk_params* p,
                          int levels_3d, double* out_pts_yx, uint8_t* status);

/* Camera -- camera.jl:1-29, the fields project_undistort / undistort_point / backproject / in_image read. */
typedef struct slamklt_camera {
    double fx, fy, cx, cy;  /* camera.jl:3-7 */
    double k1, k2, p1, p2;  /* camera.jl:9-13 */
    int64_t height, width;  /* camera.jl:18-19 */
    double Ti0[16];         /* camera.jl:21-24, 4 x 4 column-major (SMatrix storage order) */
} slamklt_camera;

typedef struct slamklt_matching_params {
    slamklt_lk_params lk;       /* window_size, pyramid_levels, max_distance = params.max_ktl_distance (map_manager.jl:454-456) */
    int32_t stereo;             /* the `stereo` argument (map_manager.jl:452) */
    int32_t pyramid_levels_3d;  /* map_manager.jl:458 (1) */
    double epipolar_error;      /* maybe_stereo_update!, map_manager.jl:580 (2.0) */
} slamklt_matching_params;

/* optical_flow_matching!(map_manager, frame, from, to, stereo) -- map_manager.jl:451-564 -- with its per-keypoint geometry
 * on the device: for every keypoint of `frame` (pixels_yx = kp.pixel, is_3d = kp.is_3d and its map
 * point exists, world_xyz = get_position(mp), undist_yx = kp.undistorted_pixel, only read in stereo mode)
 *   1. 3-D keypoints are projected with frame.cw (column-major 4 x 4; stereo: right_cam->Ti0 * cw) through
 *      project_undistort(cam, .) (frame.jl:478-484, camera.jl:73-85); inside the image (camera.jl:87-95; stereo: the right
 *      camera's bounds) they get the prior (projection - pixel) / 2^pyramid_levels_3d, outside they are not tracked (bit 3);
 *   2. tracking as in slamklt_flow_matching (prior pass, retry of the failures together with the 2-D keypoints);
 *   3. mono: update_keypoint! (frame.jl:252-270): out_pixel = tracked pixel, out_undist = undistort_point(cam, pixel),
 *      out_position = backproject(cam, undist) as (x, y, 1);
 *      stereo: maybe_stereo_update! (map_manager.jl:579-590) + update_stereo_keypoint! (frame.jl:272-287): rejected when
 *      |undist_yx[i].y - undistort_point(right_cam, tracked).y| > epipolar_error (bit 4), else out_pixel = (pixel.y, tracked.x),
 *      out_undist / out_position through the right camera.
 * status: bit0 = keypoint updated, bit1 = forward pass of the deciding attempt ok, bit2 = tracked by the prior pass,
 * bit3 = 3-D keypoint whose projection is outside the image (mono: left untouched, stereo: remove_mappoint_obs!),
 * bit4 = tracked but rejected by the epipolar gate.  Outputs of keypoints without bit0 are NaN.  The caller keeps the
 * dictionary bookkeeping: bit0 -> store the three outputs, otherwise remove_obs_from_current_frame! (mono, bit3 clear). */
int slamklt_optical_flow_matching(slamklt_ctx* ctx, const slamklt_pyr* from, const slamklt_pyr* to, const double* pixels_yx,
                                  const uint8_t* is_3d, const double* world_xyz, const double* undist_yx, int n,
                                  const double* cw, const slamklt_camera* cam, const slamklt_camera* right_cam,
                                  const slamklt_matching_params* p, double* out_pixel_yx, double* out_undist_yx,
                                  double* out_position_xyz, uint8_t* status);

/* triangulate_stereo!(map_manager, frame, max_error, cache) -- mapper.jl:142-183 -- for the stereo keypoints the caller selected
 * (not yet 3-D, map point present: mapper.jl:157-161 stays host-side bookkeeping).  Per keypoint: DLT triangulation from
 * und_yx = kp.undistorted_pixel and rund_yx = kp.right_undistorted_pixel with P1 = K, P2 = K_right * right_cam->Ti0
 * (RecoverPose.triangulate: eigenvector of A'A for the smallest eigenvalue, normalised by its 4th component), depth >= 0.1 in
 * both cameras, reprojection errors <= max_error (params.max_reprojection_error = 3.0), world point = wc * point (wc =
 * frame.wc, column-major 4 x 4).  status: 1 = update_mappoint! with out_world_xyz[i]; 2 / 3 = remove_stereo_keypoint! (depth in
 * the left / right camera), 4 / 5 = remove_stereo_keypoint! (reprojection error left / right); out_world_xyz is NaN unless 1.
 * Agreement with a LAPACK eigen-solve: 1e-9 relative on the world point (tests/test_gpu_configs.py). */
int slamklt_triangulate_stereo(slamklt_ctx* ctx, const double* und_yx, const double* rund_yx, int n, const slamklt_camera* cam,
                               const slamklt_camera* right_cam, const double* wc, double max_error, double* out_world_xyz,
                               uint8_t* status);

/* ---- BRIEF descriptors + Hamming matching (only with params.do_local_matching, params.jl:69) ------------------------- */
/* describe(e, image, keypoints) -- extractor.jl:103-105 -> ImageFeatures.create_descriptor(image, keypoints, BRIEF(size = 256)).
 * kps_yx: n pairs of Int64 (y, x), 1-based (Vector{CartesianIndex{2}}).  pairs: n_bits x (dy1, dx1, dy2, dx2) Int32 -- the
 * sampling pattern ImageFeatures draws from Julia's RNG (Random.seed!(123); gaussian(n_bits, window)); it cannot be regenerated
 * outside Julia, the shim exports it once.  The image is smoothed with the 4*ceil(sigma)+1 tap Gaussian (sigma = sqrt(2) in
 * BRIEF's defaults, replicate border, Float64) around each keypoint; bit b = smoothed[k + s1[b]] < smoothed[k + s2[b]], packed
 * LSB first into n_bits/32 UInt32 per keypoint (same order as the chunks of a BitVector).  out_valid[i] = 0 for keypoints closer
 * than ceil(window / 2) to the border: the reference drops those from the returned lists, the caller compacts. */
int slamklt_describe(slamklt_ctx* ctx, const void* img, int dtype, int H, int W, int ld, const int64_t* kps_yx, int n,
                     const int32_t* pairs, int n_bits, int window, double sigma, uint32_t* out_desc, uint8_t* out_valid);
/* Descriptor side of find_best_match (mapper.jl:392-462) for many target map points at once.  Map point s owns the descriptor
 * rows [set_off[s], set_off[s+1]) of desc (its keyframes_descriptors, `words` UInt32 each).  For target t (map point
 * target_set[t]) the candidates cand[cand_off[t] .. cand_off[t+1]) -- the surrounding keypoints' map points that passed the
 * geometric gates of mapper.jl:405-441, in the caller's order -- are scanned like the reference's loop: distance =
 * mappoint_min_distance (map_point.jl:165-174, minimum Hamming distance over all descriptor pairs, counted in bits =
 * hamming_distance * n_bits), candidates without descriptors skipped, "distance <= best" replaces the best (a later tie
 * wins).  max_distance is the starting value of best / second (256 * max_descriptor_distance, mapper.jl:402).
 * best_pos[t] = position inside the target's candidate list or -1. */
int slamklt_find_best_match(slamklt_ctx* ctx, const uint32_t* desc, int n_desc, int words, const int32_t* set_off, int n_sets,
                            const int32_t* target_set, const int32_t* cand_off, const int32_t* cand, int n_targets,
                            int max_distance, int32_t* best_pos, int32_t* best_dist, int32_t* second_dist);

/* ---- Extractor ------------------------------------------------------------------------- */
/* detect(e, image, current_points; sigma_mask) -- extractor.jl:63-95.  out_yx: cap pairs of Int64.
 * *n_out receives the number of detected keypoints (no global cap, like the reference). */
int slamklt_detect(slamklt_ctx* ctx, const void* img, int dtype, int H, int W, int ld, const double* cur_pts_yx,
                   int n_cur, const slamklt_detect_params* p, int64_t* out_yx, int cap, int* n_out);

/* ---- batched stream API (configs 2, 4, 5: many frames in flight, device-resident) --------------
 * A batch owns n_frames+1 pyramid slots: slot 0 is the previous frame carried over from the last batch,
 * slots 1..n_frames are the frames of this batch.  Pair i tracks points[i] from slot i to slot i+1
 * (preprocess! + klt_tracking! of front_end.jl:454-481 for n_frames consecutive frames).  After
 * slamklt_batch_rotate the last slot becomes slot 0 of the next batch (the copy!(previous, current)). */
int slamklt_batch_create(slamklt_ctx* ctx, int H, int W, int levels, int n_frames, int max_points_per_frame,
                         slamklt_batch** out);
int slamklt_batch_destroy(slamklt_ctx* ctx, slamklt_batch* b);
/* build slot 0 from one host image (first frame of a stream) */
int slamklt_batch_prime(slamklt_ctx* ctx, slamklt_batch* b, const void* img, int dtype, int ld, double sigma, int mode);
/* async H2D of n_frames images (frame f at img + f*frame_stride_bytes) and of n_frames x n_pts x 2 points */
int slamklt_batch_upload(slamklt_ctx* ctx, slamklt_batch* b, const void* imgs, int dtype, int ld,
                         size_t frame_stride_bytes, const double* pts_yx, int n_pts);
/* device-side work only: build all n_frames pyramids, then forward-backward track every pair */
int slamklt_batch_build(slamklt_ctx* ctx, slamklt_batch* b, double sigma, int mode);
int slamklt_batch_track(slamklt_ctx* ctx, slamklt_batch* b, const slamklt_lk_params* p);
/* stereo matching of two batches of equal shape -- optical_flow_matching!(..., kf.left_pyramid, right_pyramid, true) of
 * mapper.jl:51-60 for n_frames keyframes at once: pair i tracks the points uploaded to `to` (they live on `from`'s frame i)
 * from frame i of `from` (left) to frame i of `to` (right); results are fetched with slamklt_batch_download(to, ...) */
int slamklt_batch_track_cross(slamklt_ctx* ctx, slamklt_batch* from, slamklt_batch* to, const slamklt_lk_params* p);
/* build + track of the uploaded frames in one asynchronous call (device-resident step) */
int slamklt_batch_process(slamklt_ctx* ctx, slamklt_batch* b, double sigma, int mode, const slamklt_lk_params* p);
/* async D2H of n_frames x n_pts x 2 tracked points and n_frames x n_pts status bytes, then stream sync */
int slamklt_batch_download(slamklt_ctx* ctx, slamklt_batch* b, double* out_pts_yx, uint8_t* status);
int slamklt_batch_rotate(slamklt_ctx* ctx, slamklt_batch* b);
/* whole step through host buffers: upload + build + track + download + rotate */
int slamklt_batch_step(slamklt_ctx* ctx, slamklt_batch* b, const void* imgs, int dtype, int ld,
                       size_t frame_stride_bytes, const double* pts_yx, int n_pts, double sigma, int mode,
                       const slamklt_lk_params* p, double* out_pts_yx, uint8_t* status);
/* the same step in two halves (stream consumers: front_end.jl:454-481 over many independent sequences).  _begin queues the
 * uploads, builds, tracking and result copies and returns without waiting for the device; _end waits until out_pts_yx / status
 * are filled, then rotates.  imgs, pts_yx, out_pts_yx and status must stay valid and untouched in between.  Several batches of
 * one context may have a step in flight at once: the uploads of one overlap the kernels of the other.  A batch accepts no other
 * call between its _begin and its _end. */
int slamklt_batch_step_begin(slamklt_ctx* ctx, slamklt_batch* b, const void* imgs, int dtype, int ld,
                             size_t frame_stride_bytes, const double* pts_yx, int n_pts, double sigma, int mode,
                             const slamklt_lk_params* p, double* out_pts_yx, uint8_t* status);
int slamklt_batch_step_end(slamklt_ctx* ctx, slamklt_batch* b);
/* diagnostics: measured rates (bytes of Float64 source per second) of the two upload engines slamklt_batch_step uses for
 * page-locked Float64 frames -- the host worker pool that repacks 8-bit data and the plain copies; 0 = not measured yet */
int slamklt_upload_rates(slamklt_ctx* ctx, double* pack_bytes_per_s, double* raw_bytes_per_s);
/* parity access: view of slot `slot` (0..n_frames) as a pyramid handle owned by the batch */
int slamklt_batch_slot(slamklt_batch* b, int slot, slamklt_pyr** out);
/* batched detect on the frames last uploaded to the batch (config 5: re-extraction every frame).
 * cur_pts: n_frames x n_cur x 2 (may be NULL when n_cur == 0); out_yx: n_frames x cap pairs; n_out: n_frames. */
int slamklt_batch_detect(slamklt_ctx* ctx, slamklt_batch* b, const double* cur_pts_yx, int n_cur,
                         const slamklt_detect_params* p, int64_t* out_yx, int cap, int* n_out);

/* pinned host memory helpers (so that async copies really are asynchronous) */
int slamklt_host_alloc(size_t bytes, void** out);
int slamklt_host_free(void* p);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* SLAMKLT_H */
