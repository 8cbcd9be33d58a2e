// Internal declarations shared by the translation units of libslamklt.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/slamklt.h"

namespace sk {

constexpr int MAX_LAYERS = 8;   // pyramid_levels + 1 <= 8
constexpr unsigned FULL = 0xffffffffu;

// device planes kept per level (fp32, y contiguous, pitch = roundup4(H+1), W+1 columns, guards kept at zero)
enum DevPlane {
    DP_I = 0,              // layer
    DP_GRAD = 1,           // interleaved (Iy, Ix): float2 per pixel, occupies two plane slots
    DP_RYY = 3, DP_RXX, DP_RYX,  // exclusive prefix sums along x of the smoothed Iy*Iy, Ix*Ix, Iy*Ix (W+1 columns)
    DP_T0, DP_T1, DP_T2,   // scratch: planes after the dim-1 recursive pass
    DP_BLUR,               // LKCache.gaussian_filtered
    DP_TMP,                // scratch for parity downloads
    DP_COUNT
};

struct LevelGeom {
    int H, W, pitch;
    int pad_;
    size_t plane_elems;  // floats per plane (multiple of 32)
    size_t off;          // float offset of this level's block inside a frame
};

struct PyrGeom {
    int nl;  // layers = levels + 1
    int H0, W0;
    int pad_;
    LevelGeom lv[MAX_LAYERS];
    size_t frame_elems;  // floats per frame (all levels, all planes)
};

// A set of frames laid out with a constant stride; logical frame f lives in physical slot (slot0 + f) % n_slots.
struct FrameSet {
    float* base;
    size_t frame_elems;
    int n_slots;
    int slot0;
    __host__ __device__ int slot(int f) const { return (slot0 + f) % n_slots; }  // physical slot of logical frame f
    __host__ __device__ float* frame(int f) const { return base + (size_t)slot(f) * frame_elems; }
};

__host__ __device__ inline size_t plane_off(const LevelGeom& g, int plane) { return g.off + (size_t)plane * g.plane_elems; }

// recursive Gaussian constants (host computes in double, device uses float)
struct IirDev {
    float a1, a2, a3, scale, inv1ma;  // inv1ma = 1/(1 - (a1+a2+a3))
    float M[9];                       // Triggs-Sdika right-boundary matrix, row-major
    float P[5][9];                    // A^(K*2^j), j = 0..4, for the warp scan of the dim-1 kernel
    float PR[5][9];                   // A^(KR*2^j), j = 0..4, for the chunk-carry scan of the dim-2 kernel
    int nsc, nsr;                     // scan steps actually needed: later powers of A are below 1e-13 and are skipped
};

struct LKLevel {
    int H, W, pitch, pad_;
    size_t oI, oG, oRyy, oRxx, oRyx;
};

struct LKArgs {
    FrameSet A, B;
    int offA, offB;
    int nl, mode;  // mode 0 = optflow!, 1 = fb_tracking!, 2 = optical_flow_matching! (prior pass, then full pass)
    LKLevel lv[MAX_LAYERS];
    const double* pts;
    const double* disp_in;  // nullable
    double* disp_out;       // nullable
    double* out_pts;        // nullable (mode 1)
    uint8_t* status;
    int n_per_frame, n_frames;
    int iterations, window, levels, pad_;
    double eig_thr, eps, max_dist;
    unsigned long long* counters;  // [0] window px * iterations, [1] iterations
    const uint8_t* has_prior;      // mode 2: per point, 1 = 3-D keypoint tracked first with disp_in and levels3d, 2 = skip (status 8)
    int levels3d, pad2_;
    unsigned* work;                // nullable: device counter the persistent grid draws keypoint indices from
    int chunk, pad3_;              // keypoints drawn per request
    // TMA-staged kernel (lk_tma.cu): device arrays of per-level tensor maps (LKTmaLevel[MAX_LAYERS]) over the frame rings A and B;
    // nullptr selects the cp.async kernel of lk_patch.cu
    const void* mapsA;
    const void* mapsB;
    // optional table of first-set-up structure tensors, [keypoint][gtab_levels] entries of 16 bytes, filled by k_lk_gprep
    void* gtab;
    int gtab_levels, pad4_;
    unsigned long long npf_magic;  // ceil(2^40 / n_per_frame), set by the launcher (frame index of a keypoint without a division)
};

struct DetArgs {
    const double* img;  // n_frames images, column-major H x W, frame stride H*W
    int H, W;
    const double* cur;  // n_frames x n_cur x 2, nullable
    int n_cur, n_frames;
    int radius, grid_h, grid_w, cs, k_cell, hw, slots, pad_;
    double min_resp;
    double kw[33];      // 1-D mask blur weights, length 2*hw+1
    const void* src;    // register-tiled kernel: the staged frames in their own type (src_dtype: SLAMKLT_F64 / _F32 / _U8), converted on load
    int src_dtype, pad2_;
    signed char sy[64]; // register-tiled kernel, masked: half height of the disc at column offset dx = i - radius (-1 = empty), host-built
    int sy_valid, pad3_; // 0: the disc is wider than the table (the kernel computes the half heights itself)
    int2* bin_pts;      // register-tiled kernel, masked: per (frame, cell row) the current points within reach (k_detect_bin), or nullptr
    int* bin_cnt;
    const double* ytab; // hw == 6 only: the 2^13 sums of tap subsets (bit t set = tap t included), added in tap order (device)
    int64_t* cell_out;  // [frame][cell][slots][2]
    int* cell_cnt;      // [frame][cell]
    int64_t* out;       // [frame][cap][2]
    int* n_out;         // [frame]
    int cap;
};

// optical_flow_matching! prologue / epilogue (match.cu)
struct MatchCam { double fx, fy, cx, cy, k1, k2, p1, p2; };
struct MatchArgs {
    int n, stereo;
    MatchCam cam, rcam;
    double T[16];            // world -> camera (column-major): cw, or Ti0 * cw in stereo mode
    double bound_h, bound_w; // image bounds of the camera the projection must fall into
    double scale, epipolar;
    const double* pix;       // n x 2 kp.pixel
    const uint8_t* is_3d;    // n
    const double* world;     // n x 3 map-point positions
    const double* undist;    // n x 2 kp.undistorted_pixel (stereo)
    uint8_t* flag;           // out of the prologue: 0 no prior, 1 prior, 2 projection outside the image
    double* disp;            // out of the prologue: prior displacement
    const double* tracked;   // n x 2 result of the tracking kernel
    uint8_t* status;         // in/out
    double* out_pix; double* out_und; double* out_pos;
};
// BRIEF describe + Hamming matching (brief.cu)
struct BriefArgs {
    const double* img;      // column-major H x W (ld)
    int H, W, ld, n;
    const long long* kps;   // n x 2 (y, x), 1-based
    const int4* pairs;      // n_bits x (dy1, dx1, dy2, dx2)
    int n_bits, lim, hw, pad_;
    double kw[2 * 8 + 1];   // 1-D smoothing taps, length 2*hw+1
    unsigned* desc;         // n x n_bits/32
    uint8_t* valid;         // n
};
struct HammingArgs {
    const unsigned* desc;   // descriptor rows, `words` 32-bit words each
    const int* set_off;     // n_sets + 1: rows [set_off[s], set_off[s+1]) are the descriptors of map point s
    int words, n_targets, max_distance, pad_;
    const int* target_set;  // n_targets
    const int* cand_off;    // n_targets + 1
    const int* cand;        // candidate set ids, in the caller's order
    int* best_pos;          // position of the best candidate inside the target's list, -1 if none
    int* best_dist;
    int* second_dist;
};
void launch_brief(cudaStream_t s, const BriefArgs& a);
void launch_best_match(cudaStream_t s, const HammingArgs& a);

// triangulate_stereo! (mapper.jl:142-183)
struct TriArgs {
    int n, pad_;
    MatchCam cam, rcam;
    double Ti0[16];          // right camera: transformation from the left camera (column-major)
    double wc[16];           // frame.wc (camera -> world, column-major)
    double max_error;
    const double* und;       // n x 2 kp.undistorted_pixel (y, x)
    const double* rund;      // n x 2 kp.right_undistorted_pixel (y, x)
    double* world;           // out: n x 3 world point (NaN unless status 1)
    uint8_t* status;         // out: 1 triangulated, 2 / 3 depth < 0.1 in the left / right camera, 4 / 5 reprojection error left / right
};
void launch_triangulate_stereo(cudaStream_t s, const TriArgs& a);
void launch_match_prior(cudaStream_t s, const MatchArgs& a);
void launch_match_update(cudaStream_t s, const MatchArgs& a);

// Scratch ring for the y-filtered product planes T0..T2 (the only consumer is the x pass that follows): when a batch is built in
// groups of a few frames, the planes of frame f live in slot f % ring of a small separate buffer instead of inside the frame
// block, so every group rewrites the same addresses and -- with an L2 access-policy window over the buffer -- the planes are
// produced and consumed in L2 without ever reaching HBM.  base == nullptr: the planes inside the frame block are used.
struct TScratch {
    float* base;
    size_t stride;            // floats per ring slot (all levels)
    int ring;                 // slots
    int pad_;
    size_t off[MAX_LAYERS];   // float offset of level l's three planes inside a slot
};

// optional per-kernel profiling hook: called with the kernel's name right before each launch
struct Hook { void (*fn)(void* user, const char* name); void* user; };
inline void mark(const Hook* h, const char* name) { if (h && h->fn) h->fn(h->user, name); }

// launchers (each returns the number of kernels it launched; errors are picked up by cudaGetLastError in api.cu)
int launch_convert(cudaStream_t s, const void* src, int dtype, int ld, size_t src_frame_stride_elems, FrameSet dst, int dst_f0,
                   int n_frames, const PyrGeom& g, double* dst64 /*nullable: also keep f64 copy, frame stride H*W*/, const Hook* hk);
// streams/events for the pyramid build DAG: the blur chain (main) runs concurrently with the gradient stages (b: level 0,
// c: levels >= 1), so the small coarse-level kernels overlap the big level-0 ones
struct PyrStreams { cudaStream_t main, b, c; cudaEvent_t ev[MAX_LAYERS + 3]; bool parallel; };
int launch_pyramid(const PyrStreams& ps, FrameSet fs, int f0, int n_frames, const PyrGeom& g, double sigma, int mode,
                   const float* const* inv_ny, const float* const* inv_nx /* per level device arrays, CTOR mode only */,
                   const void* raw /* staged host image(s) or nullptr */, int dtype, const Hook* hk,
                   const TScratch* ts = nullptr /* T planes in a scratch ring */, bool join = true /* side streams rejoin main */);
int launch_smoothed_plane(cudaStream_t s, FrameSet fs, int f0, const PyrGeom& g, int level, int which, const Hook* hk);
int launch_lk(cudaStream_t s, const LKArgs& a, const Hook* hk);
bool launch_lk_patch(cudaStream_t s, const LKArgs& a);  // patch-mapped variant (lk_patch.cu), windows up to 23 x 23
bool launch_lk_tma(cudaStream_t s, const LKArgs& a);    // TMA-staged patch variant (lk_tma.cu), windows up to 19 x 19
// Tensor maps of one frame ring for the TMA-staged tracking kernel: fills MAX_LAYERS * lk_tma_level_bytes() bytes at host_out
// (to be copied to 64-byte aligned device memory).  Returns 0, or -1 with a message in err.
size_t lk_tma_level_bytes();
int lk_tma_encode(const PyrGeom& g, float* base, int n_slots, void* host_out, char* err, size_t errcap);
int launch_detect(cudaStream_t s, const DetArgs& a, const Hook* hk);
size_t detect_smem_bytes(int cs, int hw);
size_t detect2_smem_bytes(int cs, int hw);
bool detect2_supported(const DetArgs& a);  // the register-tiled kernel covers this request (it reads F32 / U8 frames directly)

void iir_design(double sigma, double a[3], double* scale, double M[9]);
void iir_dev(double sigma, int K, int KR, IirDev* out);
void iir_line_host(double* x, int n, double sigma, double iminus, double iplus);

int pick_K(int H);  // per-lane chunk of the dim-1 kernel; 0 if unsupported
constexpr int KR = 40;  // per-thread chunk of the dim-2 kernel

}  // namespace sk
