// BRIEF descriptors and Hamming matching (only active with params.do_local_matching, params.jl:69).
//   describe(e, image, keypoints)          extractor.jl:103-105 -> ImageFeatures.create_descriptor(img, keypoints, BRIEF(size = 256)) [3P]
//   mappoint_min_distance(m1, m2)          map_point.jl:165-174 (minimum Hamming distance over two sets of descriptors)
//   find_best_match(...) selection loop    mapper.jl:392-462 (best candidate with the reference's "<=": a later tie wins)
// [3P ImageFeatures 0.4, brief.jl]: the image is smoothed with imfilter(img, Kernel.gaussian(sigma)) (sigma = sqrt 2: 9 x 9 FIR,
// replicate border), keypoints closer than ceil(window / 2) to the border are dropped, bit b of a descriptor is
// smoothed[k + s1[b]] < smoothed[k + s2[b]].  The sampling pairs come from Julia's RNG (Random.seed!(123); gaussian(size,
// window)) and cannot be regenerated outside Julia: the caller passes them (the shim exports them once).
#include "common.cuh"

namespace sk {
namespace {

constexpr int BR_R = 5;             // largest sampling offset supported: ceil(window / 2) with window <= 9 (clamped to the window)
constexpr int BR_T = 8;             // largest half width of the smoothing kernel (sigma <= 4)
constexpr int BR_S = 2 * BR_R + 1;  // smoothed patch side
constexpr int BR_P = BR_S + 2 * BR_T;

// One warp per keypoint: (BR_S + 2 hw)^2 raw patch (replicate-clamped) -> y pass -> x pass -> BR_S x BR_S smoothed values in
// shared memory -> every lane evaluates bits lane, lane + 32, ... and the warp packs them with ballots.
__global__ void __launch_bounds__(128) k_brief(const BriefArgs a) {
    __shared__ double s_raw[4][BR_P][BR_P + 1];
    __shared__ double s_tmp[4][BR_S][BR_P + 1];
    __shared__ double s_sm[4][BR_S][BR_S + 1];
    const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = blockIdx.x * 4 + wl;
    if (k >= a.n) return;
    const int hw = a.hw, side = BR_S + 2 * hw;
    const long long ky = a.kps[2 * (size_t)k], kx = a.kps[2 * (size_t)k + 1];  // 1-based
    const bool valid = ky - a.lim >= 1 && kx - a.lim >= 1 && ky + a.lim <= a.H && kx + a.lim <= a.W;
    if (lane == 0) a.valid[k] = valid ? 1 : 0;
    if (!valid) {
        for (int w = lane; w < a.n_bits / 32; w += 32) a.desc[(size_t)k * (a.n_bits / 32) + w] = 0u;
        return;
    }
    const double* img = a.img;
    for (int i = lane; i < side * side; i += 32) {
        const int py = i % side, px = i / side;
        const int y = min(max((int)ky - 1 - BR_R - hw + py, 0), a.H - 1), x = min(max((int)kx - 1 - BR_R - hw + px, 0), a.W - 1);
        s_raw[wl][px][py] = img[(size_t)y + (size_t)x * a.ld];
    }
    __syncwarp();
    // dim 1 (y) first, then dim 2 (x), taps accumulated in order (imfilter with the factored Kernel.gaussian)
    for (int i = lane; i < BR_S * side; i += 32) {
        const int y = i % BR_S, px = i / BR_S;
        double acc = 0.0;
        for (int t = 0; t <= 2 * hw; ++t) acc += a.kw[t] * s_raw[wl][px][y + t];
        s_tmp[wl][y][px] = acc;
    }
    __syncwarp();
    for (int i = lane; i < BR_S * BR_S; i += 32) {
        const int y = i % BR_S, x = i / BR_S;
        double acc = 0.0;
        for (int t = 0; t <= 2 * hw; ++t) acc += a.kw[t] * s_tmp[wl][y][x + t];
        s_sm[wl][y][x] = acc;
    }
    __syncwarp();
    for (int w = 0; w < a.n_bits / 32; ++w) {
        const int b = 32 * w + lane;
        const int4 pr = a.pairs[b];  // (dy1, dx1, dy2, dx2)
        const bool bit = s_sm[wl][BR_R + pr.x][BR_R + pr.y] < s_sm[wl][BR_R + pr.z][BR_R + pr.w];
        const unsigned word = __ballot_sync(FULL, bit);
        if (lane == 0) a.desc[(size_t)k * (a.n_bits / 32) + w] = word;
    }
}

__device__ __forceinline__ int set_min_distance(const HammingArgs& a, int s1, int s2) {
    // mappoint_min_distance: minimum over all pairs of the two sets; 1e6 (here: INT_MAX) for an empty set
    int best = 0x7fffffff;
    for (int i = a.set_off[s1]; i < a.set_off[s1 + 1]; ++i)
        for (int j = a.set_off[s2]; j < a.set_off[s2 + 1]; ++j) {
            int d = 0;
            for (int w = 0; w < a.words; ++w) d += __popc(a.desc[(size_t)i * a.words + w] ^ a.desc[(size_t)j * a.words + w]);
            best = min(best, d);
        }
    return best;
}

// One thread per target map point: the candidate loop of find_best_match in the caller's order.
__global__ void k_best_match(const HammingArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n_targets) return;
    int best = a.max_distance, second = a.max_distance, best_pos = -1;
    for (int c = a.cand_off[t]; c < a.cand_off[t + 1]; ++c) {
        const int s2 = a.cand[c];
        if (a.set_off[s2 + 1] == a.set_off[s2]) continue;  // isempty(mp.descriptor)
        const int d = set_min_distance(a, a.target_set[t], s2);
        if (d <= best) { second = best; best = d; best_pos = c - a.cand_off[t]; }
        else if (d <= second) second = d;
    }
    a.best_pos[t] = best_pos;
    a.best_dist[t] = best;
    a.second_dist[t] = second;
}

}  // namespace

void launch_brief(cudaStream_t s, const BriefArgs& a) { k_brief<<<(a.n + 3) / 4, 128, 0, s>>>(a); }
void launch_best_match(cudaStream_t s, const HammingArgs& a) { k_best_match<<<(a.n_targets + 127) / 128, 128, 0, s>>>(a); }

}  // namespace sk
