// Pyramidal Lucas-Kanade with forward-backward check, TMA-staged patch variant (sm_100a).
//
// Same semantics as lk_patch.cu / lk.cu (reference: lucas_kanade.jl:9-212, utils.jl:5-45, tracker.jl:17-68, and the prior pass +
// retry of map_manager.jl:517-551 in mode 2) and the same lane mapping (8 row groups x 4 column groups, a lane owns a PR x PC
// patch of the window).  What changes is how data reaches the warp:
//   * every tile comes through the tensor-memory accelerator: per pyramid level three tensor maps describe the frame ring as a
//     3-D tensor (y, x, slot) -- the layer plane with the target-tile box, the layer plane with the template box and the
//     interleaved (Iy, Ix) plane with the template box.  One lane issues ONE cp.async.bulk.tensor.3d per tile (UTMALDG);
//     coordinates are plain element indices (no alignment rule, out-of-range elements arrive as zeros), completion is an
//     mbarrier transaction count.  This replaces ~9 predicated cp.async per lane for the target tile and 30 predicated __ldg per
//     lane for the template, with their address arithmetic;
//   * the template patch is read from shared memory with immediate offsets; rows / columns of the patch grid beyond the window
//     are neutralised by three per-lane row weights applied to the per-row accumulators and by a never-written zero column of
//     the gradient tile (no per-element predicates on the common path);
//   * the template of the NEXT forward level depends only on the keypoint, so it is requested as soon as the current level's
//     template sits in registers and arrives while the current level iterates;
//   * the 2x2 structure tensor of a level's first set-up depends only on (keypoint, level): with a.gtab it comes from a table
//     filled by k_lk_gprep (one 8-lane group per entry) instead of 32 lanes redundantly running the Float64 gate and inverse.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace sk {

// ---- geometry --------------------------------------------------------------------------------------------------------
// TRt: pitch (= box rows) of the target tile; PC * TRt must be 8 or 24 (mod 32) so that the 8 x 4 lane grid (row stride PR = 3
// words, column-group stride PC * TRt words) hits 32 distinct banks: 24 or 40 for PC = 5.
#ifndef LKT_TR
#define LKT_TR 40
#endif
#ifndef LKT_MXL
#define LKT_MXL 4
#endif
#ifndef LKT_MXR
#define LKT_MXR 4
#endif
#ifndef LKT_SLIM
#define LKT_SLIM 0
#endif
// latency experiments: G-table entry requested one stage ahead (LKT_PF_G) / next keypoint's coordinates requested during the
// last stage of the current one (LKT_PF_PT).  Measured on B200 (forward-backward kernel, 18 CTAs per SM): neither 0.810 ms, table
// entry ahead 0.811 ms, coordinates ahead 0.815 ms, both 0.822 ms -- with 18 warps per SM the kernel is bound by instruction
// issue, not by these round trips, and the extra live registers cost more than the hidden latency returns.  Both off.
#ifndef LKT_PF_G
#define LKT_PF_G 0
#endif
#ifndef LKT_PF_PT
#define LKT_PF_PT 0
#endif

template <int W2, int PR, int PC>
struct TmaTile {
    static constexpr int RGU = (W2 + PR - 1) / PR;            // row groups that own window rows (7 of 8 for 19 rows)
    static constexpr int RSPAN = (RGU - 1) * PR + PR + 1;     // tap rows touched (idle row groups alias the last used one)
    static constexpr int CSPAN = PC * 4 + 1;                  // tap columns touched
    static constexpr int TR = LKT_TR, MXL = LKT_MXL, MXR = LKT_MXR;
    static constexpr int MY = (TR - RSPAN) / 2;               // margin above the first tap row when a tile is staged
    static constexpr int TC = CSPAN + MXL + MXR;              // margins left / right of the tap columns
    static constexpr int AR = ((RGU * PR + 3 + 3) / 4) * 4;   // template tile rows: RGU*PR read + up to 3 skipped for the 16-byte aligned box start
    static constexpr int GC = W2;                             // gradient box columns
    // LKT_SLIM: the template tiles hold exactly the W2 window columns and the gradient tile W2 + 1 rows (its box start only has
    // to be even); the patch grid's 20th column / rows past the tile are never used (replaced by zeros with a select), their
    // loads are clamped or fall on neighbouring tile data.  Otherwise: 20 columns x AR rows with a zero 20th column.
    static constexpr int AC = LKT_SLIM ? W2 : PC * 4;         // layer (template) tile columns
    static constexpr int ARG = LKT_SLIM ? W2 + 1 : AR;        // gradient tile rows (float2): its 8-byte rows only need an even box start
    static constexpr int GSKIP = LKT_SLIM ? 1 : 3;            // mask of the rows skipped in the gradient tile (box start = (r0 - 1) & ~GSKIP)
    static constexpr int GCS = LKT_SLIM ? GC : PC * 4;        // gradient tile columns in shared memory
    static constexpr unsigned T_BYTES = TC * TR * 4, I_BYTES = AC * AR * 4, G_BYTES = GC * ARG * 8;
    static_assert((ARG * 8) % 16 == 0, "gradient box rows");
    static_assert(TR >= RSPAN && TR % 4 == 0 && ((PC * TR) % 32 == 8 || (PC * TR) % 32 == 24) && PR == 3, "tile geometry");
    static_assert(PR * 8 >= W2 && PC * 4 >= W2, "patch grid must cover the window");
};

// per-level tensor maps of one frame ring
struct LKTmaLevel {
    CUtensorMap tgt;  // layer plane, box TR x TC x 1
    CUtensorMap ti;   // layer plane, box AR x AC x 1
    CUtensorMap tg;   // (Iy, Ix) plane viewed as fp32 with 2*pitch rows, box 2*AR x GC x 1
};

size_t lk_tma_level_bytes() { return sizeof(LKTmaLevel); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

int lk_tma_encode(const PyrGeom& g, float* base, int n_slots, void* host_out, char* err, size_t errcap) {
    using T = TmaTile<19, 3, 5>;
    EncodeTiledFn enc = encode_fn();
    if (!enc) { snprintf(err, errcap, "cuTensorMapEncodeTiled is not available from this driver"); return -1; }
    LKTmaLevel* out = (LKTmaLevel*)host_out;
    std::memset(out, 0, sizeof(LKTmaLevel) * MAX_LAYERS);
    for (int l = 0; l < g.nl; ++l) {
        const LevelGeom& L = g.lv[l];
        const cuuint32_t es[3] = {1, 1, 1};
        auto one = [&](CUtensorMap* m, float* gaddr, cuuint64_t rows, cuuint32_t brows, cuuint32_t bcols) -> int {
            const cuuint64_t dims[3] = {rows, (cuuint64_t)L.W + 1, (cuuint64_t)n_slots};
            const cuuint64_t strides[2] = {rows * 4, (cuuint64_t)g.frame_elems * 4};
            const cuuint32_t box[3] = {brows, bcols, 1};
            CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, gaddr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { snprintf(err, errcap, "cuTensorMapEncodeTiled failed (%d) at level %d", (int)r, l); return -1; }
            return 0;
        };
        if (one(&out[l].tgt, base + plane_off(L, DP_I), (cuuint64_t)L.pitch, T::TR, T::TC)) return -1;
        if (one(&out[l].ti, base + plane_off(L, DP_I), (cuuint64_t)L.pitch, T::AR, T::AC)) return -1;
        if (one(&out[l].tg, base + plane_off(L, DP_GRAD), (cuuint64_t)L.pitch * 2, 2 * T::ARG, T::GC)) return -1;
    }
    return 0;
}

// ---- device helpers ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned s_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init1(uint64_t* bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_par(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "LKT_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra LKT_WAIT_%=;\n\t}" ::"r"(s_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_box(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(s_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(s_u32(bar))
                 : "memory");
}
// generic-proxy reads of a tile must be ordered before the async proxy overwrites it
#ifndef LKT_FENCE
#define LKT_FENCE 1
#endif
__device__ __forceinline__ void fence_async_smem() {
#if LKT_FENCE
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}

__device__ __forceinline__ void wsum2(float a, float b, int lane, float& sa, float& sb) {
    const bool hi = lane & 16;
    float v = (hi ? b : a) + __shfl_xor_sync(FULL, hi ? a : b, 16);  // lower half: a-partials, upper half: b-partials
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    sa = __shfl_sync(FULL, v, 0);
    sb = __shfl_sync(FULL, v, 16);
}
__device__ __forceinline__ void wsum3(float a, float b, float c, int lane, float& sa, float& sb, float& sc) {
    const bool hi = lane & 16, q = lane & 8;
    float v = (hi ? b : a) + __shfl_xor_sync(FULL, hi ? a : b, 16);
    c += __shfl_xor_sync(FULL, c, 16);
    v = (q ? c : v) + __shfl_xor_sync(FULL, q ? v : c, 8);  // lanes with bit 3 set now carry c-partials
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    sa = __shfl_sync(FULL, v, 0);
    sb = __shfl_sync(FULL, v, 16);
    sc = __shfl_sync(FULL, v, 8);
}

// Float64 part of a set-up: eigenvalue gate (lucas_kanade.jl:38-46) and G^-1 (utils.jl:5-45).  Singular values of the symmetric
// G = [a b; b c] are Q +- R with Q = |a+c|/2, R = sqrt(((a-c)/2)^2 + b^2), so min(S)/area < thr  <=>  |Q - R| < t, t = thr*area
// <=>  R < Q + t and R > Q - t: decided on squares, no sqrt / division on the common path.  Returns false when the gate fails.
__device__ __forceinline__ bool g_inverse(double ga, double gc, double gb, double t, double& g00, double& g01, double& g11) {
    const double E = 0.5 * (ga + gc), F = 0.5 * (ga - gc);
    const double R2 = F * F + gb * gb, Q = fabs(E);
    const double qp = Q + t, qm = Q - t;
    if (t > 0.0 && R2 < qp * qp && (qm < 0.0 || R2 > qm * qm)) return false;
    const double tol = 1.4901161193847656e-08;  // sqrt(eps(Float64)), utils.jl:37
    bool full_rank = t > tol;                     // the gate already guarantees min(S) >= t
    double R = 0.0;
    if (!full_rank) { R = sqrt(R2); full_rank = fabs(Q - R) > tol; }
    if (full_rank) {
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
    return true;
}

// ---- G table pre-pass -----------------------------------------------------------------------------------------------------
// One entry per (keypoint, level): the structure tensor of the window get_offsets(p, p) (lucas_kanade.jl:34-46) depends only on the
// keypoint and the level, so it is computed once by an 8-lane group (lane = window rows r, r+8, r+16) instead of by all 32 lanes
// of the tracking warp at every level.  Entry: g00, g01, g11 of G^-1 (fp32) and a flag (1 = gate passed, 0 = failed / window empty).
typedef float4 GEntry;  // g00, g01, g11 of G^-1 rounded to fp32, w = 1 when the gate passed

__global__ void __launch_bounds__(128) k_lk_gprep(const LKArgs a, GEntry* __restrict__ tab) {
    const int sub = threadIdx.x & 7;
    const long long item = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int total = a.n_frames * a.n_per_frame;
    const int nlv = a.gtab_levels;
    if (item >= (long long)total * nlv) return;
    const int gw = (int)(item / nlv), lvl = (int)(item - (long long)gw * nlv);
    const int f = gw / a.n_per_frame;
    const float* fbA = a.A.frame(a.offA + f);
    const double pty = a.pts[2 * (size_t)gw], ptx = a.pts[2 * (size_t)gw + 1];
    const LKLevel& L = a.lv[lvl];
    const int H = L.H, W = L.W, pitch = L.pitch, w = a.window;
    const double inv = __longlong_as_double((long long)(1023 - lvl) << 52);
    const int py = (int)floor(pty * inv), px = (int)floor(ptx * inv);
    const int up = min(w, py - 1), down = min(w, H - py), left = min(w, px - 1), right = min(w, W - px);
    const int nrows = up + down + 1, ncols = left + right + 1;
    const int r0 = py - up, c0 = px - left;
    const bool valid = !(nrows < 1 || ncols < 1 || r0 < 1 || c0 < 1 || py + down > H || px + right > W);
    float syy = 0.f, sxx = 0.f, syx = 0.f;
    if (valid) {
        const size_t lo = (size_t)(c0 - 1) * pitch, hi = (size_t)(px + right) * pitch;
        for (int r = sub; r < nrows; r += 8) {
            const float* colA = fbA + (size_t)(r0 - 1 + r);
            syy += __ldg(colA + L.oRyy + hi) - __ldg(colA + L.oRyy + lo);
            sxx += __ldg(colA + L.oRxx + hi) - __ldg(colA + L.oRxx + lo);
            syx += __ldg(colA + L.oRyx + hi) - __ldg(colA + L.oRyx + lo);
        }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        syy += __shfl_xor_sync(FULL, syy, o);
        sxx += __shfl_xor_sync(FULL, sxx, o);
        syx += __shfl_xor_sync(FULL, syx, o);
    }
    if (sub == 0) {
        double g00 = 0.0, g01 = 0.0, g11 = 0.0;
        const bool pass = valid && g_inverse((double)syy, (double)sxx, (double)syx, a.eig_thr * (double)(nrows * ncols), g00, g01, g11);
        tab[item] = make_float4((float)g00, (float)g01, (float)g11, pass ? 1.f : 0.f);
    }
}

// ---- tracking kernel --------------------------------------------------------------------------------------------------------
// 10 496 bytes (LKT_SLIM: 9 536 with a 26-column target tile): with the 1 KB the hardware reserves per CTA, twenty (twenty-two)
// one-warp CTAs fit the 227 KB of an SM.  The small members sit in the padding between the 128-byte aligned tiles.
#if LKT_SLIM
template <int W2, int PR, int PC>
struct __align__(128) LKTmaSmem {
    using T = TmaTile<W2, PR, PC>;
    float sI[T::AC][T::AR];
    __align__(16) double q[2];               // forward result (the backward pass starts from it; the distance gate needs it again)
    int dsave[4];                            // displacement at the start of the current level (optflow! keeps it when the level fails)
    __align__(8) uint64_t barT;
    uint64_t barA;
    __align__(128) float2 sG[T::GCS][T::ARG];
    __align__(128) float sT[T::TC][T::TR];
};
#else
template <int W2, int PR, int PC>
struct __align__(128) LKTmaSmem {
    using T = TmaTile<W2, PR, PC>;
    float sT[T::TC][T::TR];
    __align__(16) double q[2];               // forward result (the backward pass starts from it; the distance gate needs it again)
    int dsave[4];                            // displacement at the start of the current level (optflow! keeps it when the level fails)
    __align__(8) uint64_t barT;
    uint64_t barA;
    __align__(128) float sI[T::AC][T::AR];
    __align__(128) float2 sG[T::GCS][T::ARG];
};
static_assert(LKT_MXL + LKT_MXR != 8 || sizeof(LKTmaSmem<19, 3, 5>) == 10496, "shared memory budget of the tracking kernel (20 CTAs per SM)");
#endif

// Position bookkeeping.  The reference keeps d and c as Float64 vectors; the kernel keeps the current estimate of a level as an
// integer pixel plus an fp32 fraction in [0, 1): floor / ceil / bilinear weights / tile offsets then need no Float64 arithmetic
// inside the iteration, and the fraction's resolution (6e-8 px) is far below the fp32 noise of the window sums that drive it.
// Between levels the displacement is doubled exactly (integer part and fraction separately).
__device__ __forceinline__ void split_d(double d, int& di, float& df) {
    di = __double2int_rd(d);
    df = (float)(d - (double)di);
    if (df >= 1.f) { di += 1; df = 0.f; }  // the rounding of the fraction can reach 1
}
__device__ __forceinline__ double join_d(int di, float df) { return (double)di + (double)df; }

// one keypoint, one warp: all levels of the forward pass, the backward pass and the gates.  MODE (= a.mode) is a template parameter
// so that the forward-backward kernel carries none of optflow!'s or optical_flow_matching!'s bookkeeping
template <int W2, int PR, int PC, int MODE>
__device__ __forceinline__ void lk_point_tma(const LKArgs& a, const int gw, const int lane, LKTmaSmem<W2, PR, PC>& sm, unsigned& parT, unsigned& parA,
                                             const double2 pt0, const int next_raw, int& next, double2& pt_next, unsigned long long& acc_wpx,
                                             unsigned long long& acc_nit) {
    using T = TmaTile<W2, PR, PC>;
    constexpr int TR = T::TR, TC = T::TC, AR = T::AR;
    // frame of this keypoint: gw / n_per_frame as a 64-bit multiply by ceil(2^40 / n_per_frame) (exact for gw < 2^40 / n_per_frame)
    const int f = (int)(((unsigned long long)(unsigned)gw * a.npf_magic) >> 40);
    // template side (A) and target side (B) swap roles on the backward pass: keep the two physical slots and derive pointers
    // and tensor maps where they are used (once per set-up) instead of carrying them in registers.  slot = (slot0 + off + f) mod
    // n_slots with one conditional subtraction (slot0 < n_slots and off + f <= n_slots)
    int slot_first = a.A.slot0 + a.offA + f, slot_second = a.B.slot0 + a.offB + f;
    if (slot_first >= a.A.n_slots) slot_first -= a.A.n_slots;
    if (slot_second >= a.B.n_slots) slot_second -= a.B.n_slots;
    bool swapped = false;
#define LKT_SLOT_A (swapped ? slot_second : slot_first)
#define LKT_SLOT_B (swapped ? slot_first : slot_second)
#define LKT_MAPS_A (reinterpret_cast<const LKTmaLevel*>(swapped ? a.mapsB : a.mapsA))
#define LKT_MAPS_B (reinterpret_cast<const LKTmaLevel*>(swapped ? a.mapsA : a.mapsB))
#define LKT_FB_A (swapped ? a.B.base + (size_t)slot_second * a.B.frame_elems : a.A.base + (size_t)slot_first * a.A.frame_elems)
    // mode 2 = optical_flow_matching! (map_manager.jl:451-564): keypoints with a prior (3-D keypoints) are first tracked with
    // that prior on levels3d levels; those that fail, and all others, are tracked from a zero displacement on all levels
    const uint8_t prior_flag = (MODE == 2 && a.has_prior) ? a.has_prior[gw] : (uint8_t)0;
    if (prior_flag == 2) {  // 3-D keypoint whose projection left the image: not tracked at all (map_manager.jl:489-506)
        if (lane == 0) {
            a.status[gw] = 8;
            if (a.out_pts) {
                a.out_pts[2 * (size_t)gw] = __longlong_as_double(0x7ff8000000000000LL);
                a.out_pts[2 * (size_t)gw + 1] = __longlong_as_double(0x7ff8000000000000LL);
            }
        }
        return;
    }
    const bool prior_first = MODE == 2 && prior_flag != 0;
    int levels_cur = prior_first ? a.levels3d : a.levels;
    bool second_try = false;
    const float4* gtab = reinterpret_cast<const float4*>(a.gtab);
    float4 gnext = make_float4(0.f, 0.f, 0.f, 0.f);  // table entry of the next stage, requested one stage ahead (or at the keypoint's start)
    if (LKT_PF_G && gtab) gnext = __ldg(gtab + ((size_t)gw * a.gtab_levels + levels_cur));

    const int w = a.window;
    unsigned int wpx = 0, nit = 0;
    bool ok = true;
    uint8_t result = 0;
    int diy = 0, dix = 0;      // displacement at the scale of the current level: integer part ...
    float dfy = 0.f, dfx = 0.f;  // ... and fraction in [0, 1)
    int iy0, ix0;              // floor of the point the templates are centred on (keypoint; forward result on the backward pass)
    {
        const double2 pt = pt0;  // requested while the previous keypoint was still being tracked
        iy0 = __double2int_rd(pt.x); ix0 = __double2int_rd(pt.y);
        if (a.disp_in && !(MODE == 2 && !prior_first)) {
            const double2 d0 = *reinterpret_cast<const double2*>(a.disp_in + 2 * (size_t)gw);
            split_d(d0.x, diy, dfy); split_d(d0.y, dix, dfx);
        }
    }

    const int rgp = lane & 7, cgp = lane >> 3;                      // patch row group / column group of this lane
    const int pi0 = rgp * PR, pj0 = cgp * PC;                       // first window row / column of the patch
    const int pi0s = min(rgp, T::RGU - 1) * PR;                     // rows this lane reads (idle row groups alias the last used one)
    // template patch: I in column pairs (2q, 2q+1) as float2 + the odd last column (packed fp32 subtraction), gradients as the
    // (Iy, Ix) pair the 64-bit load delivers: one packed FMA per pixel accumulates (by, bx) together
    constexpr int NQT = PC / 2;
    static_assert(PC % 2 == 1, "odd patch width expected");
    float2 tIp[PR][NQT];
    float tIl[PR];
    float2 tG[PR][PC];
    bool pendA = false;   // a template load is in flight on barA
    int tmpl_stage = -1;  // stage whose template has been requested (or sits in shared memory)
    const float* const tT0 = &sm.sT[pj0][pi0s];  // this lane's first tap in a tile whose first tap row / column is (0, 0)

retry:
    const int nstage = levels_cur + 1 + (MODE ? 1 : 0);
    for (int s = 0; s < nstage; ++s) {
        const bool back = s > levels_cur;
        const int lvl = back ? 0 : levels_cur - s;
        if (back) {
            // tracker.jl:37-46: the backward pass starts from the forward result with the negated displacement
            const double2 pt = *reinterpret_cast<const double2*>(a.pts + 2 * (size_t)gw);
            const double qy = pt.x + join_d(diy, dfy), qx = pt.y + join_d(dix, dfx);
            if (lane == 0) {
                sm.q[0] = qy; sm.q[1] = qx;
                if (a.out_pts) { a.out_pts[2 * (size_t)gw] = qy; a.out_pts[2 * (size_t)gw + 1] = qx; }
            }
            result = 2;
            iy0 = __double2int_rd(qy); ix0 = __double2int_rd(qx);
            if (dfy > 0.f) { diy = -diy - 1; dfy = 1.f - dfy; } else diy = -diy;
            if (dfx > 0.f) { dix = -dix - 1; dfx = 1.f - dfx; } else dix = -dix;
            if (dfy >= 1.f) { diy += 1; dfy = 0.f; }
            if (dfx >= 1.f) { dix += 1; dfx = 0.f; }
            swapped = true;
        }
        if (LKT_PF_PT && next < 0 && (back || (MODE == 0 && s == levels_cur))) {
            // the work counter's answer has long arrived: broadcast it and request the next keypoint's coordinates, so that the
            // next keypoint does not start with two dependent global round trips
            next = __shfl_sync(FULL, next_raw, 0);
            if (next < a.n_frames * a.n_per_frame) pt_next = __ldg(reinterpret_cast<const double2*>(a.pts + 2 * (size_t)next));
        }
        const float eps = back ? 1e-2f : (float)a.eps;
        const int H = a.lv[lvl].H, W = a.lv[lvl].W;
        const int py = iy0 >> lvl, px = ix0 >> lvl;  // floor(p / 2^lvl) (lucas_kanade.jl:197): floor(floor(p) / 2^lvl)
        if (MODE == 0 && lane == 0) { sm.dsave[0] = diy; sm.dsave[1] = __float_as_int(dfy); sm.dsave[2] = dix; sm.dsave[3] = __float_as_int(dfx); }
        // A keypoint closer than 2^lvl to the top / left border has level coordinate 0: get_offsets then gives up (left) = -1 and the
        // window covers rows (columns) 1 .. down -- a valid window one pixel off the point (lucas_kanade.jl:199-212), reachable only
        // with an initial displacement that brings the estimate into the level.  Everything below takes up / left = -1 as it comes.
        if (!((unsigned)py <= (unsigned)H && (unsigned)px <= (unsigned)W)) { ok = false; break; }
        int up = min(w, py - 1), down = min(w, H - py), left = min(w, px - 1), right = min(w, W - px);
        int fy = py + diy, fx = px + dix;  // the estimate is (fy + wy, fx + wx)
        float wy = dfy, wx = dfx;
        // ---- target tile around the first iteration's position: its latency overlaps the set-up below
        int ty0 = (fy - up - 1 - T::MY) & ~3;  // the innermost TMA coordinate must be a multiple of 16 bytes (tools/tma_probe.cu)
        int tx0 = fx - left - 1 - T::MXL;
        __syncwarp();
        if (lane == 0) {
            fence_async_smem();
            mbar_expect(&sm.barT, T::T_BYTES);
            tma_box(&sm.sT[0][0], &LKT_MAPS_B[lvl].tgt, ty0, tx0, LKT_SLOT_B, &sm.barT);
        }
        bool pendT = true;
        bool setup = true, first_setup = true;
        float g00 = 0.f, g01 = 0.f, g11 = 0.f;
        int it = 0, it_mark = 0;
        // fast-path box of the integer estimate: inside it the staged tile covers every tap and get_offsets(p, estimate) equals
        // the current offsets, so the estimate lies in the image and nothing has to be recomputed (see the slow path below)
        // The box also carries the tile: first tap (fy - up - 1 - ty0, fx - left - 1 - tx0) within [0, TR - RSPAN] x [0, TC - CSPAN].
        int bylo = 0x40000000, byspan = 0, bxlo = 0x40000000, bxspan = 0;  // empty
        bool orig_offsets = true;
        const float* tbase = tT0;
        auto make_box = [&]() {
            // get_offsets(p, pc) keeps the offsets of get_offsets(p, p) while floor(pc) >= min(p, w + 1) and
            // ceil(pc) <= max(p, size - w) (lucas_kanade.jl:199-208); a window that was re-clipped has no such box
            const int ty_lo = ty0 + up + 1, tx_lo = tx0 + left + 1;
            // (the floor of an estimate inside the level is >= 1: explicit for level coordinates 0, where min(p, w + 1) = 0)
            const int ylo = max(max(min(py, w + 1), 1), ty_lo), yhi = min(max(py, H - w) - 1, ty_lo + (TR - T::RSPAN));
            const int xlo = max(max(min(px, w + 1), 1), tx_lo), xhi = min(max(px, W - w) - 1, tx_lo + (TC - T::CSPAN));
            const bool some = orig_offsets && yhi >= ylo && xhi >= xlo;
            bylo = some ? ylo : 0x40000000; byspan = some ? yhi - ylo : 0;
            bxlo = some ? xlo : 0x40000000; bxspan = some ? xhi - xlo : 0;
            tbase = tT0 - (tx_lo * TR + ty_lo);
        };
        while (true) {
            if (setup) {
                const int nrows = up + down + 1, ncols = left + right + 1;
                const int r0 = py - up, c0 = px - left;
                if (nrows < 1 || ncols < 1 || r0 < 1 || c0 < 1 || py + down > H || px + right > W) { ok = false; break; }
                // ---- template tiles: requested one level ahead on the forward pass, otherwise now
                if (!(first_setup && tmpl_stage == s)) {
                    if (pendA) { mbar_wait_par(&sm.barA, parA); parA ^= 1; }
                    __syncwarp();
                    if (lane == 0) {
                        fence_async_smem();
                        mbar_expect(&sm.barA, T::I_BYTES + T::G_BYTES);
                        const int ra = (r0 - 1) & ~3;  // 16-byte aligned box start; the lanes skip (r0 - 1) & 3 rows when they read
                        tma_box(&sm.sI[0][0], &LKT_MAPS_A[lvl].ti, ra, c0 - 1, LKT_SLOT_A, &sm.barA);
                        tma_box(&sm.sG[0][0], &LKT_MAPS_A[lvl].tg, 2 * ((r0 - 1) & ~T::GSKIP), c0 - 1, LKT_SLOT_A, &sm.barA);
                    }
                    pendA = true;
                    tmpl_stage = s;
                }
                // ---- G^-1: table entry of (keypoint, level) on a level's first set-up, else from the row prefix planes (lane = row)
                bool gate_ok;
                float ra0 = 0.f, ra1 = 0.f, rb0 = 0.f, rb1 = 0.f, rc0 = 0.f, rc1 = 0.f;
                const bool from_table = first_setup && !back && gtab;
                if (from_table) {
                    if (!LKT_PF_G) gnext = __ldg(gtab + ((size_t)gw * a.gtab_levels + lvl));
                    g00 = gnext.x; g01 = gnext.y; g11 = gnext.z;
                    gate_ok = gnext.w != 0.f;
                    if (LKT_PF_G && s < levels_cur) gnext = __ldg(gtab + ((size_t)gw * a.gtab_levels + lvl - 1));
                } else {
                    // request the six prefix values of this lane's row now; they are reduced after the template has been read
                    const LKLevel& L = a.lv[lvl];
                    const float* colA = LKT_FB_A + (size_t)(r0 - 1 + min(lane, nrows - 1));
                    const size_t lo = (size_t)(c0 - 1) * L.pitch, hi = (size_t)(px + right) * L.pitch;
                    ra1 = __ldg(colA + L.oRyy + hi); ra0 = __ldg(colA + L.oRyy + lo);
                    rb1 = __ldg(colA + L.oRxx + hi); rb0 = __ldg(colA + L.oRxx + lo);
                    rc1 = __ldg(colA + L.oRyx + hi); rc0 = __ldg(colA + L.oRyx + lo);
                    gate_ok = true;
                }
                // ---- template patch of this lane into registers; patch rows beyond the window read zero gradients
                mbar_wait_par(&sm.barA, parA); parA ^= 1;
                pendA = false;
                {
                    const int skip = (r0 - 1) & 3;
                    const float* const tI = &sm.sI[pj0][pi0s] + skip;
                    const float2* const tGp = &sm.sG[pj0][pi0s] + ((r0 - 1) & T::GSKIP);
                    // the patch grid's 20th column (last column of the fourth column group) is not part of the 19-column window:
                    // with slim tiles its loads are redirected to the 19th column and its gradients zeroed
                    const bool last_ok = !LKT_SLIM || pj0 + PC - 1 < W2;
                    const int jl = last_ok ? PC - 1 : PC - 2;
#pragma unroll
                    for (int i = 0; i < PR; ++i) {
                        const bool in_win = pi0 + i < nrows;  // patch rows beyond the window carry real data: their gradients are zeroed
#pragma unroll
                        for (int q = 0; q < NQT; ++q) tIp[i][q] = make_float2(tI[(2 * q) * AR + i], tI[(2 * q + 1) * AR + i]);
                        tIl[i] = tI[jl * AR + i];
#pragma unroll
                        for (int j = 0; j < PC - 1; ++j) { const float2 gv = tGp[j * T::ARG + i]; tG[i][j] = in_win ? gv : make_float2(0.f, 0.f); }
                        { const float2 gv = tGp[jl * T::ARG + i]; tG[i][PC - 1] = (in_win && last_ok) ? gv : make_float2(0.f, 0.f); }
                    }
                    if (ncols < T::GC) {  // clipped (or smaller) window: gradient columns beyond it carry real data, zero them
#pragma unroll
                        for (int j = 0; j < PC; ++j)
                            if (pj0 + j >= ncols) {
#pragma unroll
                                for (int i = 0; i < PR; ++i) tG[i][j] = make_float2(0.f, 0.f);
                            }
                    }
                }
                if (!from_table) {
                    float syy = ra1 - ra0, sxx = rb1 - rb0, syx = rc1 - rc0;
                    if (lane >= nrows) { syy = 0.f; sxx = 0.f; syx = 0.f; }
                    float fa, fc, fb;
                    wsum3(syy, sxx, syx, lane, fa, fc, fb);
                    double d00, d01, d11;
                    gate_ok = g_inverse((double)fa, (double)fc, (double)fb, a.eig_thr * (double)(nrows * ncols), d00, d01, d11);
                    g00 = (float)d00; g01 = (float)d01; g11 = (float)d11;
                }
                if (!gate_ok) { ok = false; break; }
                // ---- request the next forward level's template: it depends only on the keypoint
                if (first_setup && !back && s < levels_cur) {
                    const int npy = iy0 >> (lvl - 1), npx = ix0 >> (lvl - 1);
                    const int nr0 = npy - min(w, npy - 1), nc0 = npx - min(w, npx - 1);
                    __syncwarp();
                    if (lane == 0) {
                        fence_async_smem();
                        mbar_expect(&sm.barA, T::I_BYTES + T::G_BYTES);
                        const int ra = (nr0 - 1) & ~3;
                        tma_box(&sm.sI[0][0], &LKT_MAPS_A[lvl - 1].ti, ra, nc0 - 1, LKT_SLOT_A, &sm.barA);
                        tma_box(&sm.sG[0][0], &LKT_MAPS_A[lvl - 1].tg, 2 * ((nr0 - 1) & ~T::GSKIP), nc0 - 1, LKT_SLOT_A, &sm.barA);
                    }
                    pendA = true;
                    tmpl_stage = s + 1;
                }
                orig_offsets = first_setup;
                make_box();
                if (pendT) { mbar_wait_par(&sm.barT, parT); parT ^= 1; pendT = false; }
                setup = false; first_setup = false;
            }
            if (it >= a.iterations) {
                // the last step was applied: lucas_kanade.jl:89 still requires the new estimate to lie in the level
                if (it > 0 && !(fy >= 1 && fy + (wy > 0.f) <= H && fx >= 1 && fx + (wx > 0.f) <= W)) ok = false;
                break;
            }
            if (!((unsigned)(fy - bylo) <= (unsigned)byspan && (unsigned)(fx - bxlo) <= (unsigned)bxspan)) {
                // exact path: lies_in (1 <= pc <= size  <=>  floor >= 1 && ceil <= size) and get_offsets on floor / ceil:
                // floor(min(w, min(p, pc) - 1)) = min(w, min(p, floor pc) - 1), floor(min(w, H - max(p, pc))) = min(w, H - max(p, ceil pc))
                const int cyi = fy + (wy > 0.f ? 1 : 0), cxi = fx + (wx > 0.f ? 1 : 0);
                if (!(fy >= 1 && cyi <= H && fx >= 1 && cxi <= W)) { ok = false; break; }
                const int nup = min(w, min(py, fy) - 1), ndown = min(w, H - max(py, cyi));
                const int nleft = min(w, min(px, fx) - 1), nright = min(w, W - max(px, cxi));
                if (nup != up || ndown != down || nleft != left || nright != right) {
                    wpx += (unsigned)((up + down + 1) * (left + right + 1) * (it - it_mark)); it_mark = it;
                    up = nup; down = ndown; left = nleft; right = nright;
                    setup = true;
                    continue;
                }
                if (!((unsigned)(fy - up - 1 - ty0) <= (unsigned)(TR - T::RSPAN) && (unsigned)(fx - left - 1 - tx0) <= (unsigned)(TC - T::CSPAN))) {
                    // the estimate walked out of the staged tile (or the window was re-clipped): stage again around it
                    __syncwarp();
                    ty0 = (fy - up - 1 - T::MY) & ~3;
                    tx0 = fx - left - 1 - T::MXL;
                    if (lane == 0) {
                        fence_async_smem();
                        mbar_expect(&sm.barT, T::T_BYTES);
                        tma_box(&sm.sT[0][0], &LKT_MAPS_B[lvl].tgt, ty0, tx0, LKT_SLOT_B, &sm.barT);
                    }
                    make_box();
                    mbar_wait_par(&sm.barT, parT); parT ^= 1;
                }
            }
            // ---- prepare_linear_system (lucas_kanade.jl:159-173) on this lane's patch
            float by, bx;
            {
                const float* tb = tbase + (fx * TR + fy);
                // packed fp32 (FFMA2 / FMUL2 / FADD2 on register pairs): tap columns in pairs (2m, 2m+1), pixels in pairs
                // (2q, 2q+1) plus the last column alone; bilinear sample as a*(1-w) + b*w so no negated operand is needed
                constexpr int NP = (PC + 1) / 2, NQ = PC / 2;
                const float omwy = 1.f - wy, omwx = 1.f - wx;
                const float2 wy2 = make_float2(wy, wy), omwy2 = make_float2(omwy, omwy);
                const float2 nwx2 = make_float2(-wx, -wx), nomwx2 = make_float2(-omwx, -omwx);
                float2 V[PR][NP];
                {
                    float2 Tp[NP];
#pragma unroll
                    for (int m = 0; m < NP; ++m) Tp[m] = make_float2(tb[(2 * m) * TR], tb[(2 * m + 1) * TR]);
#pragma unroll
                    for (int i = 0; i < PR; ++i) {
#pragma unroll
                        for (int m = 0; m < NP; ++m) {
                            const float2 Tn = make_float2(tb[(2 * m) * TR + i + 1], tb[(2 * m + 1) * TR + i + 1]);
                            V[i][m] = __ffma2_rn(Tn, wy2, __fmul2_rn(Tp[m], omwy2));
                            Tp[m] = Tn;
                        }
                    }
                }
                float2 b2[PR];  // (by, bx) per patch row: PR independent packed-FMA chains
#pragma unroll
                for (int i = 0; i < PR; ++i) {
                    b2[i] = make_float2(0.f, 0.f);
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        const float2 A = V[i][q], B = make_float2(V[i][q].y, V[i][q + 1].x);
                        const float2 nval = __ffma2_rn(B, nwx2, __fmul2_rn(A, nomwx2));
                        const float2 dI = __fadd2_rn(tIp[i][q], nval);
                        b2[i] = __ffma2_rn(tG[i][2 * q], make_float2(dI.x, dI.x), b2[i]);
                        b2[i] = __ffma2_rn(tG[i][2 * q + 1], make_float2(dI.y, dI.y), b2[i]);
                    }
                    const float val = fmaf(V[i][NP - 1].y, wx, V[i][NP - 1].x * omwx);
                    const float dI = tIl[i] - val;
                    b2[i] = __ffma2_rn(tG[i][PC - 1], make_float2(dI, dI), b2[i]);
                }
#pragma unroll
                for (int i = 1; i < PR; ++i) b2[0] = __fadd2_rn(b2[0], b2[i]);
                by = b2[0].x; bx = b2[0].y;
            }
            float sby, sbx;
            wsum2(by, bx, lane, sby, sbx);
            ++it;
            // compute_flow_vector (lucas_kanade.jl:175-187): f = G^-1 b, in fp32 like the sums it is made of
            const float ffy = fmaf(g01, sbx, g00 * sby), ffx = fmaf(g11, sbx, g01 * sby);
            if (fabsf(ffy) < eps && fabsf(ffx) < eps) break;  // converged: the step is not applied (lucas_kanade.jl:80-83)
            {
                const float ny = wy + ffy, nx = wx + ffx;
                const float ky = floorf(ny), kx = floorf(nx);
                fy += (int)ky; fx += (int)kx;
                wy = ny - ky; wx = nx - kx;
            }
        }
        wpx += (unsigned)((up + down + 1) * (left + right + 1) * (it - it_mark));
        nit += (unsigned)it;
        if (pendT) { mbar_wait_par(&sm.barT, parT); parT ^= 1; }  // early exit with the target tile still in flight
        if (!ok) break;
        diy = fy - py; dix = fx - px; dfy = wy; dfx = wx;
        if (lvl > 0) {  // the next level's coordinates are twice as fine (lucas_kanade.jl:96)
            const int ky = dfy >= 0.5f, kx = dfx >= 0.5f;
            diy = 2 * diy + ky; dfy = 2.f * dfy - (float)ky;
            dix = 2 * dix + kx; dfx = 2.f * dfx - (float)kx;
        }
    }

    if (MODE != 0 && result != 0 && ok) {
        // tracker.jl:59-66: the back-tracked point must land within max_distance of the original keypoint
        __syncwarp();
        const double2 pt = *reinterpret_cast<const double2*>(a.pts + 2 * (size_t)gw);
        const double by = sm.q[0] + join_d(diy, dfy), bx = sm.q[1] + join_d(dix, dfx);
        const double ey = pt.x - by, ex = pt.y - bx;
        if (!(sqrt(ey * ey + ex * ex) >= a.max_dist)) result = 3;
    }
    if (prior_first && !second_try && result != 3) {
        // the prior pass failed: map_manager.jl:531-536 re-queues the keypoint with the 2-D ones (no prior, all levels)
        second_try = true;
        levels_cur = a.levels;
        diy = 0; dix = 0; dfy = 0.f; dfx = 0.f;
        {
            const double2 pt = *reinterpret_cast<const double2*>(a.pts + 2 * (size_t)gw);
            iy0 = __double2int_rd(pt.x); ix0 = __double2int_rd(pt.y);
        }
        ok = true; result = 0;
        swapped = false;
        tmpl_stage = -1;
        if (LKT_PF_G && gtab) gnext = __ldg(gtab + ((size_t)gw * a.gtab_levels + levels_cur));
        __syncwarp();
        goto retry;
    }
    // never leave with copies in flight (early exits only: a completed keypoint has consumed everything it requested)
    if (pendA) { mbar_wait_par(&sm.barA, parA); parA ^= 1; }
    if (MODE == 0) {
        if (lane == 0) {
            if (!ok) {  // a failed level leaves d[n] as it was (lucas_kanade.jl:43,53,67,89)
                diy = sm.dsave[0]; dfy = __int_as_float(sm.dsave[1]); dix = sm.dsave[2]; dfx = __int_as_float(sm.dsave[3]);
            }
            if (a.disp_out) { a.disp_out[2 * (size_t)gw] = join_d(diy, dfy); a.disp_out[2 * (size_t)gw + 1] = join_d(dix, dfx); }
            a.status[gw] = ok ? 1 : 0;
        }
    } else if (lane == 0) {
        if (result == 0 && a.out_pts) {  // forward pass failed; the reference leaves new_keypoints[i] undefined
            a.out_pts[2 * (size_t)gw] = __longlong_as_double(0x7ff8000000000000LL);
            a.out_pts[2 * (size_t)gw + 1] = __longlong_as_double(0x7ff8000000000000LL);
        }
        a.status[gw] = result | ((prior_first && !second_try && result == 3) ? 4 : 0);  // bit2: tracked by the prior pass
    }
    acc_wpx += wpx; acc_nit += nit;  // flushed once per CTA by the kernel
}

// resident one-warp CTAs per SM the register allocation is bounded for.  Measured on B200 (64 x 2000 keypoints, forward-backward
// kernel): 16 (110 registers) 0.840 ms, 18 (94 registers, no spills; shared memory allows 18) 0.806 ms.  A 24-row target tile
// (LKT_TR=24, 9.5 KB of shared memory) reaches 21-24 CTAs but re-stages so often that it loses: 0.976 ms at 16, 0.885 ms at 24.
// The optflow! / optical_flow_matching! instantiations carry more state and keep the 128-register bound.
#ifndef LKT_MINB
#define LKT_MINB 20
#endif

// Kernel: one warp (= one CTA) per keypoint; the grid is persistent (one-warp CTAs filling every SM, or one per keypoint when there
// are fewer keypoints than warp slots) and every warp draws keypoint indices from a device counter -- the next index is requested
// before the current keypoint is processed, so the atomic's latency is hidden.
template <int W2, int PR, int PC, int MODE>
__global__ void __launch_bounds__(32, MODE == 1 ? LKT_MINB : 16) k_lk_tma(const LKArgs a) {
    __shared__ LKTmaSmem<W2, PR, PC> sm;
    const int lane = threadIdx.x;
    // the gradient tile's columns beyond the box are never written by the TMA, nor is the zero row: clear them once
    for (int i = lane; i < TmaTile<W2, PR, PC>::GCS * TmaTile<W2, PR, PC>::ARG; i += 32) (&sm.sG[0][0])[i] = make_float2(0.f, 0.f);
    if (lane == 0) {
        mbar_init1(&sm.barT);
        mbar_init1(&sm.barA);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned parT = 0, parA = 0;
    const int total = a.n_frames * a.n_per_frame;
    int base = 0;
    if (lane == 0) base = (int)atomicAdd(a.work, 1u);
    base = __shfl_sync(FULL, base, 0);
    double2 pt = make_double2(0.0, 0.0);
    unsigned long long acc_wpx = 0, acc_nit = 0;  // executed window-pixel iterations / iterations of this CTA's keypoints
    if (base < total) pt = __ldg(reinterpret_cast<const double2*>(a.pts + 2 * (size_t)base));
    while (base < total) {
        int next_raw = 0;
        if (lane == 0) next_raw = (int)atomicAdd(a.work, 1u);
        int next = -1;
        double2 pt_next = make_double2(0.0, 0.0);
        lk_point_tma<W2, PR, PC, MODE>(a, base, lane, sm, parT, parA, pt, next_raw, next, pt_next, acc_wpx, acc_nit);
        __syncwarp();
        if (next < 0) {  // the keypoint ended early (failed before its last stage): fetch the next one here
            next = __shfl_sync(FULL, next_raw, 0);
            if (next < total) pt_next = __ldg(reinterpret_cast<const double2*>(a.pts + 2 * (size_t)next));
        }
        base = next;
        pt = pt_next;
    }
    if (lane == 0 && a.counters && acc_nit) {
        atomicAdd(a.counters, acc_wpx);
        atomicAdd(a.counters + 1, acc_nit);
    }
}

// G table for the forward pass (and for both passes of mode 2: the entries depend only on keypoint and level)
static int launch_lk_gprep(cudaStream_t s, const LKArgs& a) {
    const long long items = (long long)a.n_frames * a.n_per_frame * a.gtab_levels;
    if (items <= 0 || !a.gtab) return 0;
    const long long threads = items * 8;
    k_lk_gprep<<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(a, (GEntry*)a.gtab);
    return 1;
}

// returns false when this variant does not cover the request (window size, or no tensor maps / work counter)
bool launch_lk_tma(cudaStream_t s, const LKArgs& a) {
    const int total = a.n_frames * a.n_per_frame;
    const int w2 = 2 * a.window + 1;
    if (w2 > 19 || !a.mapsA || !a.mapsB || !a.work || a.mode < 0 || a.mode > 2) return false;
    static const int slots = [] {
        int dev = 0, sms = 148, per_sm = 1;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(k_lk_tma<19, 3, 5, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_lk_tma<19, 3, 5, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_lk_tma<19, 3, 5, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_lk_tma<19, 3, 5, 1>, 32, 0);
        const char* e = getenv("SLAMKLT_LK_SLOTS");  // experiment knob: resident one-warp CTAs per SM of the persistent grid
        const int want = e ? atoi(e) : per_sm;
        return sms * (want >= 1 && want <= per_sm ? want : per_sm);
    }();
    if ((unsigned long long)total * (unsigned long long)a.n_per_frame >= (1ull << 40)) return false;  // (the frame-index multiply)
    const int grid = total < slots ? total : slots;
    cudaMemsetAsync(a.work, 0, sizeof(unsigned), s);
    LKArgs b = a;
    b.npf_magic = ((1ull << 40) + (unsigned long long)a.n_per_frame - 1) / (unsigned long long)a.n_per_frame;
    if (b.gtab) launch_lk_gprep(s, b);
    if (b.mode == 0) k_lk_tma<19, 3, 5, 0><<<grid, 32, 0, s>>>(b);
    else if (b.mode == 1) k_lk_tma<19, 3, 5, 1><<<grid, 32, 0, s>>>(b);
    else k_lk_tma<19, 3, 5, 2><<<grid, 32, 0, s>>>(b);
    return true;
}

}  // namespace sk
