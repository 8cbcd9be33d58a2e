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
#ifndef LKT_MX
#define LKT_MX 4
#endif

template <int W2, int PR, int PC>
struct TmaTile {
    static constexpr int RGU = (W2 + PR - 1) / PR;            // row groups that own window rows (7 of 8 for 19 rows)
    static constexpr int RSPAN = (RGU - 1) * PR + PR + 1;     // tap rows touched (idle row groups alias the last used one)
    static constexpr int CSPAN = PC * 4 + 1;                  // tap columns touched
    static constexpr int TR = LKT_TR, MX = LKT_MX;
    static constexpr int MY = (TR - RSPAN) / 2;               // margin above the first tap row when a tile is staged
    static constexpr int TC = CSPAN + 2 * MX;
    static constexpr int AR = ((RGU * PR + 3 + 3) / 4) * 4;   // template tile rows: RGU*PR read + up to 3 skipped for the 16-byte aligned box start
    static constexpr int AC = PC * 4;                         // template tile columns in shared memory
    static constexpr int GC = W2;                             // gradient box columns: columns >= GC of the tile stay zero
    static constexpr unsigned T_BYTES = TC * TR * 4, I_BYTES = AC * AR * 4, G_BYTES = GC * AR * 8;
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
        if (one(&out[l].tg, base + plane_off(L, DP_GRAD), (cuuint64_t)L.pitch * 2, 2 * T::AR, T::GC)) return -1;
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
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

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
// of the tracking warp at every level.  Entry: g00, g01, g11 and a flag (1 = gate passed, 0 = failed / window empty).
struct __align__(32) GEntry { double g00, g01, g11; long long ok; };

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
        GEntry e;
        e.g00 = e.g01 = e.g11 = 0.0; e.ok = 0;
        if (valid && g_inverse((double)syy, (double)sxx, (double)syx, a.eig_thr * (double)(nrows * ncols), e.g00, e.g01, e.g11)) e.ok = 1;
        tab[item] = e;
    }
}

// ---- tracking kernel --------------------------------------------------------------------------------------------------------
template <int W2, int PR, int PC>
struct __align__(128) LKTmaSmem {
    using T = TmaTile<W2, PR, PC>;
    float sT[T::TC][T::TR];
    __align__(128) float sI[T::AC][T::AR];
    __align__(128) float2 sG[T::AC][T::AR];
    __align__(16) double dsave[2];  // displacement at the start of the current level (optflow! keeps it when the level fails)
    uint64_t barT;
    uint64_t barA;
};

// one keypoint, one warp: all levels of the forward pass, the backward pass and the gates
template <int W2, int PR, int PC>
__device__ __forceinline__ void lk_point_tma(const LKArgs& a, const int gw, const int lane, LKTmaSmem<W2, PR, PC>& sm, unsigned& parT, unsigned& parA) {
    using T = TmaTile<W2, PR, PC>;
    constexpr int TR = T::TR, TC = T::TC, AR = T::AR;
    const int f = gw / a.n_per_frame;
    // template side (A) and target side (B) swap roles on the backward pass: keep the two physical slots and derive pointers
    // and tensor maps where they are used (once per set-up) instead of carrying them in registers
    const int slot_first = a.A.slot(a.offA + f), slot_second = a.B.slot(a.offB + f);
    bool swapped = false;
#define LKT_SLOT_A (swapped ? slot_second : slot_first)
#define LKT_SLOT_B (swapped ? slot_first : slot_second)
#define LKT_MAPS_A (reinterpret_cast<const LKTmaLevel*>(swapped ? a.mapsB : a.mapsA))
#define LKT_MAPS_B (reinterpret_cast<const LKTmaLevel*>(swapped ? a.mapsA : a.mapsB))
#define LKT_FB_A (swapped ? a.B.base + (size_t)slot_second * a.B.frame_elems : a.A.base + (size_t)slot_first * a.A.frame_elems)
    double dy = 0.0, dx = 0.0;
    if (a.disp_in) { dy = a.disp_in[2 * (size_t)gw]; dx = a.disp_in[2 * (size_t)gw + 1]; }
    // mode 2 = optical_flow_matching! (map_manager.jl:451-564): keypoints with a prior (3-D keypoints) are first tracked with
    // that prior on levels3d levels; those that fail, and all others, are tracked from a zero displacement on all levels
    const uint8_t prior_flag = (a.mode == 2 && a.has_prior) ? a.has_prior[gw] : (uint8_t)0;
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
    const bool prior_first = prior_flag != 0;
    if (a.mode == 2 && !prior_first) { dy = 0.0; dx = 0.0; }
    int levels_cur = prior_first ? a.levels3d : a.levels;
    bool second_try = false;
    const GEntry* gtab = reinterpret_cast<const GEntry*>(a.gtab);

    const int w = a.window;
    unsigned int wpx = 0, nit = 0;
    double qy = a.pts[2 * (size_t)gw], qx = a.pts[2 * (size_t)gw + 1];
    bool ok = true;
    uint8_t result = 0;

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
    float2 rm[PR];  // row weights: 1 for window rows, 0 for patch rows beyond the (clipped) window
    int ty0 = 0, tx0 = 0;
    bool pendT = false, pendA = false;  // a target / template load is in flight on barT / barA
    int tmpl_stage = -1;                // stage whose template has been requested (or sits in shared memory)
    const float* const tI0 = &sm.sI[pj0][pi0s];
    const float2* const tG0 = &sm.sG[pj0][pi0s];
    const float* const tT0 = &sm.sT[pj0][pi0s];

retry:
    const int nstage = levels_cur + 1 + (a.mode ? 1 : 0);
    for (int s = 0; s < nstage; ++s) {
        const bool back = s > levels_cur;
        const int lvl = back ? 0 : levels_cur - s;
        if (back) {
            qy += dy; qx += dx;  // tracker.jl:37-46
            if (lane == 0 && a.out_pts) { a.out_pts[2 * (size_t)gw] = qy; a.out_pts[2 * (size_t)gw + 1] = qx; }
            result = 2;
            dy = -dy; dx = -dx;
            swapped = true;
        }
        const LKLevel& L = a.lv[lvl];
        const int H = L.H, W = L.W, pitch = L.pitch;
        const double inv = __longlong_as_double((long long)(1023 - lvl) << 52);  // 2^-lvl exactly, no division
        const int py = (int)floor(qy * inv), px = (int)floor(qx * inv);
        int up = min(w, py - 1), down = min(w, H - py), left = min(w, px - 1), right = min(w, W - px);
        const bool interior = py - 1 >= w + 6 && H - py >= w + 6 && px - 1 >= w + 6 && W - px >= w + 6;
        bool setup = true, first_setup = true;
        double g00 = 0, g01 = 0, g11 = 0;
        int it = 0;  // dy, dx run along with the iterations (they hold d + c of lucas_kanade.jl:50-90)
        if (a.mode == 0 && lane == 0) { sm.dsave[0] = dy; sm.dsave[1] = dx; }
        {
            // request the target tile around the first iteration's position now: its latency overlaps the set-up below
            const int fy0 = __double2int_rd((double)py + dy), fx0 = __double2int_rd((double)px + dx);
            ty0 = (fy0 - up - 1 - T::MY) & ~3;  // the innermost TMA coordinate must be a multiple of 16 bytes (tools/tma_probe.cu)
            tx0 = fx0 - left - 1 - T::MX;
            if (pendT) { mbar_wait_par(&sm.barT, parT); parT ^= 1; }  // (only after an early exit left a load in flight)
            __syncwarp();
            if (lane == 0) {
                fence_async_smem();
                mbar_expect(&sm.barT, T::T_BYTES);
                tma_box(&sm.sT[0][0], &LKT_MAPS_B[lvl].tgt, ty0, tx0, LKT_SLOT_B, &sm.barT);
            }
            pendT = true;
        }
        while (true) {
            const int nrows = up + down + 1, ncols = left + right + 1;
            if (setup) {
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
                        tma_box(&sm.sG[0][0], &LKT_MAPS_A[lvl].tg, 2 * ra, c0 - 1, LKT_SLOT_A, &sm.barA);
                    }
                    pendA = true;
                    tmpl_stage = s;
                }
                // ---- G: table entry of (keypoint, level) on a level's first set-up, else from the row prefix planes (lane = row)
                bool gate_ok;
                if (first_setup && !back && gtab) {
                    const GEntry* e = gtab + ((size_t)gw * a.gtab_levels + lvl);
                    const double2 v0 = __ldg(reinterpret_cast<const double2*>(e));
                    const double2 v1 = __ldg(reinterpret_cast<const double2*>(e) + 1);
                    g00 = v0.x; g01 = v0.y; g11 = v1.x;
                    gate_ok = __double_as_longlong(v1.y) != 0;
                } else {
                    float syy, sxx, syx;
                    {
                        const float* colA = LKT_FB_A + (size_t)(r0 - 1 + min(lane, nrows - 1));
                        const size_t lo = (size_t)(c0 - 1) * pitch, hi = (size_t)(px + right) * pitch;
                        syy = __ldg(colA + L.oRyy + hi) - __ldg(colA + L.oRyy + lo);
                        sxx = __ldg(colA + L.oRxx + hi) - __ldg(colA + L.oRxx + lo);
                        syx = __ldg(colA + L.oRyx + hi) - __ldg(colA + L.oRyx + lo);
                        if (lane >= nrows) { syy = 0.f; sxx = 0.f; syx = 0.f; }
                    }
                    float fa, fc, fb;
                    wsum3(syy, sxx, syx, lane, fa, fc, fb);
                    gate_ok = g_inverse((double)fa, (double)fc, (double)fb, a.eig_thr * (double)(nrows * ncols), g00, g01, g11);
                }
                // ---- template patch of this lane into registers
                mbar_wait_par(&sm.barA, parA); parA ^= 1;
                pendA = false;
                const float* const tI = tI0 + ((r0 - 1) & 3);
                const float2* const tGp = tG0 + ((r0 - 1) & 3);
#pragma unroll
                for (int i = 0; i < PR; ++i) {
                    const float m = (pi0 + i < nrows) ? 1.f : 0.f;
                    rm[i] = make_float2(m, m);
#pragma unroll
                    for (int q = 0; q < NQT; ++q) tIp[i][q] = make_float2(tI[(2 * q) * AR + i], tI[(2 * q + 1) * AR + i]);
                    tIl[i] = tI[(PC - 1) * AR + i];
#pragma unroll
                    for (int j = 0; j < PC; ++j) tG[i][j] = tGp[j * AR + i];
                }
                if (ncols < T::GC) {  // clipped (or smaller) window: gradient columns beyond it carry real data, zero them
#pragma unroll
                    for (int j = 0; j < PC; ++j)
                        if (pj0 + j >= ncols) {
#pragma unroll
                            for (int i = 0; i < PR; ++i) tG[i][j] = make_float2(0.f, 0.f);
                        }
                }
                if (!gate_ok) { ok = false; break; }
                // ---- request the next forward level's template: it depends only on the keypoint
                if (first_setup && !back && s < levels_cur) {
                    const int nl = lvl - 1;
                    const double ninv = __longlong_as_double((long long)(1023 - nl) << 52);
                    const int npy = (int)floor(qy * ninv), npx = (int)floor(qx * ninv);
                    const int nr0 = npy - min(w, npy - 1), nc0 = npx - min(w, npx - 1);
                    __syncwarp();
                    if (lane == 0) {
                        fence_async_smem();
                        mbar_expect(&sm.barA, T::I_BYTES + T::G_BYTES);
                        const int ra = (nr0 - 1) & ~3;
                        tma_box(&sm.sI[0][0], &LKT_MAPS_A[nl].ti, ra, nc0 - 1, LKT_SLOT_A, &sm.barA);
                        tma_box(&sm.sG[0][0], &LKT_MAPS_A[nl].tg, 2 * ra, nc0 - 1, LKT_SLOT_A, &sm.barA);
                    }
                    pendA = true;
                    tmpl_stage = s + 1;
                }
                setup = false; first_setup = false;
            }
            if (it >= a.iterations) break;
            const double pcy = (double)py + dy, pcx = (double)px + dx;
            const int fy = __double2int_rd(pcy), fx = __double2int_rd(pcx);
            // fast path: the keypoint sits >= w+6 px inside the level, its window is unclipped and the estimate is within 3 px of
            // it, so the estimate lies in the image and get_offsets(point, estimate) is (w, w, w, w) as before: nothing to recompute
            const bool fast = interior && nrows == 2 * w + 1 && ncols == 2 * w + 1 && (unsigned)(fy - py + 3) <= 6u && (unsigned)(fx - px + 3) <= 6u;
            if (!fast) {
                // floor / ceil as integers serve both lies_in (1 <= pc <= size <=> floor >= 1 && ceil <= size) and get_offsets:
                // floor(min(w, min(p, pc) - 1)) = min(w, min(p, floor pc) - 1), floor(min(w, H - max(p, pc))) = min(w, H - max(p, ceil pc))
                const int cyi = __double2int_ru(pcy), cxi = __double2int_ru(pcx);
                if (!(fy >= 1 && cyi <= H && fx >= 1 && cxi <= W)) { ok = false; break; }
                const int nup = min(w, min(py, fy) - 1), ndown = min(w, H - max(py, cyi));
                const int nleft = min(w, min(px, fx) - 1), nright = min(w, W - max(px, cxi));
                if (nup != up || ndown != down || nleft != left || nright != right) {
                    up = nup; down = ndown; left = nleft; right = nright;
                    setup = true;
                    continue;
                }
            }
            const float wy = (float)(pcy - (double)fy), wx = (float)(pcx - (double)fx);
            const int ay = fy - up - 1, ax = fx - left - 1;  // 0-based first tap row / column
            int oy = ay - ty0, ox = ax - tx0;
            if (oy < 0 || oy + T::RSPAN > TR || ox < 0 || ox + T::CSPAN > TC) {
                // the estimate walked out of the staged tile (or the window was re-clipped): stage again around it
                if (pendT) { mbar_wait_par(&sm.barT, parT); parT ^= 1; }
                __syncwarp();
                ty0 = (ay - T::MY) & ~3;
                tx0 = ax - T::MX;
                oy = ay - ty0; ox = T::MX;
                if (lane == 0) {
                    fence_async_smem();
                    mbar_expect(&sm.barT, T::T_BYTES);
                    tma_box(&sm.sT[0][0], &LKT_MAPS_B[lvl].tgt, ty0, tx0, LKT_SLOT_B, &sm.barT);
                }
                pendT = true;
            }
            if (pendT) { mbar_wait_par(&sm.barT, parT); parT ^= 1; pendT = false; }
            // ---- prepare_linear_system (lucas_kanade.jl:159-173) on this lane's patch
            const float* tb = tT0 + (ox * TR + oy);
            float by, bx;
            {
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
                float2 acc = __fmul2_rn(b2[0], rm[0]);
#pragma unroll
                for (int i = 1; i < PR; ++i) acc = __ffma2_rn(b2[i], rm[i], acc);
                by = acc.x; bx = acc.y;
            }
            float fby, fbx;
            wsum2(by, bx, lane, fby, fbx);
            const double sby = (double)fby, sbx = (double)fbx;
            wpx += (unsigned)(nrows * ncols);
            nit += 1;
            ++it;
            const double ffy = g00 * sby + g01 * sbx, ffx = g01 * sby + g11 * sbx;
            const double eps = back ? 1e-2 : a.eps;
            if (fabs(ffy) < eps && fabs(ffx) < eps) break;
            dy += ffy; dx += ffx;
            if (!(fast && fabs(ffy) < 2.0 && fabs(ffx) < 2.0)) {  // on the fast path a step below 2 px cannot leave the image
                const double ny = pcy + ffy, nx = pcx + ffx;
                if (!(__double2int_rd(ny) >= 1 && __double2int_ru(ny) <= H && __double2int_rd(nx) >= 1 && __double2int_ru(nx) <= W)) { ok = false; break; }
            }
        }
        if (!ok) break;
        if (lvl > 0) { dy *= 2.0; dx *= 2.0; }
    }

    if (a.mode != 0 && result != 0 && ok) {
        // tracker.jl:59-66: the back-tracked point must land within max_distance of the original keypoint
        const double by = qy + dy, bx = qx + dx;
        const double ey = a.pts[2 * (size_t)gw] - by, ex = a.pts[2 * (size_t)gw + 1] - bx;
        if (!(sqrt(ey * ey + ex * ex) >= a.max_dist)) result = 3;
    }
    if (prior_first && !second_try && result != 3) {
        // the prior pass failed: map_manager.jl:531-536 re-queues the keypoint with the 2-D ones (no prior, all levels)
        second_try = true;
        levels_cur = a.levels;
        dy = 0.0; dx = 0.0; qy = a.pts[2 * (size_t)gw]; qx = a.pts[2 * (size_t)gw + 1];
        ok = true; result = 0;
        swapped = false;
        tmpl_stage = -1;
        goto retry;
    }
    // never leave with copies in flight (early exits only: a completed keypoint has consumed everything it requested)
    if (pendT) { mbar_wait_par(&sm.barT, parT); parT ^= 1; }
    if (pendA) { mbar_wait_par(&sm.barA, parA); parA ^= 1; }
    if (a.mode == 0) {
        if (lane == 0) {
            if (!ok) { dy = sm.dsave[0]; dx = sm.dsave[1]; }  // a failed level leaves d[n] as it was (lucas_kanade.jl:43,53,67,89)
            if (a.disp_out) { a.disp_out[2 * (size_t)gw] = dy; a.disp_out[2 * (size_t)gw + 1] = dx; }
            a.status[gw] = ok ? 1 : 0;
        }
    } else if (lane == 0) {
        if (result == 0 && a.out_pts) {  // forward pass failed; the reference leaves new_keypoints[i] undefined
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

#ifndef LKT_MINB
#define LKT_MINB 16
#endif

// Kernel: one warp (= one CTA) per keypoint.  With a.work != nullptr the grid is persistent (one-warp CTAs filling every SM) and
// every warp draws keypoint indices from a device counter -- the next index is requested before the current keypoint is
// processed, so the atomic's latency is hidden; without it CTA i handles keypoint i.
template <int W2, int PR, int PC>
__global__ void __launch_bounds__(32, LKT_MINB) k_lk_tma(const LKArgs a) {
    __shared__ LKTmaSmem<W2, PR, PC> sm;
    const int lane = threadIdx.x;
    // the gradient tile's columns beyond the box are never written by the TMA: zero the tile once
    for (int i = lane; i < TmaTile<W2, PR, PC>::AC * TmaTile<W2, PR, PC>::AR; i += 32) (&sm.sG[0][0])[i] = make_float2(0.f, 0.f);
    if (lane == 0) {
        mbar_init1(&sm.barT);
        mbar_init1(&sm.barA);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned parT = 0, parA = 0;
    const int total = a.n_frames * a.n_per_frame;
    if (a.work == nullptr) {
        const int gw = blockIdx.x;
        if (gw < total) lk_point_tma<W2, PR, PC>(a, gw, lane, sm, parT, parA);
        return;
    }
    int base = 0;
    if (lane == 0) base = (int)atomicAdd(a.work, 1u);
    base = __shfl_sync(FULL, base, 0);
    while (base < total) {
        int next = 0;
        if (lane == 0) next = (int)atomicAdd(a.work, 1u);
        lk_point_tma<W2, PR, PC>(a, base, lane, sm, parT, parA);
        __syncwarp();
        base = __shfl_sync(FULL, next, 0);
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

// returns false when this variant does not cover the request (window size, or no tensor maps)
bool launch_lk_tma(cudaStream_t s, const LKArgs& a) {
    const int total = a.n_frames * a.n_per_frame;
    const int w2 = 2 * a.window + 1;
    if (w2 > 19 || !a.mapsA || !a.mapsB) return false;
    auto kern = k_lk_tma<19, 3, 5>;
    static const int slots = [&] {
        int dev = 0, sms = 148, per_sm = 1;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, 0);
        const char* e = getenv("SLAMKLT_LK_SLOTS");  // experiment knob: resident one-warp CTAs per SM of the persistent grid
        const int want = e ? atoi(e) : per_sm;
        return sms * (want >= 1 && want <= per_sm ? want : per_sm);
    }();
    LKArgs b = a;
    static const bool persistent = getenv("SLAMKLT_LK_STATIC") == nullptr;
    if (!persistent) b.work = nullptr;
    int grid = total;
    if (b.work) {
        if (total > slots) { grid = slots; cudaMemsetAsync(b.work, 0, sizeof(unsigned), s); }
        else b.work = nullptr;  // fewer keypoints than warp slots: one CTA each
    }
    if (b.gtab) launch_lk_gprep(s, b);
    kern<<<grid, 32, 0, s>>>(b);
    return true;
}

}  // namespace sk
