// Pyramidal Lucas-Kanade with forward-backward check, patch-mapped variant (sm_100a).
//
// Same semantics as lk.cu (reference: lucas_kanade.jl:9-212, utils.jl:5-45, tracker.jl:17-68); different mapping.
// lk.cu gives every window ROW to a lane: a 19-row window keeps 20 of 32 lanes busy and needs one shuffle per tap.
// Here the 32 lanes tile the window as 8 row-groups x 4 column-groups; a lane owns a PR x PC patch of the template
// (3 x 5 for the 19 x 19 window) in registers and reads its own (PR+1) x (PC+1) bilinear taps from a target tile staged
// in shared memory: no shuffles in the sampling loop, all lanes busy, loads with immediate offsets.
// The target tile (TC columns x TR rows around the current estimate) is staged once per level with 16-byte loads and
// re-staged only if the estimate walks out of its margin.
#include <cstdlib>

#include "common.cuh"

namespace sk {

template <typename T>
__device__ __forceinline__ const T* col_ptr2(const T* base, unsigned stride_bytes, unsigned k) {
    return reinterpret_cast<const T*>(reinterpret_cast<const char*>(base) + (unsigned long long)stride_bytes * k);
}

__device__ __forceinline__ float warp_sum_f2(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// Sums of two / three per-lane values over the warp with shared butterfly steps: after the first exchange the lower and
// upper half-warps carry different quantities, so the remaining steps reduce both at once (7 shuffles instead of 10, 9
// instead of 15).  Every total is the same butterfly tree as warp_sum_f2 computes => bit-identical results.
__device__ __forceinline__ void warp_sum2(float a, float b, int lane, float& sa, float& sb) {
    const bool hi = lane & 16;
    float v = (hi ? b : a) + __shfl_xor_sync(FULL, hi ? a : b, 16);  // lower half: a-partials, upper half: b-partials
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    sa = __shfl_sync(FULL, v, 0);
    sb = __shfl_sync(FULL, v, 16);
}
__device__ __forceinline__ void warp_sum3(float a, float b, float c, int lane, float& sa, float& sb, float& sc) {
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

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
// ---- TMA (bulk async copy engine) staging: one cp.async.bulk per tile column, completion on a per-warp mbarrier
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!done);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Measured on B200 (64 x 2000 keypoints): TMA bulk staging (UBLKCP, 29 copies of 160 B per tile, L2-sourced, mbarrier wait)
// 1.38 ms vs 1.14 ms for 16-byte cp.async.ca (LDGSTS, L1-allocating): neighbouring keypoints' tiles overlap, so the L1 hits
// are worth more than the saved issue slots.  The TMA path stays selectable (-DLK_TMA=1) and is covered by the same tests.
#ifndef LK_TMA
#define LK_TMA 0
#endif
// warp sums with shared butterfly steps (fewer shuffles, one more dependent step) or plain butterflies.  Measured with the
// packed loop: shared 1.075 ms, plain 1.089 ms (the shared-memory / shuffle pipe is the scarcer resource)
#ifndef LK_SHARED_SUMS
#define LK_SHARED_SUMS 1
#endif
// packed fp32 arithmetic in the window loop (sm_100 FFMA2 / FADD2 / FMUL2): loop body 144 -> 100 instructions, 1.111 -> 1.075 ms
#ifndef LK_F32X2
#define LK_F32X2 1
#endif
// warps (= keypoints) per CTA and the occupancy the register allocation is bounded for (65536 / (32 * LK_WPB * LK_MINB) registers);
// LK_MAXNREG caps the registers directly instead.  Measured on B200 (64 x 2000 keypoints): 128 registers / 16 warps per SM 1.11 ms;
// uncapped the kernel wants 160 registers; 120 (16 warps) 1.11 ms, 112 (18 warps, 24 B spill) 1.18 ms, 104 (18 warps) 1.25 ms,
// 96 (20 warps, ~200 B spill) slower still: the spills cost more than the extra warps hide.  Fewer resident warps through
// SLAMKLT_LK_PAD_KB: 12 warps 1.34 ms, 8 warps 2.01 ms.  CTA size at 16 warps per SM: 4 / 2 / 1 warps per CTA 1.074 / 1.055 / 0.996 ms --
// keypoints need very different numbers of iterations, and a CTA keeps its slot until its slowest warp is done, so one-warp CTAs
// (no barrier is used anywhere in the kernel) keep all 16 warp slots busy.
#ifndef LK_WPB
#define LK_WPB 1
#endif
#ifndef LK_MINB
#define LK_MINB (16 / LK_WPB)
#endif

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// tile geometry of the patch kernel: rows a multiple of 4 (16-byte staging), >= RSPAN + 8, and PC*TR = 8 (mod 32) so that the
// 8 x 4 lane grid (row stride PR = 3 words, column-group stride PC*TR words) hits 32 distinct banks: TR = 40 for PC = 5, 44 for 6
template <int PR, int PC>
struct PatchTile {
    static constexpr int RSPAN = PR * 8 + 1, CSPAN = PC * 4 + 1;  // tap rows / columns touched by the lanes
    static constexpr int TR = PC == 5 ? 40 : 44;
    static constexpr int TC = CSPAN + 8;
    static_assert(TR >= RSPAN + 8 && TR % 4 == 0 && (PC * TR) % 32 == 8 && PR == 3, "tile geometry");
};

// one keypoint, one warp: all levels of the forward pass, the backward pass and the gates
template <int W2, int PR, int PC>
__device__ __forceinline__ void lk_point(const LKArgs& a, const int gw, const int lane, float (*sT)[PatchTile<PR, PC>::TR], uint64_t* bar,
                                         unsigned& parity) {
    static_assert(PR * 8 >= W2 && PC * 4 >= W2, "patch grid must cover the window");
    constexpr int RSPAN = PatchTile<PR, PC>::RSPAN, CSPAN = PatchTile<PR, PC>::CSPAN;
    constexpr int TR = PatchTile<PR, PC>::TR, TC = PatchTile<PR, PC>::TC;
    constexpr int RG = TR / 4, CGN = 32 / RG;               // staging: RG row groups per column, CGN columns per instruction
    const int f = gw / a.n_per_frame;
    const float* const fbA0 = a.A.frame(a.offA + f);
    const float* const fbB0 = a.B.frame(a.offB + f);
    const double pty = a.pts[2 * (size_t)gw], ptx = a.pts[2 * (size_t)gw + 1];
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

    const int w = a.window;
    unsigned int wpx = 0, nit = 0;
    double qy = pty, qx = ptx;
    bool ok = true;
    uint8_t result = 0;
    const float* fbA = fbA0;
    const float* fbB = fbB0;

    const int rgp = lane & 7, cgp = lane >> 3;  // patch row group / column group of this lane
    const int pi0 = rgp * PR, pj0 = cgp * PC;   // first window row / column of the patch
    // template patch: I in column pairs (2q, 2q+1) as float2 + the odd last column (packed fp32 subtraction), gradients as the
    // (Iy, Ix) pair the 64-bit load delivers: one packed FMA per pixel accumulates (by, bx) together
    constexpr int NQT = PC / 2;
    float2 tIp[PR][NQT];
    float tIl[PR];
    float2 tG[PR][PC];
    (void)tIl;
#define TI_REF(i, j) (*((j) < 2 * NQT ? (((j) & 1) ? &tIp[i][(j) / 2 < NQT ? (j) / 2 : 0].y : &tIp[i][(j) / 2 < NQT ? (j) / 2 : 0].x) : &tIl[i]))
    int ty0 = 0, tx0 = 0;
    bool pending = false;
    const int srg = lane % RG, scg = lane / RG;  // staging role of this lane: 16-byte row group / column group

retry:
    const int nstage = levels_cur + 1 + (a.mode ? 1 : 0);
    for (int s = 0; s < nstage; ++s) {
        const bool back = s > levels_cur;
        const int lvl = back ? 0 : levels_cur - s;
        if (back) {
            qy = pty + dy; qx = ptx + dx;  // tracker.jl:37-46
            if (lane == 0 && a.out_pts) { a.out_pts[2 * (size_t)gw] = qy; a.out_pts[2 * (size_t)gw + 1] = qx; }
            result = 2;
            dy = -dy; dx = -dx;
            const float* t = fbA; fbA = fbB; fbB = t;
        }
        const double eps = back ? 1e-2 : a.eps;
        const LKLevel& L = a.lv[lvl];
        const int H = L.H, W = L.W, pitch = L.pitch;
        const unsigned pitch4 = (unsigned)L.pitch * 4u;
        const double inv = __longlong_as_double((long long)(1023 - lvl) << 52);  // 2^-lvl exactly, no division
        const int py = (int)floor(qy * inv), px = (int)floor(qx * inv);
        int up = min(w, py - 1), down = min(w, H - py), left = min(w, px - 1), right = min(w, W - px);
        const bool interior = py - 1 >= w + 6 && H - py >= w + 6 && px - 1 >= w + 6 && W - px >= w + 6;
        bool setup = true;
        double g00 = 0, g01 = 0, g11 = 0;
        double cy = 0.0, cx = 0.0;
        int it = 0;
        {
            // stage the target tile for the first iteration's position now, so that its latency overlaps the G and
            // template loads and the 2x2 solve below (asynchronous global->shared copies, no registers involved)
            const int fy0 = __double2int_rd((double)py + dy), fx0 = __double2int_rd((double)px + dx);
            // the estimate may be anywhere (bounds are checked later, lies_in): clamp so the copies stay inside the frame block
            const int ay0 = min(max(fy0 - up - 1, 0), H), ax0 = min(max(fx0 - left - 1, 0), W);
            if (pending) {
                if (LK_TMA) { mbar_wait(bar, parity); parity ^= 1; } else cp_async_wait_all();
            }
            __syncwarp();
            ty0 = max(0, ay0 - 4) & ~3;
            tx0 = max(0, ax0 - 4);
            if (LK_TMA) {
                // TMA: one bulk copy of TR floats per tile column (y is contiguous), all completing on this warp's mbarrier
                fence_proxy_async();
                if (lane == 0) mbar_expect_tx(bar, (unsigned)(TC * TR * sizeof(float)));
                __syncwarp();
                const float* src = fbB + L.oI + (size_t)tx0 * pitch + ty0;
                for (int col = lane; col < TC; col += 32) bulk_g2s(&sT[col][0], col_ptr2(src, pitch4, (unsigned)col), TR * sizeof(float), bar);
            } else {
                const float* src = fbB + L.oI + (size_t)(tx0 + scg) * pitch + (ty0 + 4 * srg);
#pragma unroll
                for (int j = 0; j < (TC + CGN - 1) / CGN; ++j)
                    if (scg < CGN && CGN * j + scg < TC) cp_async16(&sT[CGN * j + scg][4 * srg], col_ptr2(src, (unsigned)CGN * pitch4, j));
                cp_async_commit();
            }
            pending = true;
        }
        while (true) {
            const int nrows = up + down + 1, ncols = left + right + 1;
            if (setup) {
                const int r0 = py - up, c0 = px - left;
                if (nrows < 1 || ncols < 1 || r0 < 1 || c0 < 1 || py + down > H || px + right > W) { ok = false; break; }
                // ---- G from the row prefix planes, lane = window row (as in lk.cu)
                float syy, sxx, syx;
                {
                    const float* colA = fbA + (size_t)(r0 - 1 + min(lane, nrows - 1));
                    const size_t lo = (size_t)(c0 - 1) * pitch, hi = (size_t)(px + right) * pitch;
                    syy = __ldg(colA + L.oRyy + hi) - __ldg(colA + L.oRyy + lo);
                    sxx = __ldg(colA + L.oRxx + hi) - __ldg(colA + L.oRxx + lo);
                    syx = __ldg(colA + L.oRyx + hi) - __ldg(colA + L.oRyx + lo);
                    if (lane >= nrows) { syy = 0.f; sxx = 0.f; syx = 0.f; }
                }
                // ---- template patch of this lane into registers; entries outside the (clipped) window get zero gradients
                {
                    const float* pI = fbA + L.oI + (size_t)(r0 - 1 + pi0) + (size_t)(c0 - 1 + pj0) * pitch;
                    const float2* pG = reinterpret_cast<const float2*>(fbA + L.oG) + (size_t)(r0 - 1 + pi0) + (size_t)(c0 - 1 + pj0) * pitch;
#pragma unroll
                    for (int j = 0; j < PC; ++j) {
                        const float* cI = col_ptr2(pI, pitch4, j);
                        const float2* cG = col_ptr2(pG, 2u * pitch4, j);
#pragma unroll
                        for (int i = 0; i < PR; ++i) {
                            const bool okk = (pi0 + i < nrows) && (pj0 + j < ncols);
                            float iv = 0.f;
                            float2 g2 = make_float2(0.f, 0.f);
                            if (okk) { iv = __ldg(cI + i); g2 = __ldg(cG + i); }
                            TI_REF(i, j) = iv; tG[i][j] = g2;
                        }
                    }
                }
                float fa, fc, fb;
                if (LK_SHARED_SUMS) warp_sum3(syy, sxx, syx, lane, fa, fc, fb);
                else { fa = warp_sum_f2(syy); fc = warp_sum_f2(sxx); fb = warp_sum_f2(syx); }
                const double ga = (double)fa, gc = (double)fc, gb = (double)fb;
                // eigenvalue gate (lucas_kanade.jl:38-46): singular values of the symmetric G = [a b; b c] are Q +- R with
                // Q = |a+c|/2, R = sqrt(((a-c)/2)^2 + b^2) (utils.jl:5-27 with H = 0), so min(S)/area < thr  <=>  |Q - R| < t,
                // t = thr*area  <=>  R < Q + t  and  R > Q - t: decided on squares, no sqrt / division on the common path
                const double E = 0.5 * (ga + gc), F = 0.5 * (ga - gc);
                const double R2 = F * F + gb * gb, Q = fabs(E);
                const double t = a.eig_thr * (double)(nrows * ncols);
                const double qp = Q + t, qm = Q - t;
                if (t > 0.0 && R2 < qp * qp && (qm < 0.0 || R2 > qm * qm)) { ok = false; break; }
                const double tol = 1.4901161193847656e-08;  // sqrt(eps(Float64)), utils.jl:37
                bool full_rank = t > tol;  // the gate already guarantees min(S) >= t
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
                setup = false;
            }
            if (it >= a.iterations) break;
            const double pcy = (double)py + (dy + cy), pcx = (double)px + (dx + cx);
            const int fy = __double2int_rd(pcy), fx = __double2int_rd(pcx);
            // fast path: the keypoint sits >= w+6 px inside the level and the estimate is within 3 px of it, so the estimate
            // lies in the image and get_offsets(point, estimate) is (w, w, w, w) as before: nothing to recompute
            // (the window must still be the unclipped one: an earlier far-off estimate may have clipped it)
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
            if (oy < 0 || oy + RSPAN > TR || ox < 0 || ox + CSPAN > TC) {
                // the estimate walked out of the staged tile (or the window was re-clipped): stage again around it
                if (pending) {
                    if (LK_TMA) { mbar_wait(bar, parity); parity ^= 1; } else cp_async_wait_all();
                }
                __syncwarp();
                ty0 = max(0, ay - 4) & ~3;
                tx0 = max(0, ax - 4);
                oy = ay - ty0; ox = ax - tx0;
                if (LK_TMA) {
                    fence_proxy_async();
                    if (lane == 0) mbar_expect_tx(bar, (unsigned)(TC * TR * sizeof(float)));
                    __syncwarp();
                    const float* src = fbB + L.oI + (size_t)tx0 * pitch + ty0;
                    for (int col = lane; col < TC; col += 32) bulk_g2s(&sT[col][0], col_ptr2(src, pitch4, (unsigned)col), TR * sizeof(float), bar);
                } else {
                    const float* src = fbB + L.oI + (size_t)(tx0 + scg) * pitch + (ty0 + 4 * srg);
#pragma unroll
                    for (int j = 0; j < (TC + CGN - 1) / CGN; ++j)
                        if (scg < CGN && CGN * j + scg < TC) cp_async16(&sT[CGN * j + scg][4 * srg], col_ptr2(src, (unsigned)CGN * pitch4, j));
                    cp_async_commit();
                }
                pending = true;
            }
            if (pending) {
                if (LK_TMA) { mbar_wait(bar, parity); parity ^= 1; } else { cp_async_wait_all(); __syncwarp(); }
                pending = false;
            }
            // ---- prepare_linear_system (lucas_kanade.jl:159-173) on this lane's patch
            const float* tb = &sT[ox + pj0][oy + pi0];
            float by, bx;
#if LK_F32X2
            if constexpr (PC % 2 == 1) {
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
            } else
#endif
            {
            float byr[PR], bxr[PR];  // one accumulator pair per patch row: PR independent FFMA chains instead of one
            float vprev[PR];         // vertical lerps of the previous tap column
#pragma unroll
            for (int i = 0; i < PR; ++i) {
                const float t0 = tb[i], t1 = tb[i + 1];
                vprev[i] = fmaf(wy, t1 - t0, t0);
                byr[i] = 0.f; bxr[i] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < PC; ++j) {
                float tcol[PR + 1];
#pragma unroll
                for (int i = 0; i <= PR; ++i) tcol[i] = tb[(j + 1) * TR + i];
#pragma unroll
                for (int i = 0; i < PR; ++i) {
                    const float vcur = fmaf(wy, tcol[i + 1] - tcol[i], tcol[i]);
                    const float val = fmaf(wx, vcur - vprev[i], vprev[i]);
                    const float dI = TI_REF(i, j) - val;
                    byr[i] = fmaf(dI, tG[i][j].x, byr[i]);
                    bxr[i] = fmaf(dI, tG[i][j].y, bxr[i]);
                    vprev[i] = vcur;
                }
            }
            by = byr[0]; bx = bxr[0];
#pragma unroll
            for (int i = 1; i < PR; ++i) { by += byr[i]; bx += bxr[i]; }
            }
            float fby, fbx;
            if (LK_SHARED_SUMS) warp_sum2(by, bx, lane, fby, fbx);
            else { fby = warp_sum_f2(by); fbx = warp_sum_f2(bx); }
            const double sby = (double)fby, sbx = (double)fbx;
            wpx += (unsigned)(nrows * ncols);
            nit += 1;
            ++it;
            const double ffy = g00 * sby + g01 * sbx, ffx = g01 * sby + g11 * sbx;
            if (fabs(ffy) < eps && fabs(ffx) < eps) break;
            cy += ffy; cx += ffx;
            if (!(fast && fabs(ffy) < 2.0 && fabs(ffx) < 2.0)) {  // on the fast path a step below 2 px cannot leave the image
                const double ny = pcy + ffy, nx = pcx + ffx;
                if (!(__double2int_rd(ny) >= 1 && __double2int_ru(ny) <= H && __double2int_rd(nx) >= 1 && __double2int_ru(nx) <= W)) { ok = false; break; }
            }
        }
        if (!ok) break;
        dy += cy; dx += cx;
        if (lvl > 0) { dy *= 2.0; dx *= 2.0; }
    }

    if (a.mode != 0 && result != 0 && ok) {
        // tracker.jl:59-66: the back-tracked point must land within max_distance of the original keypoint
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
    if (pending) {  // never leave with copies in flight
        if (LK_TMA) mbar_wait(bar, parity); else cp_async_wait_all();
    }
    if (a.mode == 0) {
        if (lane == 0) {
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

// Kernel: one warp per keypoint.  With a.work != nullptr the grid is persistent (one-warp CTAs filling every SM) and every warp
// draws keypoint indices from a device counter -- the next index is requested before the current keypoint is processed, so the
// atomic's latency is hidden; without it CTA i handles keypoints i*LK_WPB .. i*LK_WPB + LK_WPB-1.
template <int W2, int PR, int PC>
#ifdef LK_MAXNREG
__global__ void __maxnreg__(LK_MAXNREG) k_lk_patch(const LKArgs a) {
#else
__global__ void __launch_bounds__(32 * LK_WPB, LK_MINB) k_lk_patch(const LKArgs a) {
#endif
    constexpr int TR = PatchTile<PR, PC>::TR, TC = PatchTile<PR, PC>::TC;
    __shared__ __align__(16) float sTile[LK_WPB][TC][TR];
    __shared__ __align__(8) uint64_t sBar[LK_WPB];
    float (*sT)[TR] = sTile[threadIdx.x >> 5];
    uint64_t* bar = &sBar[threadIdx.x >> 5];
    unsigned parity = 0;
    if (LK_TMA && (threadIdx.x & 31) == 0) mbar_init(bar, 1);
    __syncwarp();
    const int lane = threadIdx.x & 31;
    const int total = a.n_frames * a.n_per_frame;
    if (a.work == nullptr) {
        const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        if (gw < total) lk_point<W2, PR, PC>(a, gw, lane, sT, bar, parity);
        return;
    }
    // a warp draws `chunk` consecutive keypoints per request (SLAMKLT_LK_CHUNK, default 1).  Measured: 1 / 2 / 4 / 8 keypoints per
    // request = 0.955 / 0.997 / 1.114 / 1.241 ms -- neighbours in the caller's order are better handled at the same time by
    // different warps (they meet in L2) than back to back by one warp
    const unsigned chunk = (unsigned)a.chunk;
    int base = 0;
    if (lane == 0) base = (int)atomicAdd(a.work, chunk);
    base = __shfl_sync(FULL, base, 0);
    while (base < total) {
        int next = 0;
        if (lane == 0) next = (int)atomicAdd(a.work, chunk);
        const int end = min(base + (int)chunk, total);
        for (int gw = base; gw < end; ++gw) {
            lk_point<W2, PR, PC>(a, gw, lane, sT, bar, parity);
            __syncwarp();
        }
        base = __shfl_sync(FULL, next, 0);
    }
}

// returns false when this variant does not cover the window size
bool launch_lk_patch(cudaStream_t s, const LKArgs& a) {
    const int total = a.n_frames * a.n_per_frame;
    const int blocks = (total + LK_WPB - 1) / LK_WPB;
    const int w2 = 2 * a.window + 1;
    // experiment knob: SLAMKLT_LK_PAD_KB adds unused dynamic shared memory per CTA to lower the occupancy (4 -> 3 -> 2 CTAs/SM)
    static const int pad = [] {
        const char* e = getenv("SLAMKLT_LK_PAD_KB");
        const int kb = e ? atoi(e) : 0;
        if (kb > 0) cudaFuncSetAttribute(k_lk_patch<19, 3, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, kb * 1024);
        return kb * 1024;
    }();
    if (w2 > 23) return false;
    int grid = blocks;
    static const bool persistent = getenv("SLAMKLT_LK_STATIC") == nullptr;
    LKArgs b = a;
    if (!persistent) b.work = nullptr;
    if (b.work) {
        // persistent grid: every warp slot of every SM, keypoints drawn from the counter (zeroed on this stream first)
        static const int slots = [] {
            int dev = 0, sms = 148;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            // SLAMKLT_LK_SLOTS: resident one-warp CTAs per SM of the persistent grid (default: all 16 the registers allow);
            // fewer leave room for the next batch's pyramid kernels to run next to the tracking kernel
            const char* e = getenv("SLAMKLT_LK_SLOTS");
            const int per_sm = e ? atoi(e) : LK_MINB;
            return sms * (per_sm >= 1 && per_sm <= LK_MINB ? per_sm : LK_MINB);
        }();
        static const int chunk = [] { const char* e = getenv("SLAMKLT_LK_CHUNK"); const int v = e ? atoi(e) : 1; return v < 1 ? 1 : v; }();
        b.chunk = chunk;
        if (blocks > slots * chunk) { grid = slots; cudaMemsetAsync(b.work, 0, sizeof(unsigned), s); }
        else b.work = nullptr;  // fewer keypoints than warp slots: one CTA each
    }
    if (w2 <= 19) k_lk_patch<19, 3, 5><<<grid, 32 * LK_WPB, pad, s>>>(b);
    else k_lk_patch<23, 3, 6><<<grid, 32 * LK_WPB, 0, s>>>(b);
    return true;
}

}  // namespace sk
