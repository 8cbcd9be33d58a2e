// Pyramid construction kernels (sm_100a): input conversion, Young-van Vliet recursive Gaussian with
// Triggs-Sdika boundaries along both dimensions, Scharr gradients + gradient products, bilinear decimation.
//
// Reference behaviour: pyramid.jl:40-137 (LKPyramid ctor, update!, gaussian_pyramid!, imgradients_yx!),
// lucas_kanade.jl:102-129 (compute_partial_derivatives!).  The 2-D integral images of lucas_kanade.jl:131-138 are
// replaced by 1-D exclusive prefix sums along x of the smoothed planes (accumulated in Float64, stored fp32): the
// LK kernel gets a window sum from two coalesced loads per row and plane, and fp32 keeps enough digits because a
// row prefix is ~65x a window sum, not ~10^5x like a 2-D integral image.
//
// Layout: every plane is fp32, y contiguous (Julia column-major), pitch = roundup4(H+1) floats, W+1 columns; the guard
// row(s) y >= H and the guard column x = W are kept at zero by every kernel here.  Iy and Ix are interleaved
// (float2 per pixel) so that the LK template costs one 8-byte load per pixel.
//
// The recursion u[i] = x[i] + a1 u[i-1] + a2 u[i-2] + a3 u[i-3] is sequential along a line.  Both kernels
// cut each line into chunks, run every chunk from a zero state to obtain its outgoing state, combine the
// chunk states with powers of the companion matrix A (exact by linearity), and re-run every chunk from its
// true incoming state.  Along y (contiguous, short lines) one warp owns a whole column in registers (vector
// loads, lane l owns rows [l*K, l*K+K)) and the carries travel through a Kogge-Stone scan on shuffles; along x
// (strided, long lines) a CTA owns a strip of rows, chunks are spread over warps and the carries travel through
// shared memory.
#include <cstdio>

#include "common.cuh"
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

// experiment knob: software prefetch of the raw Float64 column COLS_PREFETCH_DIST columns ahead (1 = into L2, 2 = into L1)
#ifndef COLS_PREFETCH
#define COLS_PREFETCH 0
#endif
// raw Float64 columns through a cp.async shared-memory stage, one column ahead (0 = direct __ldg loads).  Measured on B200:
// 0.280 ms vs 0.269 ms direct -- long-scoreboard stalls drop (3.1 -> 1.0 cycles/instr) but the extra LDS/LDGSTS traffic moves
// them to the shared-memory pipe (short scoreboard 0.8 -> 1.9, MIO throttle 0.2 -> 0.9), which the scan's shuffles already load.
#ifndef COLS_STAGE
#define COLS_STAGE 0
#endif
#ifndef COLS_PREFETCH_DIST
#define COLS_PREFETCH_DIST 2
#endif
// UInt8 / Float32 / fp32-plane source columns requested one column iteration ahead into registers (K / 4 words for bytes, K for
// floats): the loads of column x + 2 are in flight while column x is filtered
#ifndef COLS_REGPF
#define COLS_REGPF 1
#endif

namespace sk {

// ----------------------------------------------------------------------------------------------
// input conversion: host layout (dtype, ld) -> level-0 layer (fp32, pitch) [+ optional f64 copy for detect]
// ----------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ double to_unit(T v);
template <> __device__ __forceinline__ double to_unit<double>(double v) { return v; }
template <> __device__ __forceinline__ double to_unit<float>(float v) { return (double)v; }
template <> __device__ __forceinline__ double to_unit<uint8_t>(uint8_t v) { return (double)v / 255.0; }

template <typename T>
__global__ void k_convert(const T* __restrict__ src, int ld, size_t src_stride, FrameSet dst, int dst_f0, size_t off_I, int H, int W,
                          int pitch, double* __restrict__ dst64) {
    const int f = blockIdx.z;
    const int x = blockIdx.y;
    const T* s = src + (size_t)f * src_stride + (size_t)x * ld;
    float* d = dst.frame(dst_f0 + f) + off_I + (size_t)x * pitch;
    double* d64 = dst64 ? dst64 + ((size_t)f * W + x) * H : nullptr;
    for (int y = blockIdx.x * blockDim.x + threadIdx.x; y < H; y += gridDim.x * blockDim.x) {
        double v = to_unit<T>(s[y]);
        d[y] = (float)v;
        if (d64) d64[y] = v;
    }
}

int launch_convert(cudaStream_t s, const void* src, int dtype, int ld, size_t src_stride, FrameSet dst, int dst_f0, int n_frames,
                   const PyrGeom& g, double* dst64, const Hook* hk) {
    const LevelGeom& l0 = g.lv[0];
    mark(hk, "k_convert");
    dim3 grid((l0.H + 127) / 128, l0.W, n_frames), block(128);
    size_t off = plane_off(l0, DP_I);
    if (dtype == SLAMKLT_F64)
        k_convert<double><<<grid, block, 0, s>>>((const double*)src, ld, src_stride, dst, dst_f0, off, l0.H, l0.W, l0.pitch, dst64);
    else if (dtype == SLAMKLT_F32)
        k_convert<float><<<grid, block, 0, s>>>((const float*)src, ld, src_stride, dst, dst_f0, off, l0.H, l0.W, l0.pitch, dst64);
    else
        k_convert<uint8_t><<<grid, block, 0, s>>>((const uint8_t*)src, ld, src_stride, dst, dst_f0, off, l0.H, l0.W, l0.pitch, dst64);
    return 1;
}

// ----------------------------------------------------------------------------------------------
// column I/O: lane l owns rows [l*K, l*K+K) of a column; K is even, so 8- or 16-byte vectors are always aligned
// ----------------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ void load_col(const float* __restrict__ col, int y0, int pitch, float (&x)[K]) {
    if constexpr (K % 4 == 0) {
#pragma unroll
        for (int v = 0; v < K / 4; ++v) {
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (y0 + 4 * v < pitch) t = __ldg(reinterpret_cast<const float4*>(col + y0 + 4 * v));
            x[4 * v] = t.x; x[4 * v + 1] = t.y; x[4 * v + 2] = t.z; x[4 * v + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int v = 0; v < K / 2; ++v) {
            float2 t = make_float2(0.f, 0.f);
            if (y0 + 2 * v < pitch) t = __ldg(reinterpret_cast<const float2*>(col + y0 + 2 * v));
            x[2 * v] = t.x; x[2 * v + 1] = t.y;
        }
    }
}

// rows >= H (guard rows) are written as zero
template <int K>
__device__ __forceinline__ void store_col(float* __restrict__ col, int y0, int pitch, int H, const float (&x)[K]) {
    if constexpr (K % 4 == 0) {
#pragma unroll
        for (int v = 0; v < K / 4; ++v) {
            const int y = y0 + 4 * v;
            if (y < pitch) {
                float4 t;
                t.x = y < H ? x[4 * v] : 0.f; t.y = y + 1 < H ? x[4 * v + 1] : 0.f;
                t.z = y + 2 < H ? x[4 * v + 2] : 0.f; t.w = y + 3 < H ? x[4 * v + 3] : 0.f;
                *reinterpret_cast<float4*>(col + y) = t;
            }
        }
    } else {
#pragma unroll
        for (int v = 0; v < K / 2; ++v) {
            const int y = y0 + 2 * v;
            if (y < pitch) {
                float2 t;
                t.x = y < H ? x[2 * v] : 0.f; t.y = y + 1 < H ? x[2 * v + 1] : 0.f;
                *reinterpret_cast<float2*>(col + y) = t;
            }
        }
    }
}

// Row validity at group granularity.  When H is a multiple of G (and K is), a lane's K rows fall into K/G groups that are
// valid or invalid as a whole, so the K per-row tests `y0 + j < H` collapse into K/G loop-invariant predicates (G = 1 is the
// general case).  Guard rows (>= H) are zero from the allocation's memset and no kernel ever writes anything else there,
// so stores of invalid groups are simply skipped.
template <int K, int G>
struct RowGroups {
    bool g[K / G];
    __device__ __forceinline__ RowGroups(int y0, int H) {
#pragma unroll
        for (int i = 0; i < K / G; ++i) g[i] = y0 + i * G < H;
    }
    __device__ __forceinline__ bool valid(int j) const { return g[j / G]; }
};

template <int K, int G>
__device__ __forceinline__ void store_col_g(float* __restrict__ col, int y0, const RowGroups<K, G>& rg, const float (&x)[K]) {
    constexpr int V = (G % 4 == 0) ? 4 : 2;
    static_assert(G % 2 == 0 && K % V == 0, "vector stores need even groups");
#pragma unroll
    for (int v = 0; v < K / V; ++v) {
        if (rg.valid(V * v)) {
            if constexpr (V == 4) *reinterpret_cast<float4*>(col + y0 + 4 * v) = make_float4(x[4 * v], x[4 * v + 1], x[4 * v + 2], x[4 * v + 3]);
            else *reinterpret_cast<float2*>(col + y0 + 2 * v) = make_float2(x[2 * v], x[2 * v + 1]);
        }
    }
}

// ----------------------------------------------------------------------------------------------
// warp-wide recursive filter of one line held in registers: lane l owns elements [l*K, l*K+K)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mat3_acc(const float* __restrict__ P, float r0, float r1, float r2, float& q0, float& q1, float& q2) {
    q0 = fmaf(P[0], r0, fmaf(P[1], r1, fmaf(P[2], r2, q0)));
    q1 = fmaf(P[3], r0, fmaf(P[4], r1, fmaf(P[5], r2, q1)));
    q2 = fmaf(P[6], r0, fmaf(P[7], r1, fmaf(P[8], r2, q2)));
}

// NL independent lines are filtered together so that their dependent FMA / shuffle chains interleave (ILP = NL).
// The line is extended to the full 32*K elements with its border value (x[n-1] for replicate, 0 for Fill(0)): the
// Triggs-Sdika boundary is exactly the constant-extension assumption, so the result on [0, n) is unchanged while
// every lane becomes a full chunk and no per-element predicate is needed.
// HV = 2: the line is shared by a pair of warps (half 0: elements [0, 32 K), half 1: [32 K, 64 K); the launcher guarantees
// n > 32 K, so the line ends in half 1).  Each half scans its own chunks; after the forward scan half 0 hands the true state at
// its end to half 1, whose lane l adds A^(K (l + 1)) times that state (table PL, one 3 x 3 matrix per lane); after the backward
// scan half 1 hands the state at its start to half 0, whose lane l adds A^(K (32 - l)) times it.  Two 64-thread named barriers
// per call; the exchange buffers xch (18 floats per pair) are separate for the two directions, so calls can follow each other.
__device__ __forceinline__ void pair_sync(int bar) { asm volatile("bar.sync %0, 64;" ::"r"(bar) : "memory"); }

template <int K, int NL, int G = 1, int HV = 1>
__device__ __forceinline__ void warp_iir_lines(float (&x)[NL][K], const int n, const int lane, const IirDev& c, const bool zero_border,
                                               const int half = 0, float* xch = nullptr, const float* __restrict__ PL = nullptr, const int bar = 0) {
    const bool first_half = HV == 1 || half == 0, last_half = HV == 1 || half == 1;
    const int y0 = (HV == 2 ? half * 32 + lane : lane) * K;
    const int jl = n - 1 - y0;  // slot of the last element of the line, if it lives in this lane
    const float a1 = c.a1, a2 = c.a2, a3 = c.a3;
    const int ln = ((n - 1) / K) & 31;
    float um[NL], iplus[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        float first = __shfl_sync(FULL, x[l][0], 0);
        float lastv = 0.f;
#pragma unroll
        for (int j = G - 1; j < K; j += G)  // n is a multiple of G: the last element closes a group
            if (j == jl) lastv = x[l][j];
        lastv = __shfl_sync(FULL, lastv, ln);
        um[l] = (zero_border ? 0.f : first) * c.inv1ma;
        iplus[l] = zero_border ? 0.f : lastv;
        if constexpr (G == 1) {
#pragma unroll
            for (int j = 0; j < K; ++j)
                if (j > jl) x[l][j] = iplus[l];  // constant extension past the end of the line
        } else {
#pragma unroll
            for (int i = 0; i < K / G; ++i) {
                const bool past = i * G > jl;
#pragma unroll
                for (int j = i * G; j < i * G + G; ++j) x[l][j] = past ? iplus[l] : x[l][j];
            }
        }
    }
    // forward, phase 1: chunk-local pass (lane 0 starts from the true left boundary state)
    float s0[NL], s1[NL], s2[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) { s0[l] = (lane == 0 && first_half) ? um[l] : 0.f; s1[l] = s0[l]; s2[l] = s0[l]; }
#pragma unroll
    for (int j = 0; j < K; ++j) {
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            float u = fmaf(a1, s0[l], fmaf(a2, s1[l], fmaf(a3, s2[l], x[l][j])));  // newest state last: 1 FMA on the recurrence's critical path
            s2[l] = s1[l]; s1[l] = s0[l]; s0[l] = u;
        }
    }
    // phase 2: inclusive scan of the chunk states, q_l += A^(K d) q_{l-d}
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        if (j >= c.nsc) break;
        const int d = 1 << j;
        float r0[NL], r1[NL], r2[NL];
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            r0[l] = __shfl_up_sync(FULL, s0[l], d); r1[l] = __shfl_up_sync(FULL, s1[l], d); r2[l] = __shfl_up_sync(FULL, s2[l], d);
        }
        if (lane >= d) {
#pragma unroll
            for (int l = 0; l < NL; ++l) mat3_acc(c.P[j], r0[l], r1[l], r2[l], s0[l], s1[l], s2[l]);
        }
    }
    float sin_[NL][3];  // HV = 2, half 1: the state half 0 ends with
    if constexpr (HV == 2) {
        if (half == 0 && lane == 31) {
#pragma unroll
            for (int l = 0; l < NL; ++l) { xch[3 * l] = s0[l]; xch[3 * l + 1] = s1[l]; xch[3 * l + 2] = s2[l]; }
        }
        pair_sync(bar);
        if (half == 1) {
            float Pm[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) Pm[i] = __ldg(PL + lane * 9 + i);
#pragma unroll
            for (int l = 0; l < NL; ++l) {
                sin_[l][0] = xch[3 * l]; sin_[l][1] = xch[3 * l + 1]; sin_[l][2] = xch[3 * l + 2];
                mat3_acc(Pm, sin_[l][0], sin_[l][1], sin_[l][2], s0[l], s1[l], s2[l]);
            }
        }
    }
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        float i0 = __shfl_up_sync(FULL, s0[l], 1), i1 = __shfl_up_sync(FULL, s1[l], 1), i2 = __shfl_up_sync(FULL, s2[l], 1);
        if (lane == 0) {
            if (first_half) { i0 = um[l]; i1 = um[l]; i2 = um[l]; }
            else if constexpr (HV == 2) { i0 = sin_[l][0]; i1 = sin_[l][1]; i2 = sin_[l][2]; }
        }
        s0[l] = i0; s1[l] = i1; s2[l] = i2;
    }
    // phase 3: true pass
#pragma unroll
    for (int j = 0; j < K; ++j) {
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            float u = fmaf(a1, s0[l], fmaf(a2, s1[l], fmaf(a3, s2[l], x[l][j])));  // newest state last: 1 FMA on the recurrence's critical path
            x[l][j] = u;
            s2[l] = s1[l]; s1[l] = s0[l]; s0[l] = u;
        }
    }
    // right boundary (Triggs & Sdika eq. 14) at the end of the extended line: lane 31 holds (u[N], u[N-1], u[N-2])
    float vr0[NL], vr1[NL], vr2[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        const float e0 = __shfl_sync(FULL, s0[l], 31), e1 = __shfl_sync(FULL, s1[l], 31), e2 = __shfl_sync(FULL, s2[l], 31);
        const float up = iplus[l] * c.inv1ma, vp = up * c.inv1ma;
        const float d0 = e0 - up, d1 = e1 - up, d2 = e2 - up;
        vr0[l] = fmaf(c.M[0], d0, fmaf(c.M[1], d1, fmaf(c.M[2], d2, vp)));
        vr1[l] = fmaf(c.M[3], d0, fmaf(c.M[4], d1, fmaf(c.M[5], d2, vp)));
        vr2[l] = fmaf(c.M[6], d0, fmaf(c.M[7], d1, fmaf(c.M[8], d2, vp)));
    }
    // backward, phase 1 (the very last element takes the boundary value instead of the recursion)
    float t0[NL], t1[NL], t2[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) { t0[l] = 0.f; t1[l] = 0.f; t2[l] = 0.f; }
#pragma unroll
    for (int j = K - 1; j >= 0; --j) {
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            float v = fmaf(a1, t0[l], fmaf(a2, t1[l], fmaf(a3, t2[l], x[l][j])));
            t2[l] = t1[l]; t1[l] = t0[l]; t0[l] = v;
            if (j == K - 1 && lane == 31 && last_half) { t0[l] = vr0[l]; t1[l] = vr1[l]; t2[l] = vr2[l]; }
        }
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        if (j >= c.nsc) break;
        const int d = 1 << j;
        float r0[NL], r1[NL], r2[NL];
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            r0[l] = __shfl_down_sync(FULL, t0[l], d); r1[l] = __shfl_down_sync(FULL, t1[l], d); r2[l] = __shfl_down_sync(FULL, t2[l], d);
        }
        if (lane + d < 32) {
#pragma unroll
            for (int l = 0; l < NL; ++l) mat3_acc(c.P[j], r0[l], r1[l], r2[l], t0[l], t1[l], t2[l]);
        }
    }
    float tin_[NL][3];  // HV = 2, half 0: the state half 1 starts with
    if constexpr (HV == 2) {
        if (half == 1 && lane == 0) {
#pragma unroll
            for (int l = 0; l < NL; ++l) { xch[9 + 3 * l] = t0[l]; xch[9 + 3 * l + 1] = t1[l]; xch[9 + 3 * l + 2] = t2[l]; }
        }
        pair_sync(bar);
        if (half == 0) {
            float Pm[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) Pm[i] = __ldg(PL + (31 - lane) * 9 + i);
#pragma unroll
            for (int l = 0; l < NL; ++l) {
                tin_[l][0] = xch[9 + 3 * l]; tin_[l][1] = xch[9 + 3 * l + 1]; tin_[l][2] = xch[9 + 3 * l + 2];
                mat3_acc(Pm, tin_[l][0], tin_[l][1], tin_[l][2], t0[l], t1[l], t2[l]);
            }
        }
    }
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        float i0 = __shfl_down_sync(FULL, t0[l], 1), i1 = __shfl_down_sync(FULL, t1[l], 1), i2 = __shfl_down_sync(FULL, t2[l], 1);
        if (lane == 31) {
            if (last_half) { i0 = 0.f; i1 = 0.f; i2 = 0.f; }
            else if constexpr (HV == 2) { i0 = tin_[l][0]; i1 = tin_[l][1]; i2 = tin_[l][2]; }
        }
        t0[l] = i0; t1[l] = i1; t2[l] = i2;
    }
    const float sc = c.scale;
#pragma unroll
    for (int j = K - 1; j >= 0; --j) {
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            float v = fmaf(a1, t0[l], fmaf(a2, t1[l], fmaf(a3, t2[l], x[l][j])));
            if (j == K - 1 && lane == 31 && last_half) { v = vr0[l]; t0[l] = vr0[l]; t1[l] = vr1[l]; t2[l] = vr2[l]; }
            else { t2[l] = t1[l]; t1[l] = t0[l]; t0[l] = v; }
            x[l][j] = v * sc;
        }
    }
}

// ----------------------------------------------------------------------------------------------
// dim-1 kernels (along y)
// ----------------------------------------------------------------------------------------------
struct ColArgs {
    FrameSet fs;
    int f0, n_frames;
    int H, W, pitch;
    int zero_border;       // blur: NA mode; grad: Fill(0) Scharr border
    int strip;             // k_cols_all: columns per warp strip
    size_t o_in;           // layer
    size_t o_out0;         // blur: T0;   grad: T0,T1,T2 consecutive
    size_t o_grad;         // grad only: interleaved (Iy, Ix)
    size_t plane_elems;
    const float* inv_n;    // blur NA mode: 1/ny[y]
    // fused kernel only: level 0 reads the staged host image directly (and writes the layer), levels < L also run the
    // y pass of the pyramid blur
    const void* raw;       // nullptr: read the layer plane
    int raw_ld;
    size_t raw_stride;     // elements per frame
    int do_blur;
    size_t o_tmp;          // y-filtered blur output
    // T planes in a scratch ring (TScratch) instead of at o_out0 inside the frame
    float* t_base;
    size_t t_stride, o_t;
    int t_ring;
    // two warps per column (k_cols_all<..., HV = 2>): per-lane matrices A^(K (l + 1)) of the sigma = 4 and the blur recursion
    const float* pl4;
    const float* pl1;
};

// One warp per (frame, column).
template <int K, int G>
__global__ void __launch_bounds__(256) k_cols_blur(ColArgs a, IirDev c) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int total = a.n_frames * a.W;
    const int y0 = lane * K;
    const RowGroups<K, G> rg(y0, a.H);
    for (int w = warp; w < total; w += nwarps) {
        const int f = w / a.W, xcol = w - f * a.W;
        float* fb = a.fs.frame(a.f0 + f);
        const float* in = fb + a.o_in + (size_t)xcol * a.pitch;
        float* out = fb + a.o_out0 + (size_t)xcol * a.pitch;
        float x[1][K];
        load_col<K>(in, y0, a.pitch, x[0]);
        warp_iir_lines<K, 1, G>(x, a.H, lane, c, a.zero_border != 0);
        if (a.inv_n) {
#pragma unroll
            for (int j = 0; j < K; ++j)
                if (rg.valid(j)) x[0][j] *= __ldg(a.inv_n + y0 + j);
        }
        if constexpr (G > 1) store_col_g<K, G>(out, y0, rg, x[0]);
        else store_col<K>(out, y0, a.pitch, a.H, x[0]);
    }
}

// Scharr gradients (pyramid.jl:59,75,98-103), gradient products and the y pass of their sigma=4 smoothing
// (lucas_kanade.jl:116-126).  A warp walks a strip of CS consecutive columns with a 3-column sliding window in
// registers, so every layer column is loaded once per strip (+2 halo columns).  Writes the interleaved (Iy, Ix)
// plane and the three y-filtered product planes T0 (yy), T1 (xx), T2 (yx).
constexpr int GRAD_CS = 8;

// (float)((double)k / 255.0) for k = 0 .. 255 -- the layer value of an 8-bit pixel (Gray{Float64}(k / 255) converted to fp32) --
// without the Float64 division: q = k * (1/255) corrected by one Newton step on the exact remainder equals the correctly rounded
// fp32 quotient, and that equals the doubly rounded reference value for every one of the 256 inputs (checked exhaustively:
// tests/test_oracle.py::test_u8_unit_conversion_is_exact).
__device__ __forceinline__ float u8_unit(unsigned k) {
    const float kf = (float)k, r = 1.0f / 255.0f;
    const float q = kf * r;
    return fmaf(fmaf(-q, 255.0f, kf), r, q);
}

// SRC: 0 layer plane (fp32), 1 raw Float64, 2 raw Float32, 3 raw UInt8 (value / 255)
template <int K, int SRC, int G = 1>
__device__ __forceinline__ void load_col_any(const ColArgs& a, const float* __restrict__ I, int f, int xc, int y0, float (&x)[K],
                                             const RowGroups<K, G>& rg) {
    if constexpr (SRC == 0) {
        load_col<K>(I + (size_t)xc * a.pitch, y0, a.pitch, x);
    } else if constexpr (SRC == 1 && G > 1) {
        // aligned variant (the launcher guarantees raw_ld even): whole groups are valid or not
        const double* col = reinterpret_cast<const double*>(a.raw) + (size_t)f * a.raw_stride + (size_t)xc * a.raw_ld;
#pragma unroll
        for (int v = 0; v < K / 2; ++v) {
            double2 t = make_double2(0.0, 0.0);
            if (rg.valid(2 * v)) t = __ldg(reinterpret_cast<const double2*>(col + y0 + 2 * v));
            x[2 * v] = (float)t.x; x[2 * v + 1] = (float)t.y;
        }
    } else if constexpr (SRC == 1) {
        const double* col = reinterpret_cast<const double*>(a.raw) + (size_t)f * a.raw_stride + (size_t)xc * a.raw_ld;
        if ((a.raw_ld & 1) == 0) {  // 16-byte aligned column starts: K is even, so pairs never straddle the end of a valid run
#pragma unroll
            for (int v = 0; v < K / 2; ++v) {
                const int y = y0 + 2 * v;
                double2 t = make_double2(0.0, 0.0);
                if (y + 1 < a.H) t = __ldg(reinterpret_cast<const double2*>(col + y));
                else if (y < a.H) t.x = __ldg(col + y);
                x[2 * v] = (float)t.x; x[2 * v + 1] = (float)t.y;
            }
        } else {
#pragma unroll
            for (int j = 0; j < K; ++j) x[j] = (y0 + j < a.H) ? (float)__ldg(col + y0 + j) : 0.f;
        }
    } else if constexpr (SRC == 2) {
        const float* col = reinterpret_cast<const float*>(a.raw) + (size_t)f * a.raw_stride + (size_t)xc * a.raw_ld;
#pragma unroll
        for (int j = 0; j < K; ++j) x[j] = rg.valid(j) ? __ldg(col + y0 + j) : 0.f;
    } else {
        const uint8_t* col = reinterpret_cast<const uint8_t*>(a.raw) + (size_t)f * a.raw_stride + (size_t)xc * a.raw_ld;
        if constexpr (G >= 4 && K % 4 == 0) {
            // aligned variant (H, and with it every column start, is a multiple of 4): four rows per 32-bit load
#pragma unroll
            for (int v = 0; v < K / 4; ++v) {
                const unsigned w = rg.valid(4 * v) ? __ldg(reinterpret_cast<const unsigned*>(col + y0 + 4 * v)) : 0u;
                x[4 * v] = u8_unit(w & 0xffu); x[4 * v + 1] = u8_unit((w >> 8) & 0xffu);
                x[4 * v + 2] = u8_unit((w >> 16) & 0xffu); x[4 * v + 3] = u8_unit(w >> 24);
            }
        } else if constexpr (G >= 2) {
#pragma unroll
            for (int v = 0; v < K / 2; ++v) {
                const unsigned w = rg.valid(2 * v) ? (unsigned)__ldg(reinterpret_cast<const unsigned short*>(col + y0 + 2 * v)) : 0u;
                x[2 * v] = u8_unit(w & 0xffu); x[2 * v + 1] = u8_unit(w >> 8);
            }
        } else {
#pragma unroll
            for (int j = 0; j < K; ++j) x[j] = rg.valid(j) ? u8_unit(__ldg(col + y0 + j)) : 0.f;
        }
    }
}

// one source element (row y of column xc), converted exactly like load_col_any converts it
template <int SRC>
__device__ __forceinline__ float load_px_any(const ColArgs& a, const float* __restrict__ I, int f, int xc, int y) {
    if constexpr (SRC == 0) return __ldg(I + (size_t)xc * a.pitch + y);
    else if constexpr (SRC == 1) return (float)__ldg(reinterpret_cast<const double*>(a.raw) + (size_t)f * a.raw_stride + (size_t)xc * a.raw_ld + y);
    else if constexpr (SRC == 2) return __ldg(reinterpret_cast<const float*>(a.raw) + (size_t)f * a.raw_stride + (size_t)xc * a.raw_ld + y);
    else return u8_unit(__ldg(reinterpret_cast<const uint8_t*>(a.raw) + (size_t)f * a.raw_stride + (size_t)xc * a.raw_ld + y));
}

// e[0] = row y0-1, e[1..K] = rows y0..y0+K-1, e[K+1] = row y0+K from the lane's own rows x[] (border rule along y applied here).
// HV = 2: `cross` is the row on the other side of the boundary between the two halves (row 32 K for half 0, row 32 K - 1 for half 1)
template <int K, int G, int HV = 1>
__device__ __forceinline__ void halo_from_col(const float (&x)[K], int y0, int H, int lane, bool zb, float (&e)[K + 2], int half = 0, float cross = 0.f) {
    float upv = __shfl_up_sync(FULL, x[K - 1], 1);
    float dnv = __shfl_down_sync(FULL, x[0], 1);
    if (lane == 0) upv = (HV == 2 && half == 1) ? cross : (zb ? 0.f : x[0]);
    if (lane == 31) dnv = (HV == 2 && half == 0) ? cross : 0.f;
    e[0] = upv;
#pragma unroll
    for (int j = 0; j < K; ++j) e[j + 1] = x[j];
    e[K + 1] = dnv;
    if (!zb) {
#pragma unroll
        for (int j = 1; j <= K + 1; j += G)  // row H opens a group
            if (y0 + j - 1 == H) e[j] = e[j - 1];
    }
}

template <int K, int SRC, int G = 1, int HV = 1>
__device__ __forceinline__ void load_col_halo_any(const ColArgs& a, const float* __restrict__ I, int f, int xcol, int y0, int lane, bool zb,
                                                  float (&e)[K + 2], const RowGroups<K, G>& rg, int half = 0) {
    const int W = a.W, H = a.H;
    float x[K];
    float cross = 0.f;
    const bool inside = xcol >= 0 && xcol < W;
    if (inside || !zb) {
        const int xc = xcol < 0 ? 0 : (xcol >= W ? W - 1 : xcol);
        load_col_any<K, SRC, G>(a, I, f, xc, y0, x, rg);
        if constexpr (HV == 2) cross = load_px_any<SRC>(a, I, f, xc, half == 0 ? 32 * K : 32 * K - 1);  // (H > 32 K: both rows exist)
    } else {
#pragma unroll
        for (int j = 0; j < K; ++j) x[j] = 0.f;
    }
    halo_from_col<K, G, HV>(x, y0, H, lane, zb, e, half, cross);
}

// Raw Float64 columns staged through shared memory with cp.async, one column ahead of the column loop: a lane copies its
// own K rows (K/2 16-byte pieces) and later reads back only what it copied itself, so no barrier is needed, just
// cp.async.wait_group.  The lane stride is an odd number of 16-byte units => conflict-free 128-bit shared loads.
template <int K>
struct RawStage {
    static constexpr int U = K / 2;                       // 16-byte units per lane
    static constexpr int LU = (U % 2 == 0) ? U + 1 : U;   // lane stride in units (odd)
    static constexpr int STAGE_UNITS = 32 * LU;
};
__device__ __forceinline__ void cp_async_cg16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all(bool wait) {
    if (wait) asm volatile("cp.async.wait_group 0;" ::: "memory");
    else asm volatile("cp.async.commit_group;" ::: "memory");
}

// Fused column kernel: [level 0: input conversion + layer store] + Scharr + products + y pass of the sigma=4 filter
// + [levels < L: y pass of the pyramid blur].  One read of the source column per output column (+2 halo columns per strip).
// experiment knob: minimum resident 4-warp CTAs per SM the register allocation of k_cols_all is bounded for (0 = unbounded)
#ifndef COLS_MINB
#define COLS_MINB 4
#endif
// 16 rows per lane (frames of 385 .. 512 rows, e.g. 752 x 480): 203 registers unbounded = 2 CTAs per SM.  Measured on B200, build of
// 64 frames of 480 x 752: unbounded 0.633 ms, bounded for 3 CTAs (168 registers, no spills) 0.577 ms, for 4 CTAs (128 registers, 120
// bytes of spills) 0.623 ms; two warps per column with 8 rows per lane 0.611 ms (128 registers) / 0.688 ms (96 registers).
#ifndef COLS_MINB16
#define COLS_MINB16 3
#endif
#ifndef COLS_MINB_TALL
#define COLS_MINB_TALL 0
#endif
// two warps per column (HV = 2): resident 4-warp CTAs per SM the register allocation is bounded for
#ifndef COLS_PAIR_MINB8
#define COLS_PAIR_MINB8 5
#endif
#ifndef COLS_PAIR_MINB12
#define COLS_PAIR_MINB12 4
#endif
#ifndef COLS_PAIR_MINB18
#define COLS_PAIR_MINB18 3
#endif
// rows per lane of the pair kernel for frames taller than 768 rows (18: 94 % of the lanes' rows used at 1080 rows, 8-byte vectors;
// 20: 84 %, 16-byte vectors).  Measured on B200, 16 frames of 1080 x 1920 (c5), level-0 column kernel: one warp per column (34 rows
// per lane, 255 registers) 0.699 ms; pairs with 18 rows at 168 registers 0.485 ms, at 128 registers (spills) 0.540 ms; pairs with 20
// rows at 168 registers 0.362 ms, at 128 registers 0.443 ms.
// COLS_PAIR_SMALL=1 (experiment) also sends frames of 193 .. 512 rows through pairs (6 or 8 rows per lane).
#ifndef COLS_PAIR_SMALL
#define COLS_PAIR_SMALL 0
#endif
#ifndef COLS_PAIR_KBIG
#define COLS_PAIR_KBIG 20
#endif
// HV = 2 (frames taller than 512 rows): a pair of warps shares a strip, warp `half` owns rows [32 K half, 32 K (half + 1)) of every
// column -- K = 12 / 18 rows per lane instead of 24 / 34, half the registers, twice the resident warps.
template <int K, int SRC, int G, int HV = 1>
__global__ void __launch_bounds__(128, HV == 2 ? (K <= 8 ? COLS_PAIR_MINB8 : (K <= 12 ? COLS_PAIR_MINB12 : COLS_PAIR_MINB18))
                                              : ((K <= 12 && COLS_MINB > 0) ? COLS_MINB : ((K == 16 && COLS_MINB16 > 0) ? COLS_MINB16 : ((K >= 24 && COLS_MINB_TALL > 0) ? COLS_MINB_TALL : 1))))
    k_cols_all(ColArgs a, IirDev c4, IirDev c1) {
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int warp = HV == 2 ? gwarp >> 1 : gwarp;                                  // strip walker: a warp, or a pair of warps
    const int nwarps = ((gridDim.x * blockDim.x) >> 5) >> (HV == 2 ? 1 : 0);
    const int half = HV == 2 ? (gwarp & 1) : 0;
    const int bar = 1 + ((threadIdx.x >> 6) & 1);                                     // named barrier of this pair (two pairs per CTA)
    __shared__ float sXch[HV == 2 ? 2 * 2 * 18 : 1];
    float* const xch4 = sXch + (HV == 2 ? ((threadIdx.x >> 6) & 1) * 36 : 0);      // sigma = 4 filter (3 lines)
    float* const xch1 = xch4 + (HV == 2 ? 18 : 0);                                    // pyramid blur (1 line)
    const int H = a.H, W = a.W, pitch = a.pitch;
    const int cs = a.strip;  // columns per warp strip (GRAD_CS for large batches, narrower when there are few frames)
    const int strips = (W + cs - 1) / cs;
    const int total = a.n_frames * strips;
    const int y0 = (half * 32 + lane) * K;
    const bool zb = a.zero_border != 0;
    const RowGroups<K, G> rg(y0, H);
    constexpr bool STAGED = COLS_STAGE && SRC == 1 && G > 1 && K <= 16 && HV == 1;
    __shared__ __align__(16) double2 sRaw[STAGED ? 4 * 2 * RawStage<K>::STAGE_UNITS : 1];
    double2* const sMine = sRaw + (threadIdx.x >> 5) * 2 * RawStage<K>::STAGE_UNITS + lane * RawStage<K>::LU;
    for (int w = warp; w < total; w += nwarps) {
        const int f = w / strips, xb = (w - f * strips) * cs;
        float* fb = a.fs.frame(a.f0 + f);
        const float* I = fb + a.o_in;
        float em[K + 2], ec[K + 2], ep[K + 2];
        const int xe = min(xb + cs, W);
        // column c (clamped into the image) -> stage c & 1; issued one iteration before it is consumed
        auto stage_issue = [&](int c) {
            if constexpr (STAGED) {
                if (c >= W) { if (zb) return; c = W - 1; }
                const double* col = reinterpret_cast<const double*>(a.raw) + (size_t)f * a.raw_stride + (size_t)c * a.raw_ld + y0;
                double2* dst = sMine + (c & 1) * RawStage<K>::STAGE_UNITS;
#pragma unroll
                for (int v = 0; v < K / 2; ++v)
                    if (rg.valid(2 * v)) cp_async_cg16(dst + v, col + 2 * v);
                cp_async_commit_wait_all(false);
            }
        };
        // register prefetch (UInt8 source, aligned variant): raw words of the column one iteration ahead
        constexpr bool PF = COLS_REGPF && SRC == 3 && G >= 4 && K % 4 == 0 && HV == 1;
        unsigned nxt[PF ? K / 4 : 1];
        bool nxt_zero = false;
        auto pf_issue = [&](int c) {
            if constexpr (PF) {
                nxt_zero = zb && c >= W;
                const int cc = c >= W ? W - 1 : c;
                const uint8_t* col = reinterpret_cast<const uint8_t*>(a.raw) + (size_t)f * a.raw_stride + (size_t)cc * a.raw_ld + y0;
#pragma unroll
                for (int v = 0; v < K / 4; ++v) nxt[v] = rg.valid(4 * v) ? __ldg(reinterpret_cast<const unsigned*>(col + 4 * v)) : 0u;
            }
        };
        pf_issue(xb + 1);
        stage_issue(xb + 1);
        load_col_halo_any<K, SRC, G, HV>(a, I, f, xb - 1, y0, lane, zb, em, rg, half);
        load_col_halo_any<K, SRC, G, HV>(a, I, f, xb, y0, lane, zb, ec, rg, half);
        for (int xcol = xb; xcol < xe; ++xcol) {
            if constexpr (STAGED) {
                float x[K];
                int c = xcol + 1;
                cp_async_commit_wait_all(true);
                if (c >= W && zb) {
#pragma unroll
                    for (int j = 0; j < K; ++j) x[j] = 0.f;
                } else {
                    if (c >= W) c = W - 1;
                    const double2* src = sMine + (c & 1) * RawStage<K>::STAGE_UNITS;
#pragma unroll
                    for (int v = 0; v < K / 2; ++v) {
                        double2 t = make_double2(0.0, 0.0);
                        if (rg.valid(2 * v)) t = src[v];
                        x[2 * v] = (float)t.x; x[2 * v + 1] = (float)t.y;
                    }
                }
                halo_from_col<K, G>(x, y0, H, lane, zb, ep);
                if (xcol + 1 < xe) stage_issue(xcol + 2);
            } else {
#if COLS_PREFETCH
            if (SRC == 1 && xcol + 1 + COLS_PREFETCH_DIST < W && y0 < H) {
                const double* pf = reinterpret_cast<const double*>(a.raw) + (size_t)f * a.raw_stride + (size_t)(xcol + 1 + COLS_PREFETCH_DIST) * a.raw_ld + y0;
#if COLS_PREFETCH == 1
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
#else
                asm volatile("prefetch.global.L1 [%0];" ::"l"(pf));
#endif
            }
#endif
            if constexpr (PF) {
                float x[K];
#pragma unroll
                for (int v = 0; v < K / 4; ++v) {
                    const unsigned wd = nxt_zero ? 0u : nxt[v];
                    x[4 * v] = u8_unit(wd & 0xffu); x[4 * v + 1] = u8_unit((wd >> 8) & 0xffu);
                    x[4 * v + 2] = u8_unit((wd >> 16) & 0xffu); x[4 * v + 3] = u8_unit(wd >> 24);
                }
                halo_from_col<K, G>(x, y0, H, lane, zb, ep);
                if (xcol + 1 < xe) pf_issue(xcol + 2);
            } else
            load_col_halo_any<K, SRC, G, HV>(a, I, f, xcol + 1, y0, lane, zb, ep, rg, half);
            }
            float pp[3][K];
            {
                float gi[2 * K];
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    const float s0 = 3.f / 16.f, s1 = 10.f / 16.f;
                    float gy = s0 * (0.5f * (em[j + 2] - em[j])) + s1 * (0.5f * (ec[j + 2] - ec[j])) + s0 * (0.5f * (ep[j + 2] - ep[j]));
                    float gx = s0 * (0.5f * (ep[j] - em[j])) + s1 * (0.5f * (ep[j + 1] - em[j + 1])) + s0 * (0.5f * (ep[j + 2] - em[j + 2]));
                    gi[2 * j] = gy; gi[2 * j + 1] = gx;
                    pp[0][j] = gy * gy; pp[1][j] = gx * gx; pp[2][j] = gy * gx;
                }
                float* og = fb + a.o_grad + (size_t)xcol * (2 * pitch);
#pragma unroll
                for (int v = 0; v < K / 2; ++v) {
                    const int y = y0 + 2 * v;
                    if constexpr (G > 1) {
                        if (rg.valid(2 * v)) *reinterpret_cast<float4*>(og + 2 * y) = make_float4(gi[4 * v], gi[4 * v + 1], gi[4 * v + 2], gi[4 * v + 3]);
                    } else if (y < pitch) {
                        float4 t;
                        t.x = y < H ? gi[4 * v] : 0.f; t.y = y < H ? gi[4 * v + 1] : 0.f;
                        t.z = y + 1 < H ? gi[4 * v + 2] : 0.f; t.w = y + 1 < H ? gi[4 * v + 3] : 0.f;
                        *reinterpret_cast<float4*>(og + 2 * y) = t;
                    }
                }
            }
            warp_iir_lines<K, 3, G, HV>(pp, H, lane, c4, false, half, xch4, a.pl4, bar);
            float* o0 = (a.t_base ? a.t_base + (size_t)((a.f0 + f) % a.t_ring) * a.t_stride + a.o_t : fb + a.o_out0) + (size_t)xcol * pitch;
            if constexpr (G > 1) {
                store_col_g<K, G>(o0, y0, rg, pp[0]);
                store_col_g<K, G>(o0 + a.plane_elems, y0, rg, pp[1]);
                store_col_g<K, G>(o0 + 2 * a.plane_elems, y0, rg, pp[2]);
            } else {
                store_col<K>(o0, y0, pitch, H, pp[0]);
                store_col<K>(o0 + a.plane_elems, y0, pitch, H, pp[1]);
                store_col<K>(o0 + 2 * a.plane_elems, y0, pitch, H, pp[2]);
            }
            if (SRC != 0 || a.do_blur) {
                float bl[1][K];
                if constexpr (G > 1) {
                    // rows >= H are never stored and the filter overwrites them with the border value
#pragma unroll
                    for (int j = 0; j < K; ++j) bl[0][j] = ec[j + 1];
                    if (SRC != 0) store_col_g<K, G>(fb + a.o_in + (size_t)xcol * pitch, y0, rg, bl[0]);  // the converted layer
                } else {
#pragma unroll
                    for (int j = 0; j < K; ++j) bl[0][j] = (y0 + j < H) ? ec[j + 1] : 0.f;
                    if (SRC != 0) store_col<K>(fb + a.o_in + (size_t)xcol * pitch, y0, pitch, H, bl[0]);  // the converted layer
                }
                if (a.do_blur) {
                    warp_iir_lines<K, 1, G, HV>(bl, H, lane, c1, zb, half, xch1, a.pl1, bar);
                    if (a.inv_n) {
#pragma unroll
                        for (int j = 0; j < K; ++j)
                            if (rg.valid(j)) bl[0][j] *= __ldg(a.inv_n + y0 + j);
                    }
                    if constexpr (G > 1) store_col_g<K, G>(fb + a.o_tmp + (size_t)xcol * pitch, y0, rg, bl[0]);
                    else store_col<K>(fb + a.o_tmp + (size_t)xcol * pitch, y0, pitch, H, bl[0]);
                }
            }
#pragma unroll
            for (int j = 0; j < K + 2; ++j) { em[j] = ec[j]; ec[j] = ep[j]; }
        }
    }
}


// ----------------------------------------------------------------------------------------------
// dim-2 kernel (along x).  CTA = LR rows x NC chunks of KRt elements; thread (row, chunk) keeps its chunk in registers.
// MODE 0: plain output (v * scale [* inv_n])       -- blur chain, and parity download of the smoothed planes
// MODE 1: exclusive prefix sum along x of v * scale, accumulated in Float64, W+1 columns (column 0 stays 0)
// ----------------------------------------------------------------------------------------------
struct RowArgs {
    FrameSet fs;
    int f0, n_frames;
    int H, W, pitch, nplanes;
    int zero_border;
    size_t o_in0, o_out0, plane_elems;  // nplanes consecutive input planes -> nplanes consecutive output planes
    const float* inv_n;                 // NA mode: 1/nx[x]
    // input planes in a scratch ring (TScratch) instead of at o_in0 inside the frame
    const float* t_base;
    size_t t_stride, o_t;
    int t_ring;
    // MODE 2 (blur x pass fused with the bilinear decimation): next level's layer plane and the source-index / weight tables
    int keep_blur;          // also store the blurred plane (LKCache.gaussian_filtered) at o_out0
    int Ho, Wo, pout;
    size_t o_next;
    const int* iy_tab; const float* wy_tab;   // per output row: 1-based source row iy (rows iy-1, iy 0-based) and weight
    const int* ix_tab; const float* wx_tab;   // per output column
};

// base + k columns as one IMAD.WIDE.U32 (u32 x u32 + u64)
__device__ __forceinline__ const float* row_ptr(const float* base, unsigned pitch4, unsigned k) {
    return reinterpret_cast<const float*>(reinterpret_cast<const char*>(base) + (unsigned long long)pitch4 * k);
}
__device__ __forceinline__ float* row_ptr(float* base, unsigned pitch4, unsigned k) {
    return reinterpret_cast<float*>(reinterpret_cast<char*>(base) + (unsigned long long)pitch4 * k);
}

template <int KRt, int LR, int MODE>
__global__ void __launch_bounds__(32 * LR, (LR == 16 && KRt <= 40) ? 2 : 1) k_rows(RowArgs a, IirDev c) {
    constexpr int NCMAX = 32;
    constexpr int LRP = LR + 1;  // padded: the carry scans read these arrays with lane = chunk
    __shared__ float sF[NCMAX][3][LRP];
    __shared__ float sB[NCMAX][3][LRP];
    __shared__ double sP[MODE == 1 ? NCMAX : 1][LRP];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int rl = threadIdx.x % LR;
    const int ch = threadIdx.x / LR;
    const int NC = (a.W + KRt - 1) / KRt;
    // MODE 2: the CTA owns 8 output rows of the next level; its LR = 16 source rows start at the first row they sample
    const int srow0 = MODE == 2 ? __ldg(a.iy_tab + min(8 * (int)blockIdx.x, a.Ho - 1)) - 1 : 0;
    const int r = MODE == 2 ? min(srow0 + rl, a.H - 1) : blockIdx.x * LR + rl;
    const int plane = blockIdx.y;
    const int f = blockIdx.z;
    const bool chok = ch < NC;  // the block is rounded up to whole warps
    const bool rowok = r < a.H && chok;
    float* fb = a.fs.frame(a.f0 + f);
    const float* in = (a.t_base ? a.t_base + (size_t)((a.f0 + f) % a.t_ring) * a.t_stride + a.o_t : fb + a.o_in0) + (size_t)plane * a.plane_elems + (rowok ? r : 0);
    float* out = fb + a.o_out0 + (size_t)plane * a.plane_elems + (rowok ? r : 0);
    const int x0 = chok ? ch * KRt : 0;
    const int n = a.W;
    const unsigned pitch4 = (unsigned)a.pitch * 4u;
    const float a1 = c.a1, a2 = c.a2, a3 = c.a3;
    const bool zb = a.zero_border != 0;
    const float iminus = (zb || !rowok) ? 0.f : __ldg(in);
    const float iplus = (zb || !rowok) ? 0.f : __ldg(row_ptr(in, pitch4, n - 1));
    const float um = iminus * c.inv1ma;

    // the line is extended to NC*KRt elements with its border value (see warp_iir_lines): every chunk is full
    float x[KRt];
    {
        const float* p = row_ptr(in, pitch4, x0);
#pragma unroll
        for (int j = 0; j < KRt; ++j) x[j] = __ldg(row_ptr(p, pitch4, (unsigned)max(min(j, n - 1 - x0), 0)));
        if (x0 + KRt > n) {
#pragma unroll
            for (int j = 0; j < KRt; ++j)
                if (x0 + j >= n) x[j] = iplus;
        }
        if (!rowok) {
#pragma unroll
            for (int j = 0; j < KRt; ++j) x[j] = 0.f;
        }
    }

    // forward phase 1
    float s0 = ch == 0 ? um : 0.f, s1 = s0, s2 = s0;
#pragma unroll
    for (int j = 0; j < KRt; ++j) {
        float u = fmaf(a1, s0, fmaf(a2, s1, fmaf(a3, s2, x[j])));
        s2 = s1; s1 = s0; s0 = u;
    }
    sF[ch][0][rl] = s0; sF[ch][1][rl] = s1; sF[ch][2][rl] = s2;
    __syncthreads();
    // carries: warp `wid` takes rows wid, wid+nwarp, ...; lane = chunk; Kogge-Stone scan q_k += A^(KR d) q_{k-d}
    for (int rr = wid; rr < LR; rr += nwarp) {
        float q0 = 0.f, q1 = 0.f, q2 = 0.f;
        if (lane < NC) { q0 = sF[lane][0][rr]; q1 = sF[lane][1][rr]; q2 = sF[lane][2][rr]; }
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            if (j >= c.nsr) break;
            const int d = 1 << j;
            float r0 = __shfl_up_sync(FULL, q0, d), r1 = __shfl_up_sync(FULL, q1, d), r2 = __shfl_up_sync(FULL, q2, d);
            if (lane >= d) mat3_acc(c.PR[j], r0, r1, r2, q0, q1, q2);
        }
        q0 = __shfl_up_sync(FULL, q0, 1); q1 = __shfl_up_sync(FULL, q1, 1); q2 = __shfl_up_sync(FULL, q2, 1);
        if (lane >= 1 && lane < NC) { sF[lane][0][rr] = q0; sF[lane][1][rr] = q1; sF[lane][2][rr] = q2; }  // incoming state of chunk `lane`
    }
    __syncthreads();
    if (ch == 0) { s0 = um; s1 = um; s2 = um; }
    else { s0 = sF[ch][0][rl]; s1 = sF[ch][1][rl]; s2 = sF[ch][2][rl]; }
#pragma unroll
    for (int j = 0; j < KRt; ++j) {
        float u = fmaf(a1, s0, fmaf(a2, s1, fmaf(a3, s2, x[j])));
        x[j] = u;
        s2 = s1; s1 = s0; s0 = u;
    }
    // right boundary at the end of the extended line: only the last chunk's values are used
    const float up = iplus * c.inv1ma, vp = up * c.inv1ma;
    const float d0 = s0 - up, d1 = s1 - up, d2 = s2 - up;
    const float vr0 = fmaf(c.M[0], d0, fmaf(c.M[1], d1, fmaf(c.M[2], d2, vp)));
    const float vr1 = fmaf(c.M[3], d0, fmaf(c.M[4], d1, fmaf(c.M[5], d2, vp)));
    const float vr2 = fmaf(c.M[6], d0, fmaf(c.M[7], d1, fmaf(c.M[8], d2, vp)));
    const bool lastc = ch == NC - 1;

    // backward phase 1
    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int j = KRt - 1; j >= 0; --j) {
        float v = fmaf(a1, t0, fmaf(a2, t1, fmaf(a3, t2, x[j])));
        t2 = t1; t1 = t0; t0 = v;
        if (j == KRt - 1 && lastc) { t0 = vr0; t1 = vr1; t2 = vr2; }
    }
    sB[ch][0][rl] = t0; sB[ch][1][rl] = t1; sB[ch][2][rl] = t2;
    __syncthreads();
    for (int rr = wid; rr < LR; rr += nwarp) {
        float q0 = 0.f, q1 = 0.f, q2 = 0.f;
        if (lane < NC) { q0 = sB[lane][0][rr]; q1 = sB[lane][1][rr]; q2 = sB[lane][2][rr]; }
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            if (j >= c.nsr) break;
            const int d = 1 << j;
            float r0 = __shfl_down_sync(FULL, q0, d), r1 = __shfl_down_sync(FULL, q1, d), r2 = __shfl_down_sync(FULL, q2, d);
            if (lane + d < 32) mat3_acc(c.PR[j], r0, r1, r2, q0, q1, q2);
        }
        q0 = __shfl_down_sync(FULL, q0, 1); q1 = __shfl_down_sync(FULL, q1, 1); q2 = __shfl_down_sync(FULL, q2, 1);
        if (lane < NC - 1) { sB[lane][0][rr] = q0; sB[lane][1][rr] = q1; sB[lane][2][rr] = q2; }
    }
    __syncthreads();
    if (!lastc) { t0 = sB[ch][0][rl]; t1 = sB[ch][1][rl]; t2 = sB[ch][2][rl]; }
    else { t0 = 0.f; t1 = 0.f; t2 = 0.f; }
    const float sc = c.scale;
#pragma unroll
    for (int j = KRt - 1; j >= 0; --j) {
        float v = fmaf(a1, t0, fmaf(a2, t1, fmaf(a3, t2, x[j])));
        if (j == KRt - 1 && lastc) { v = vr0; t0 = vr0; t1 = vr1; t2 = vr2; }
        else { t2 = t1; t1 = t0; t0 = v; }
        x[j] = v * sc;
    }

    const bool fullc = x0 + KRt <= n;
    if (MODE == 0 || (MODE == 2 && a.keep_blur)) {
        float* q = row_ptr(out, pitch4, x0);
        if (a.inv_n) {
#pragma unroll
            for (int j = 0; j < KRt; ++j)
                if (fullc || x0 + j < n) x[j] *= __ldg(a.inv_n + x0 + j);
        }
        if (rowok) {
#pragma unroll
            for (int j = 0; j < KRt; ++j)
                if (fullc || x0 + j < n) *row_ptr(q, pitch4, j) = x[j];
        }
    }
    if constexpr (MODE == 2) {
        // ---- bilinear decimation (pyramid.jl:120-121, [3P] imresize!) of the rows this CTA just blurred, same arithmetic and
        // order as k_resize: vertical lerp of the two source rows, then horizontal lerp of the two source columns.
        extern __shared__ float sV[];                   // 8 output rows x (NC * KRt + 1) vertically interpolated values
        const int wpad = NC * KRt + 1;
        if (!(MODE == 0 || a.keep_blur) && a.inv_n) {   // (the NA normalisation was applied above only when the plane is stored)
#pragma unroll
            for (int j = 0; j < KRt; ++j)
                if (fullc || x0 + j < n) x[j] *= __ldg(a.inv_n + x0 + j);
        }
        int mi = -1;
        float wy = 0.f;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int i = 8 * (int)blockIdx.x + m;
            if (i < a.Ho && __ldg(a.iy_tab + i) - 1 - srow0 == rl) { mi = m; wy = __ldg(a.wy_tab + i); }
        }
#pragma unroll
        for (int j = 0; j < KRt; ++j) {
            const float below = __shfl_down_sync(FULL, x[j], 1);  // row rl + 1 of the same chunk (rl <= 14 for every producing row)
            if (mi >= 0 && chok && (fullc || x0 + j < n)) sV[mi * wpad + x0 + j] = (1.f - wy) * x[j] + wy * below;
        }
        __syncthreads();
        float* nxt = fb + a.o_next;
        for (int idx = threadIdx.x; idx < 8 * a.Wo; idx += blockDim.x) {
            const int m = idx & 7, jo = idx >> 3;
            const int i = 8 * (int)blockIdx.x + m;
            if (i < a.Ho) {
                const int ix = __ldg(a.ix_tab + jo);
                const float wx = __ldg(a.wx_tab + jo);
                const float c0 = sV[m * wpad + ix - 1], c1 = sV[m * wpad + ix];
                nxt[i + (size_t)jo * a.pout] = (1.f - wx) * c0 + wx * c1;
            }
        }
    }
    if constexpr (MODE == 1) {
        // exclusive prefix along x in Float64: out[., j+1] = sum_{x' <= j} v[x']
        if (!fullc) {
#pragma unroll
            for (int j = 0; j < KRt; ++j)
                if (x0 + j >= n) x[j] = 0.f;  // the extension does not belong to the sums
        }
        // chunk totals in fp32 (4 partial sums), scanned across chunks in Float64; inside the chunk the running sum is
        // fp32 on top of a two-float (hi + lo) copy of the Float64 base, which keeps the conversion pipe (XU) out of the
        // per-element path.  Errors of the base are common to both ends of a window difference and cancel.
        float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
#pragma unroll
        for (int j = 0; j + 3 < KRt; j += 4) { p0 += x[j]; p1 += x[j + 1]; p2 += x[j + 2]; p3 += x[j + 3]; }
#pragma unroll
        for (int j = KRt & ~3; j < KRt; ++j) p0 += x[j];
        sP[ch][rl] = (double)((p0 + p1) + (p2 + p3));
        __syncthreads();
        for (int rr = wid; rr < LR; rr += nwarp) {  // exclusive scan of the chunk totals, lane = chunk
            double t = lane < NC ? sP[lane][rr] : 0.0;
            double inc = t;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                double o = __shfl_up_sync(FULL, inc, d);
                if (lane >= d) inc += o;
            }
            if (lane < NC) sP[lane][rr] = inc - t;
        }
        __syncthreads();
        const double base = sP[ch][rl];
        const float bh = (float)base, bl = (float)(base - (double)bh);
        if (rowok) {
            float* q = row_ptr(out, pitch4, x0 + 1);
            float local = 0.f;
#pragma unroll
            for (int j = 0; j < KRt; ++j) {
                local += x[j];
                if (fullc || x0 + j < n) *row_ptr(q, pitch4, j) = bh + (local + bl);
            }
        }
    }
}

// ----------------------------------------------------------------------------------------------
// bilinear decimation to ceil(size/2)  (pyramid.jl:120-121,132-133; [3P] imresize!)
// ----------------------------------------------------------------------------------------------
constexpr int RESIZE_CPB = 8;  // output columns per CTA

__global__ void k_resize(FrameSet fs, int f0, size_t o_in, int Hi, int Wi, int pin, size_t o_out, int Ho, int Wo, int pout) {
    const int f = blockIdx.z;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Ho) return;
    float* fb = fs.frame(f0 + f);
    const float* in = fb + o_in;
    float* out = fb + o_out;
    // 1-based source coordinate sf*(i1 - 0.5) + 0.5 with i1 = i+1; computed in double like the reference
    const double sy = (double)Hi / (double)Ho, sx = (double)Wi / (double)Wo;
    const double ry = sy * ((double)i + 0.5) + 0.5;
    int iy = (int)floor(ry);
    iy = min(max(iy, 1), Hi - 1);
    const float wy = (float)(ry - iy);
    // a CTA walks RESIZE_CPB output columns: eight times fewer, eight times longer CTAs than one column each
    const int j0 = blockIdx.y * RESIZE_CPB, j1 = min(j0 + RESIZE_CPB, Wo);
    for (int j = j0; j < j1; ++j) {
        const double rx = sx * ((double)j + 0.5) + 0.5;
        int ix = (int)floor(rx);
        ix = min(max(ix, 1), Wi - 1);
        const float wx = (float)(rx - ix);
        const float* p = in + (iy - 1) + (size_t)(ix - 1) * pin;
        const float c0 = (1.f - wy) * p[0] + wy * p[1];
        const float c1 = (1.f - wy) * p[pin] + wy * p[pin + 1];
        out[i + (size_t)j * pout] = (1.f - wx) * c0 + wx * c1;
    }
}

// ----------------------------------------------------------------------------------------------
// Levels beyond the register-tiled kernels (more than 1088 rows, level 0 more than 1280: a lane would own more than 34 rows of a column; more than 2048
// columns: more than 32 chunks per row) -- portrait 1080p, 4K frames.  Same stages, same planes, written for generality instead of
// speed: one thread per line runs the recursion sequentially in Float64 (forward pass stored as fp32 in the output plane, backward
// pass in place), a pointwise kernel forms the Scharr gradients and their products.  Coarser levels that fit go back to the
// register-tiled kernels.
// ----------------------------------------------------------------------------------------------
struct IirGen { double a1, a2, a3, scale, inv1ma, M[9]; };

static IirGen iir_gen(double sigma) {
    IirGen o;
    double a[3];
    iir_design(sigma, a, &o.scale, o.M);
    o.a1 = a[0]; o.a2 = a[1]; o.a3 = a[2];
    o.inv1ma = 1.0 / (1.0 - (a[0] + a[1] + a[2]));
    return o;
}

// Scharr (pyramid.jl:59,75,98-103; Fill(0) border on the ctor path, replicate on update!) + products: same fp32 expressions as k_cols_all
__global__ void k_gen_grad(FrameSet fs, int f0, int H, int W, int pitch, int zb, size_t o_in, size_t o_grad, size_t o_t, size_t plane_elems) {
    const int y = blockIdx.x * blockDim.x + threadIdx.x, x = blockIdx.y;
    if (y >= H) return;
    float* fb = fs.frame(f0 + blockIdx.z);
    const float* I = fb + o_in;
    auto at = [&](int yy, int xx) -> float {
        if (zb) return (yy < 0 || yy >= H || xx < 0 || xx >= W) ? 0.f : I[yy + (size_t)xx * pitch];
        yy = min(max(yy, 0), H - 1); xx = min(max(xx, 0), W - 1);
        return I[yy + (size_t)xx * pitch];
    };
    float em[3], ec[3], ep[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { em[k] = at(y - 1 + k, x - 1); ec[k] = at(y - 1 + k, x); ep[k] = at(y - 1 + k, x + 1); }
    const float s0 = 3.f / 16.f, s1 = 10.f / 16.f;
    const float gy = s0 * (0.5f * (em[2] - em[0])) + s1 * (0.5f * (ec[2] - ec[0])) + s0 * (0.5f * (ep[2] - ep[0]));
    const float gx = s0 * (0.5f * (ep[0] - em[0])) + s1 * (0.5f * (ep[1] - em[1])) + s0 * (0.5f * (ep[2] - em[2]));
    const size_t p = y + (size_t)x * pitch;
    reinterpret_cast<float2*>(fb + o_grad)[p] = make_float2(gy, gx);
    float* T = fb + o_t;
    T[p] = gy * gy; T[p + plane_elems] = gx * gx; T[p + 2 * plane_elems] = gy * gx;
}

// One thread per line of n elements with element stride es: out = IIR(in) [* inv_n[i]] (MODE 0), or the exclusive prefix sums of
// IIR(in) along the line into n + 1 elements (MODE 1: out[0] = 0, accumulated in Float64).  in == out is allowed for MODE 0.
// Lines are numbered fastest-first: line -> (i_line, plane, frame); the line's first element sits at i_line * ls.
template <int MODE>
__global__ void k_gen_iir(FrameSet fs, int f0, int n_lines, int nplanes, int n, size_t es, size_t ls, size_t o_in, size_t o_out, size_t plane_elems,
                          int zb, const float* __restrict__ inv_n, IirGen c) {
    const int il = blockIdx.x * blockDim.x + threadIdx.x;
    if (il >= n_lines) return;
    float* fb = fs.frame(f0 + blockIdx.z);
    const float* in = fb + o_in + (size_t)blockIdx.y * plane_elems + (size_t)il * ls;
    float* out = fb + o_out + (size_t)blockIdx.y * plane_elems + (size_t)il * ls + (MODE == 1 ? es : 0);
    const double iminus = zb ? 0.0 : (double)in[0], iplus = zb ? 0.0 : (double)in[(size_t)(n - 1) * es];
    const double um = iminus * c.inv1ma;
    // blocks of U elements: the loads of a block are independent of the recursion and of each other (the line is bound by memory
    // latency, not by the Float64 chain), and a block is read before it is written, so in == out stays legal
    constexpr int U = 8;
    double s0 = um, s1 = um, s2 = um;
    for (int i0 = 0; i0 < n; i0 += U) {
        float xv[U];
#pragma unroll
        for (int k = 0; k < U; ++k) xv[k] = i0 + k < n ? in[(size_t)(i0 + k) * es] : 0.f;
#pragma unroll
        for (int k = 0; k < U; ++k) {
            if (i0 + k < n) {
                const double u = (double)xv[k] + c.a1 * s0 + c.a2 * s1 + c.a3 * s2;
                out[(size_t)(i0 + k) * es] = (float)u;
                s2 = s1; s1 = s0; s0 = u;
            }
        }
    }
    // right boundary (Triggs & Sdika): (v[n-1], v[n], v[n+1]) from the last three forward values
    const double up = iplus * c.inv1ma, vp = up * c.inv1ma;
    const double d0 = s0 - up, d1 = s1 - up, d2 = s2 - up;
    double t0 = c.M[0] * d0 + c.M[1] * d1 + c.M[2] * d2 + vp;
    double t1 = c.M[3] * d0 + c.M[4] * d1 + c.M[5] * d2 + vp;
    double t2 = c.M[6] * d0 + c.M[7] * d1 + c.M[8] * d2 + vp;
    {
        float v = (float)(t0 * c.scale);
        if (MODE == 0 && inv_n) v *= inv_n[n - 1];
        out[(size_t)(n - 1) * es] = v;
    }
    for (int i1 = n - 2; i1 >= 0; i1 -= U) {   // elements i1, i1 - 1, ..., i1 - U + 1
        float xv[U], nv[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            xv[k] = i1 - k >= 0 ? out[(size_t)(i1 - k) * es] : 0.f;
            nv[k] = (MODE == 0 && inv_n && i1 - k >= 0) ? inv_n[i1 - k] : 1.f;
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
            if (i1 - k >= 0) {
                const double v = (double)xv[k] + c.a1 * t0 + c.a2 * t1 + c.a3 * t2;
                t2 = t1; t1 = t0; t0 = v;
                float o = (float)(v * c.scale);
                if (MODE == 0 && inv_n) o *= nv[k];
                out[(size_t)(i1 - k) * es] = o;
            }
        }
    }
    if (MODE == 1) {
        double run = 0.0;
        for (int i0 = 0; i0 < n; i0 += U) {
            float xv[U];
#pragma unroll
            for (int k = 0; k < U; ++k) xv[k] = i0 + k < n ? out[(size_t)(i0 + k) * es] : 0.f;
#pragma unroll
            for (int k = 0; k < U; ++k) {
                if (i0 + k < n) { run += (double)xv[k]; out[(size_t)(i0 + k) * es] = (float)run; }
            }
        }
    }
}

// line filters of one level: along y (lines = columns) or along x (lines = rows)
static void gen_iir(cudaStream_t s, FrameSet fs, int f0, int n_frames, const LevelGeom& L, bool along_y, int mode, int nplanes, size_t o_in, size_t o_out,
                    bool zb, const float* inv_n, double sigma) {
    const int n_lines = along_y ? L.W : L.H, n = along_y ? L.H : L.W;
    const size_t es = along_y ? 1 : (size_t)L.pitch, ls = along_y ? (size_t)L.pitch : 1;
    const int threads = along_y ? 32 : 128;  // columns are read with a stride: few lines per CTA keep their cache lines resident
    dim3 grid((n_lines + threads - 1) / threads, nplanes, n_frames);
    const IirGen c = iir_gen(sigma);
    if (mode == 1) k_gen_iir<1><<<grid, threads, 0, s>>>(fs, f0, n_lines, nplanes, n, es, ls, o_in, o_out, L.plane_elems, zb ? 1 : 0, inv_n, c);
    else k_gen_iir<0><<<grid, threads, 0, s>>>(fs, f0, n_lines, nplanes, n, es, ls, o_in, o_out, L.plane_elems, zb ? 1 : 0, inv_n, c);
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
int pick_K(int H) {
    const int need = (H + 31) / 32;
    const int ks[] = {2, 4, 6, 8, 12, 16, 24, 34};
    for (int k : ks)
        if (k >= need) return k;
    return 0;
}

// does the level fit the register-tiled kernels?  (SLAMKLT_FORCE_GENERIC=1 sends every level through the general kernels: tests)
// Up to 1088 rows one warp, or a pair of warps, owns a column; level 0 -- always built by the fused column kernel -- also takes
// the pair kernel's full range of 1280 rows (coarser levels need the one-warp blur kernel of the layer chain as well).
static int pick_K_pair(int H);
static bool level_tiled(const LevelGeom& L, int level) {
    if (L.W > 2048 || getenv("SLAMKLT_FORCE_GENERIC") != nullptr) return false;
    return pick_K(L.H) != 0 || (level == 0 && pick_K_pair(L.H) != 0);
}

template <int K>
static void launch_cols_blur(cudaStream_t s, const ColArgs& a, const IirDev& c) {
    const int maxb = 148 * 16;
    const int total_warps = a.n_frames * a.W;
    const int wpb = 8;
    int blocks = (total_warps + wpb - 1) / wpb;
    if (blocks > maxb) blocks = maxb;
    constexpr int GA = (K % 4 == 0) ? 4 : 2;
    if (a.H % GA == 0 && getenv("SLAMKLT_COLS_GENERIC") == nullptr) k_cols_blur<K, GA><<<blocks, wpb * 32, 0, s>>>(a, c);
    else k_cols_blur<K, 1><<<blocks, wpb * 32, 0, s>>>(a, c);
}

static void dispatch_cols_blur(cudaStream_t s, int K, const ColArgs& a, const IirDev& c) {
    switch (K) {
        case 2: launch_cols_blur<2>(s, a, c); break;
        case 4: launch_cols_blur<4>(s, a, c); break;
        case 6: launch_cols_blur<6>(s, a, c); break;
        case 8: launch_cols_blur<8>(s, a, c); break;
        case 12: launch_cols_blur<12>(s, a, c); break;
        case 16: launch_cols_blur<16>(s, a, c); break;
        case 24: launch_cols_blur<24>(s, a, c); break;
        case 34: launch_cols_blur<34>(s, a, c); break;
    }
}

static void dispatch_rows(cudaStream_t s, const RowArgs& a, const IirDev& c, int mode) {
    // 16 rows per CTA, two CTAs per SM; 32 rows per CTA (one 1024-thread CTA per SM, 128-byte row segments) measured slower:
    // prefix kernel of level 0 0.231 vs 0.176 ms
    if (a.W <= 32 * 40) {
        dim3 grid((a.H + 15) / 16, a.nplanes, a.n_frames);
        const int NC = (a.W + 39) / 40;
        const int threads = ((16 * NC + 31) / 32) * 32;
        if (mode == 0) k_rows<40, 16, 0><<<grid, threads, 0, s>>>(a, c);
        else k_rows<40, 16, 1><<<grid, threads, 0, s>>>(a, c);
    } else {
        dim3 grid((a.H + 15) / 16, a.nplanes, a.n_frames);
        const int NC = (a.W + 63) / 64;
        const int threads = ((16 * NC + 31) / 32) * 32;
        if (mode == 0) k_rows<64, 16, 0><<<grid, threads, 0, s>>>(a, c);
        else k_rows<64, 16, 1><<<grid, threads, 0, s>>>(a, c);
    }
}

static int krow_of(int W) { return W <= 32 * 40 ? 40 : 64; }

// blur x pass fused with the decimation (MODE 2): 8 output rows per CTA, 16 source rows, dynamic shared memory for the
// vertically interpolated rows
static void dispatch_rows_resize(cudaStream_t s, const RowArgs& a, const IirDev& c) {
    dim3 grid((a.Ho + 7) / 8, 1, a.n_frames);
    if (a.W <= 32 * 40) {
        const int NC = (a.W + 39) / 40;
        const int threads = ((16 * NC + 31) / 32) * 32;
        const size_t smem = (size_t)8 * (NC * 40 + 1) * sizeof(float);
        static const bool once = [] { cudaFuncSetAttribute(k_rows<40, 16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (32 * 40 + 1) * 4); return true; }();
        (void)once;
        k_rows<40, 16, 2><<<grid, threads, smem, s>>>(a, c);
    } else {
        const int NC = (a.W + 63) / 64;
        const int threads = ((16 * NC + 31) / 32) * 32;
        const size_t smem = (size_t)8 * (NC * 64 + 1) * sizeof(float);
        static const bool once = [] { cudaFuncSetAttribute(k_rows<64, 16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (32 * 64 + 1) * 4); return true; }();
        (void)once;
        k_rows<64, 16, 2><<<grid, threads, smem, s>>>(a, c);
    }
}

// source index / weight tables of the bilinear decimation n_in -> n_out ([3P] imresize!: 1-based source coordinate
// sf * (i - 1/2) + 1/2, computed in double like the reference), cached on the device per (n_in, n_out)
struct ResizeTab { int* idx; float* w; };
static ResizeTab resize_table(int n_in, int n_out) {
    static std::map<std::pair<int, std::pair<int, int>>, ResizeTab> cache;  // (device, (n_in, n_out))
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    int dev = 0;
    cudaGetDevice(&dev);
    auto key = std::make_pair(dev, std::make_pair(n_in, n_out));
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    std::vector<int> idx(n_out);
    std::vector<float> w(n_out);
    const double sf = (double)n_in / (double)n_out;
    for (int i = 0; i < n_out; ++i) {
        const double r = sf * ((double)i + 0.5) + 0.5;
        int ii = (int)std::floor(r);
        ii = std::min(std::max(ii, 1), n_in - 1);
        idx[i] = ii; w[i] = (float)(r - ii);
    }
    ResizeTab t{nullptr, nullptr};
    cudaMalloc(&t.idx, sizeof(int) * n_out);
    cudaMalloc(&t.w, sizeof(float) * n_out);
    cudaMemcpy(t.idx, idx.data(), sizeof(int) * n_out, cudaMemcpyHostToDevice);
    cudaMemcpy(t.w, w.data(), sizeof(float) * n_out, cudaMemcpyHostToDevice);
    cache[key] = t;
    return t;
}

// Strip width of the fused column kernel.  A strip of cs columns costs about cs + 0.6 column-times (two halo columns that only
// load); the grid runs in ceil(units / resident warps) rounds and the last round is rarely full, so the width is chosen to
// minimise rounds * (cs + 0.6): at 64 KITTI frames cs = 8 gives 4.2 -> 5 rounds (84 % full), cs = 7 gives 4.8 -> 5 rounds (96 %).
// With few frames the same rule picks narrow strips so that every SM gets warps.
static int pick_strip(int n_frames, int W, int resident_warps) {
    int best = GRAD_CS;
    double best_cost = 1e300;
    for (int cs = 1; cs <= 12; ++cs) {
        const long long units = (long long)n_frames * ((W + cs - 1) / cs);
        const long long rounds = (units + resident_warps - 1) / resident_warps;
        const double cost = (double)rounds * (cs + 0.6);
        if (cost < best_cost * (1.0 - 1e-9)) { best_cost = cost; best = cs; }
    }
    return best;
}

template <typename Kern>
static int resident_warps_of(Kern kern, int threads) {
    static std::map<const void*, int> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find((const void*)kern);
    if (it != cache.end()) return it->second;
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0);
    const int r = sms * (per_sm > 0 ? per_sm : 1) * (threads / 32);
    cache[(const void*)kern] = r;
    return r;
}

template <int K, int SRC, int G, int HV>
static void launch_cols_all_k(cudaStream_t s, ColArgs a, const IirDev& c4, const IirDev& c1) {
    const int wpb = 4;
    auto kern = k_cols_all<K, SRC, G, HV>;
    static const int forced = [] { const char* e = getenv("SLAMKLT_COLS_STRIP"); return e ? atoi(e) : 0; }();
    a.strip = forced > 0 ? forced : pick_strip(a.n_frames, a.W, resident_warps_of(kern, wpb * 32) / HV);
    const int total_warps = HV * a.n_frames * ((a.W + a.strip - 1) / a.strip);
    const int blocks = (total_warps + wpb - 1) / wpb;  // one strip per warp (pair); the hardware hands CTAs to SMs as slots free up
    kern<<<blocks, wpb * 32, 0, s>>>(a, c4, c1);
}

template <int K, int G, int HV>
static void launch_cols_all_g(cudaStream_t s, int src, const ColArgs& a, const IirDev& c4, const IirDev& c1) {
    switch (src) {
        case 0: launch_cols_all_k<K, 0, G, HV>(s, a, c4, c1); break;
        case 1: launch_cols_all_k<K, 1, G, HV>(s, a, c4, c1); break;
        case 2: launch_cols_all_k<K, 2, G, HV>(s, a, c4, c1); break;
        default: launch_cols_all_k<K, 3, G, HV>(s, a, c4, c1); break;
    }
}

template <int K, int HV = 1>
static void launch_cols_all(cudaStream_t s, int src, const ColArgs& a, const IirDev& c4, const IirDev& c1) {
    // group-aligned variant when the image height is a multiple of the row-group size (4 rows, or 2 when K is not a multiple of 4)
    constexpr int GA = (K % 4 == 0) ? 4 : 2;
    const bool aligned = a.H % GA == 0 && (src != 1 || ((a.raw_ld & 1) == 0 && (reinterpret_cast<uintptr_t>(a.raw) & 15) == 0)) &&
                         getenv("SLAMKLT_COLS_GENERIC") == nullptr;
    if (aligned) launch_cols_all_g<K, GA, HV>(s, src, a, c4, c1);
    else launch_cols_all_g<K, 1, HV>(s, src, a, c4, c1);
}

// Two warps per column for frames taller than 512 rows: rows per lane (0: one warp per column).  SLAMKLT_COLS_PAIR=0 keeps the
// one-warp kernels (K = 24 / 34 rows per lane, 255 registers).
static int pick_K_pair(int H) {
    const char* e = getenv("SLAMKLT_COLS_PAIR");
    if (e && atoi(e) == 0) return 0;
#if COLS_PAIR_SMALL
    if (H > 192 && H <= 384) return 6;
    if (H > 384 && H <= 512) return 8;
#endif
    if (H <= 512) return 0;
    return H <= 64 * 12 ? 12 : (H <= 64 * COLS_PAIR_KBIG ? COLS_PAIR_KBIG : 0);
}

// per-lane matrices A^(K (l + 1)), l = 0 .. 31, of the recursion with the given sigma (fp32, row-major 3 x 3), cached on the device
static const float* lane_powers(double sigma, int K) {
    static std::map<std::pair<int, std::pair<long long, int>>, float*> cache;  // (device, (sigma bits, K))
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    int dev = 0;
    cudaGetDevice(&dev);
    long long bits;
    static_assert(sizeof(bits) == sizeof(sigma), "");
    memcpy(&bits, &sigma, sizeof(bits));
    auto key = std::make_pair(dev, std::make_pair(bits, K));
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    double a[3], sc, M[9];
    iir_design(sigma, a, &sc, M);
    const double A[9] = {a[0], a[1], a[2], 1, 0, 0, 0, 1, 0};
    double AK[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, P[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    auto mul = [](const double* X, const double* Y, double* Z) {
        double t[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) { double v = 0; for (int k = 0; k < 3; ++k) v += X[3 * i + k] * Y[3 * k + j]; t[3 * i + j] = v; }
        for (int i = 0; i < 9; ++i) Z[i] = t[i];
    };
    for (int i = 0; i < K; ++i) mul(AK, A, AK);  // A^K
    std::vector<float> tab(32 * 9);
    for (int l = 0; l < 32; ++l) {
        mul(P, AK, P);                           // A^(K (l + 1))
        for (int i = 0; i < 9; ++i) tab[l * 9 + i] = (float)P[i];
    }
    float* d = nullptr;
    cudaMalloc(&d, tab.size() * sizeof(float));
    cudaMemcpy(d, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice);
    cache[key] = d;
    return d;
}

static void dispatch_cols_all(cudaStream_t s, int K, int src, const ColArgs& a, const IirDev& c4, const IirDev& c1) {
    switch (K) {
        case 2: launch_cols_all<2>(s, src, a, c4, c1); break;
        case 4: launch_cols_all<4>(s, src, a, c4, c1); break;
        case 6: launch_cols_all<6>(s, src, a, c4, c1); break;
        case 8: launch_cols_all<8>(s, src, a, c4, c1); break;
        case 12: launch_cols_all<12>(s, src, a, c4, c1); break;
        case 16: launch_cols_all<16>(s, src, a, c4, c1); break;
        case 24: launch_cols_all<24>(s, src, a, c4, c1); break;
        case 34: launch_cols_all<34>(s, src, a, c4, c1); break;
    }
}
static void dispatch_cols_all_pair(cudaStream_t s, int K2, int src, const ColArgs& a, const IirDev& c4, const IirDev& c1) {
#if COLS_PAIR_SMALL
    if (K2 == 6) { launch_cols_all<6, 2>(s, src, a, c4, c1); return; }
    if (K2 == 8) { launch_cols_all<8, 2>(s, src, a, c4, c1); return; }
#endif
    if (K2 == 12) launch_cols_all<12, 2>(s, src, a, c4, c1);
    else launch_cols_all<COLS_PAIR_KBIG, 2>(s, src, a, c4, c1);
}

// Build the pyramids of n_frames frames.  raw != nullptr: level 0 is read from the staged host image (dtype, compact
// ld = H0) and the layer plane is written on the way; raw == nullptr: the level-0 layer is already in place.
// Schedule (3 streams): main runs the layer chain -- level 0: k_cols_all (conversion + gradients + blur y pass), then per
// level k_rows blur -> k_resize -> [k_cols_blur of the next level]; the structure planes of level 0 (k_rows prefix) go to
// side stream B, and the gradient stages of the coarser levels (k_cols_grad + k_rows prefix) to side stream C as soon as
// their layer exists, so the small coarse-level kernels overlap the layer chain instead of extending it.
int launch_pyramid(const PyrStreams& ps, FrameSet fs, int f0, int n_frames, const PyrGeom& g, double sigma, int mode,
                   const float* const* inv_ny, const float* const* inv_nx, const void* raw, int dtype, const Hook* hk, const TScratch* ts,
                   bool join) {
    int launches = 0;
    char nm[48];
    const bool ctor = mode == SLAMKLT_MODE_CTOR;
    const bool par = ps.parallel && hk == nullptr;  // per-kernel profiling needs a serial stream
    cudaStream_t sA = ps.main, sB = par ? ps.b : ps.main, sC = par ? ps.c : ps.main;
    cudaEvent_t evB = ps.ev[MAX_LAYERS + 1], evC = ps.ev[MAX_LAYERS + 2];

    auto col_args = [&](int l) {
        const LevelGeom& L = g.lv[l];
        ColArgs ca{};
        ca.fs = fs; ca.f0 = f0; ca.n_frames = n_frames; ca.H = L.H; ca.W = L.W; ca.pitch = L.pitch;
        ca.zero_border = ctor; ca.o_in = plane_off(L, DP_I); ca.o_out0 = plane_off(L, DP_T0);
        ca.o_grad = plane_off(L, DP_GRAD); ca.plane_elems = L.plane_elems;
        ca.o_tmp = plane_off(L, DP_TMP);
        ca.raw = nullptr; ca.raw_ld = g.H0; ca.raw_stride = (size_t)g.H0 * g.W0;
        if (ts && ts->base) { ca.t_base = ts->base; ca.t_stride = ts->stride; ca.t_ring = ts->ring; ca.o_t = ts->off[l]; }
        return ca;
    };
    auto rows_struct = [&](cudaStream_t s, int l, const IirDev& c4) {
        const LevelGeom& L = g.lv[l];
        RowArgs ra{};
        ra.fs = fs; ra.f0 = f0; ra.n_frames = n_frames; ra.H = L.H; ra.W = L.W; ra.pitch = L.pitch; ra.nplanes = 3;
        ra.zero_border = 0; ra.o_in0 = plane_off(L, DP_T0); ra.o_out0 = plane_off(L, DP_RYY); ra.plane_elems = L.plane_elems;
        ra.inv_n = nullptr;
        if (ts && ts->base) { ra.t_base = ts->base; ra.t_stride = ts->stride; ra.t_ring = ts->ring; ra.o_t = ts->off[l]; }
        snprintf(nm, sizeof(nm), "k_rows_struct_L%d", l); mark(hk, nm);
        dispatch_rows(s, ra, c4, 1);
        launches += 1;
    };

    for (int l = 0; l < g.nl; ++l) {
        const LevelGeom& L = g.lv[l];
        const bool blur = l + 1 < g.nl;
        if (!level_tiled(L, l)) {
            // general kernels, all on the main stream: [conversion] -> gradients + products -> sigma = 4 along y (in place) -> along x
            // with the row prefix sums -> [pyramid blur along y, along x -> decimation]
            if (l == 0 && raw) {
                mark(hk, "k_convert");
                dim3 cgrid((L.H + 127) / 128, L.W, n_frames);
                const size_t st = (size_t)g.H0 * g.W0, oI = plane_off(L, DP_I);
                if (dtype == SLAMKLT_F64) k_convert<double><<<cgrid, 128, 0, sA>>>((const double*)raw, g.H0, st, fs, f0, oI, L.H, L.W, L.pitch, nullptr);
                else if (dtype == SLAMKLT_F32) k_convert<float><<<cgrid, 128, 0, sA>>>((const float*)raw, g.H0, st, fs, f0, oI, L.H, L.W, L.pitch, nullptr);
                else k_convert<uint8_t><<<cgrid, 128, 0, sA>>>((const uint8_t*)raw, g.H0, st, fs, f0, oI, L.H, L.W, L.pitch, nullptr);
                launches += 1;
            }
            snprintf(nm, sizeof(nm), "k_gen_grad_L%d", l); mark(hk, nm);
            k_gen_grad<<<dim3((L.H + 127) / 128, L.W, n_frames), 128, 0, sA>>>(fs, f0, L.H, L.W, L.pitch, ctor ? 1 : 0, plane_off(L, DP_I),
                                                                              plane_off(L, DP_GRAD), plane_off(L, DP_T0), L.plane_elems);
            snprintf(nm, sizeof(nm), "k_gen_iir_y_L%d", l); mark(hk, nm);
            gen_iir(sA, fs, f0, n_frames, L, true, 0, 3, plane_off(L, DP_T0), plane_off(L, DP_T0), false, nullptr, 4.0);
            snprintf(nm, sizeof(nm), "k_gen_iir_x_prefix_L%d", l); mark(hk, nm);
            gen_iir(sA, fs, f0, n_frames, L, false, 1, 3, plane_off(L, DP_T0), plane_off(L, DP_RYY), false, nullptr, 4.0);
            launches += 3;
            if (blur) {
                const LevelGeom& N = g.lv[l + 1];
                snprintf(nm, sizeof(nm), "k_gen_blur_y_L%d", l); mark(hk, nm);
                gen_iir(sA, fs, f0, n_frames, L, true, 0, 1, plane_off(L, DP_I), plane_off(L, DP_TMP), ctor, ctor ? inv_ny[l] : nullptr, sigma);
                snprintf(nm, sizeof(nm), "k_gen_blur_x_L%d", l); mark(hk, nm);
                gen_iir(sA, fs, f0, n_frames, L, false, 0, 1, plane_off(L, DP_TMP), plane_off(L, DP_BLUR), ctor, ctor ? inv_nx[l] : nullptr, sigma);
                snprintf(nm, sizeof(nm), "k_resize_L%d", l); mark(hk, nm);
                const int nby = (N.H + 127) / 128;
                const int rthreads = (((N.H + nby - 1) / nby) + 31) / 32 * 32;
                k_resize<<<dim3(nby, (N.W + RESIZE_CPB - 1) / RESIZE_CPB, n_frames), rthreads, 0, sA>>>(
                    fs, f0, plane_off(L, DP_BLUR), L.H, L.W, L.pitch, plane_off(N, DP_I), N.H, N.W, N.pitch);
                launches += 3;
            }
            continue;
        }
        const int K = pick_K(L.H);
        IirDev c4, c1;
        iir_dev(4.0, K, krow_of(L.W), &c4);  // lucas_kanade.jl:112
        iir_dev(sigma, K, krow_of(L.W), &c1);
        ColArgs ca = col_args(l);
        ca.inv_n = (ctor && blur) ? inv_ny[l] : nullptr;
        ca.do_blur = blur;
        // frames taller than 512 rows: the fused column kernel runs with two warps per column (own scan matrices: K2 rows per lane)
        const int K2 = pick_K_pair(L.H);
        IirDev c4p, c1p;
        if (K2) {
            iir_dev(4.0, K2, krow_of(L.W), &c4p);
            iir_dev(sigma, K2, krow_of(L.W), &c1p);
            ca.pl4 = lane_powers(4.0, K2);
            ca.pl1 = lane_powers(sigma, K2);
        }
        if (l == 0 || !par) {
            // fused column kernel on the main stream
            const int src = (l == 0 && raw) ? (dtype == SLAMKLT_F64 ? 1 : (dtype == SLAMKLT_F32 ? 2 : 3)) : 0;
            ca.raw = src ? raw : nullptr;
            snprintf(nm, sizeof(nm), "k_cols_all_L%d", l); mark(hk, nm);
            if (K2) dispatch_cols_all_pair(sA, K2, src, ca, c4p, c1p);
            else dispatch_cols_all(sA, K, src, ca, c4, c1);
            launches += 1;
            // with a scratch ring the x pass follows on the same stream: the next group's column kernel (also on this stream)
            // overwrites the ring slots, and the planes are still hot in L2
            const bool ring = ts && ts->base;
            if (par && !ring) { cudaEventRecord(ps.ev[0], sA); cudaStreamWaitEvent(sB, ps.ev[0], 0); }
            rows_struct(par && !ring ? sB : sA, l, c4);
        } else {
            // coarser levels: the layer chain only needs the blur y pass; gradients + structure planes run on stream C
            cudaEventRecord(ps.ev[l], sA);  // layer l exists (k_resize of level l-1 ran on main)
            cudaStreamWaitEvent(sC, ps.ev[l], 0);
            ColArgs cg = ca;
            cg.inv_n = nullptr;
            cg.do_blur = 0;  // the fused kernel without its blur half (SRC = 0: nothing to convert)
            snprintf(nm, sizeof(nm), "k_cols_grad_L%d", l); mark(hk, nm);
            if (K2) dispatch_cols_all_pair(sC, K2, 0, cg, c4p, c1p);
            else dispatch_cols_all(sC, K, 0, cg, c4, c1);
            launches += 1;
            rows_struct(sC, l, c4);
            if (blur) {
                ColArgs cb = ca;
                cb.o_out0 = plane_off(L, DP_TMP);
                snprintf(nm, sizeof(nm), "k_cols_blur_L%d", l); mark(hk, nm);
                dispatch_cols_blur(sA, K, cb, c1);
                launches += 1;
            }
        }
        if (blur) {
            const LevelGeom& N = g.lv[l + 1];
            RowArgs rb{};
            rb.fs = fs; rb.f0 = f0; rb.n_frames = n_frames; rb.H = L.H; rb.W = L.W; rb.pitch = L.pitch; rb.nplanes = 1;
            rb.zero_border = ctor; rb.o_in0 = plane_off(L, DP_TMP); rb.o_out0 = plane_off(L, DP_BLUR); rb.plane_elems = L.plane_elems;
            rb.inv_n = ctor ? inv_nx[l] : nullptr;
            // blur x pass and decimation in one kernel: the blurred plane is only stored when somebody may read it back
            // (single pyramids: LKCache.gaussian_filtered, parity downloads); SLAMKLT_NO_FUSED_RESIZE=1 keeps the two kernels
            static const bool fused = getenv("SLAMKLT_NO_FUSED_RESIZE") == nullptr;
            if (fused && L.H >= 16) {
                const ResizeTab ty = resize_table(L.H, N.H), tx = resize_table(L.W, N.W);
                rb.keep_blur = n_frames == 1 ? 1 : 0;
                rb.Ho = N.H; rb.Wo = N.W; rb.pout = N.pitch; rb.o_next = plane_off(N, DP_I);
                rb.iy_tab = ty.idx; rb.wy_tab = ty.w; rb.ix_tab = tx.idx; rb.wx_tab = tx.w;
                snprintf(nm, sizeof(nm), "k_rows_blur_resize_L%d", l); mark(hk, nm);
                dispatch_rows_resize(sA, rb, c1);
                launches += 1;
                continue;
            }
            snprintf(nm, sizeof(nm), "k_rows_blur_L%d", l); mark(hk, nm);
            dispatch_rows(sA, rb, c1, 0);
            snprintf(nm, sizeof(nm), "k_resize_L%d", l); mark(hk, nm);
            // block = an even share of the output column rounded up to whole warps (188 rows -> 2 x 96 threads, not 128 + 60)
            const int nby = (N.H + 127) / 128;
            const int rthreads = (((N.H + nby - 1) / nby) + 31) / 32 * 32;
            dim3 grid(nby, (N.W + RESIZE_CPB - 1) / RESIZE_CPB, n_frames);
            k_resize<<<grid, rthreads, 0, sA>>>(fs, f0, plane_off(L, DP_BLUR), L.H, L.W, L.pitch, plane_off(N, DP_I), N.H, N.W, N.pitch);
            launches += 2;
        }
    }
    if (par && join) {
        cudaEventRecord(evB, sB);
        cudaEventRecord(evC, sC);
        cudaStreamWaitEvent(sA, evB, 0);
        cudaStreamWaitEvent(sA, evC, 0);
    }
    return launches;
}

// parity access: smoothed product plane `which` (0 yy, 1 xx, 2 yx) of one level, recomputed from the y-filtered
// scratch plane into DP_TMP (the pyramid itself only keeps the row-prefix form)
int launch_smoothed_plane(cudaStream_t s, FrameSet fs, int f0, const PyrGeom& g, int level, int which, const Hook* hk) {
    const LevelGeom& L = g.lv[level];
    if (!level_tiled(L, level)) {
        mark(hk, "k_gen_iir_x_plain");
        gen_iir(s, fs, f0, 1, L, false, 0, 1, plane_off(L, DP_T0 + which), plane_off(L, DP_TMP), false, nullptr, 4.0);
        return 1;
    }
    IirDev c;
    iir_dev(4.0, pick_K(L.H), krow_of(L.W), &c);
    RowArgs ra{};
    ra.fs = fs; ra.f0 = f0; ra.n_frames = 1; ra.H = L.H; ra.W = L.W; ra.pitch = L.pitch; ra.nplanes = 1;
    ra.zero_border = 0; ra.o_in0 = plane_off(L, DP_T0 + which); ra.o_out0 = plane_off(L, DP_TMP); ra.plane_elems = L.plane_elems;
    ra.inv_n = nullptr;
    mark(hk, "k_rows_plain");
    dispatch_rows(s, ra, c, 0);
    return 1;
}

}  // namespace sk
