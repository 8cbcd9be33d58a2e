// Shi-Tomasi extraction with avoidance mask, strict 3x3 NMS, per-cell top-k and grid bucketing (sm_100a).
//
// Reference behaviour: extractor.jl:24-42 (_shi_tomasi), :63-95 (detect), :116-122 (get_mask); third-party
// semantics (Images.shi_tomasi / findlocalmaxima, Kernel.gaussian, ImageDraw circle) as listed in SURVEY.md A.8-A.10.
//
// One CTA per 35x35 cell keeps the whole cell in shared memory.  The path is Float64 end to end and this file
// is compiled with -fmad=false so that responses are bit-identical to the reference arithmetic order: keypoint
// sets must be identical except for exact score ties, and near-ties must not flip either.
#include "common.cuh"

namespace sk {

constexpr int DET_THREADS = 256;
constexpr int MAX_NEAR = 1024;  // current points kept in shared memory per cell

size_t detect_smem_bytes(int cs, int hw) {
    const size_t cell = (size_t)cs * cs;
    const size_t reg = (size_t)(cs + 2 * hw) * (cs + 2 * hw);
    size_t b = 0;
    b += 5 * cell * sizeof(double);               // img, gyy, gyx, gxx, R
    b += (size_t)cs * (cs + 2 * hw) * sizeof(double);  // tmp after the y pass of the mask blur
    b += reg * sizeof(float);                     // m0
    b += cell * sizeof(double) + cell * sizeof(int);  // candidates
    b += MAX_NEAR * 2 * sizeof(int);
    b += 64 * sizeof(int);
    return b;
}

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

    const size_t cell = (size_t)cs * cs;
    double* s_img = (double*)smem_raw;
    double* s_gyy = s_img + cell;
    double* s_gyx = s_gyy + cell;
    double* s_gxx = s_gyx + cell;
    double* s_R = s_gxx + cell;
    double* s_tmp = s_R + cell;
    double* s_candr = s_tmp + (size_t)cs * (cs + 2 * hw);
    float* s_m0 = (float*)(s_candr + cell);
    int* s_candi = (int*)(s_m0 + (size_t)(cs + 2 * hw) * (cs + 2 * hw));
    int* s_near = s_candi + cell;
    int* s_misc = s_near + 2 * MAX_NEAR;  // [0] near count, [1..8] warp scan scratch

    const double* img = a.img + (size_t)f * H * W;
    const int tid = threadIdx.x;

    // ---- image * mask (extractor.jl:66-71) ---------------------------------------------------
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
        // binary mask on the halo region; coordinates clamped to the image (replicate border of the blur)
        for (int i = tid; i < rh * rw; i += DET_THREADS) {
            const int yy = i % rh, xx = i / rh;
            const int Y = clampi(y0 - hw + yy, 0, H - 1) + 1, X = clampi(x0 - hw + xx, 0, W - 1) + 1;  // 1-based
            float m = 1.f;
            if (n_near <= MAX_NEAR) {
                for (int k = 0; k < n_near; ++k) {
                    const int dy = Y - s_near[2 * k], dx = X - s_near[2 * k + 1];
                    if (dy * dy + dx * dx <= r2) { m = 0.f; break; }
                }
            } else {
                for (int k = 0; k < a.n_cur; ++k) {
                    const int dy = Y - (int)rint(cur[2 * k]), dx = X - (int)rint(cur[2 * k + 1]);
                    if (dy * dy + dx * dx <= r2) { m = 0.f; break; }
                }
            }
            s_m0[i] = m;
        }
        __syncthreads();
        if (hw > 0) {
            // y pass -> s_tmp[h][rw]
            for (int i = tid; i < h * rw; i += DET_THREADS) {
                const int y = i % h, xx = i / h;
                double acc = 0.0;
                for (int t = 0; t <= 2 * hw; ++t) acc += a.kw[t] * (double)s_m0[(y + t) + xx * rh];
                s_tmp[i] = acc;
            }
            __syncthreads();
            for (int i = tid; i < h * w; i += DET_THREADS) {
                const int y = i % h, x = i / h;
                double acc = 0.0;
                for (int t = 0; t <= 2 * hw; ++t) acc += a.kw[t] * s_tmp[y + (x + t) * h];
                s_img[i] = img[(size_t)(y0 + y) + (size_t)(x0 + x) * H] * acc;
            }
        } else {
            for (int i = tid; i < h * w; i += DET_THREADS) {
                const int y = i % h, x = i / h;
                s_img[i] = img[(size_t)(y0 + y) + (size_t)(x0 + x) * H] * (double)s_m0[i];
            }
        }
    } else {
        for (int i = tid; i < h * w; i += DET_THREADS) {
            const int y = i % h, x = i / h;
            s_img[i] = img[(size_t)(y0 + y) + (size_t)(x0 + x) * H];
        }
    }
    __syncthreads();

    // ---- Shi-Tomasi response on the cell sub-image (replicate border at the cell edge) --------
#define CP(yy, xx) s_img[clampi(yy, 0, h - 1) + clampi(xx, 0, w - 1) * h]
    for (int i = tid; i < h * w; i += DET_THREADS) {
        const int y = i % h, x = i / h;
        const double g_y = ((CP(y + 1, x - 1) - CP(y - 1, x - 1)) + 2.0 * (CP(y + 1, x) - CP(y - 1, x)) + (CP(y + 1, x + 1) - CP(y - 1, x + 1))) / 8.0;
        const double g_x = ((CP(y - 1, x + 1) - CP(y - 1, x - 1)) + 2.0 * (CP(y, x + 1) - CP(y, x - 1)) + (CP(y + 1, x + 1) - CP(y + 1, x - 1))) / 8.0;
        s_gyy[i] = g_y * g_y; s_gyx[i] = g_y * g_x; s_gxx[i] = g_x * g_x;
    }
#undef CP
    __syncthreads();
    for (int i = tid; i < h * w; i += DET_THREADS) {
        const int y = i % h, x = i / h;
        double sa = 0.0, sb = 0.0, sc = 0.0;
        for (int dx = -1; dx <= 1; ++dx)
            for (int dy = -1; dy <= 1; ++dy) {
                const int j = clampi(y + dy, 0, h - 1) + clampi(x + dx, 0, w - 1) * h;
                sa += s_gyy[j]; sb += s_gyx[j]; sc += s_gxx[j];
            }
        sa /= 9.0; sb /= 9.0; sc /= 9.0;
        s_R[i] = ((sa + sc) - sqrt((sa - sc) * (sa - sc) + 4.0 * sb * sb)) / 2.0;
    }
    __syncthreads();

    // ---- strict 3x3 local maxima, enumerated column-major (findlocalmaxima) -------------------
    int n_cand = 0;
    for (int base = 0; base < h * w; base += DET_THREADS) {
        const int i = base + tid;
        int ismax = 0;
        if (i < h * w) {
            const int y = i % h, x = i / h;
            const double r = s_R[i];
            ismax = 1;
            for (int dx = -1; dx <= 1; ++dx)
                for (int dy = -1; dy <= 1; ++dy) {
                    if (!dx && !dy) continue;
                    const int yy = y + dy, xx = x + dx;
                    if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
                    if (!(s_R[yy + xx * h] < r)) ismax = 0;
                }
        }
        int tot;
        const int pos = block_excl_scan(ismax, s_misc + 1, tot);
        if (ismax) { s_candr[n_cand + pos] = s_R[i]; s_candi[n_cand + pos] = i; }
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
            const int idx = s_candi[i];
            out[2 * (n_sel + pos)] = (int64_t)(idx % h) + 1 + y0;
            out[2 * (n_sel + pos) + 1] = (int64_t)(idx / h) + 1 + x0;
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

int launch_detect(cudaStream_t s, const DetArgs& a, const Hook* hk) {
    const int n_cells = a.grid_h * a.grid_w;
    const size_t smem = detect_smem_bytes(a.cs, a.hw);
    cudaFuncSetAttribute(k_detect_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(n_cells, a.n_frames);
    mark(hk, "k_detect_cells");
    k_detect_cells<<<grid, DET_THREADS, smem, s>>>(a);
    mark(hk, "k_detect_compact");
    k_detect_compact<<<a.n_frames, DET_THREADS, 0, s>>>(a, n_cells);
    return 2;
}

}  // namespace sk
