// C ABI of libslamklt.so (see include/slamklt.h).  Host-side plumbing only: contexts, device memory,
// copies, launch sequencing.  All compute is in pyramid.cu / lk.cu / detect.cu; there is no CPU fallback.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <thread>

#include <sched.h>
#include <string>
#include <utility>
#include <vector>

#include "common.cuh"

using namespace sk;

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) return fail(SLAMKLT_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define CKL()                                                                                             \
    do {                                                                                                  \
        cudaError_t e_ = cudaGetLastError();                                                              \
        if (e_ != cudaSuccess) return fail(SLAMKLT_E_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// ---------------------------------------------------------------------------------------------
// recursive Gaussian design (host, double).  [3P] ImageFiltering KernelFactors.IIRGaussian / TriggsSdika.
// ---------------------------------------------------------------------------------------------
namespace sk {

void iir_design(double sigma, double a[3], double* scale, double M[9]) {
    const double m0 = 1.16680, m1 = 1.10783, m2 = 1.40586;
    const double q = 1.31564 * (std::sqrt(1.0 + 0.490811 * sigma * sigma) - 1.0);
    const double as = (m0 + q) * (m1 * m1 + m2 * m2 + 2 * m1 * q + q * q);
    double B = m0 * (m1 * m1 + m2 * m2) / as;
    *scale = B * B;
    const double a1 = q * (2 * m0 * m1 + m1 * m1 + m2 * m2 + (2 * m0 + 4 * m1) * q + 3 * q * q) / as;
    const double a2 = -q * q * (m0 + 2 * m1 + 3 * q) / as;
    const double a3 = q * q * q / as;
    a[0] = a1; a[1] = a2; a[2] = a3;
    const double den = (1 + a1 - a2 + a3) * (1 - a1 - a2 - a3) * (1 + a2 + (a1 - a3) * a3);
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

static void mat3_mul(const double* A, const double* B, double* C) {
    double t[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += A[3 * i + k] * B[3 * k + j];
            t[3 * i + j] = s;
        }
    std::memcpy(C, t, sizeof(t));
}

static void mat3_pow(const double* A, int n, double* out) {
    double r[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, b[9];
    std::memcpy(b, A, sizeof(b));
    while (n > 0) {
        if (n & 1) mat3_mul(r, b, r);
        mat3_mul(b, b, b);
        n >>= 1;
    }
    std::memcpy(out, r, sizeof(r));
}

void iir_dev(double sigma, int K, int KRr, IirDev* o) {
    double a[3], sc, M[9];
    iir_design(sigma, a, &sc, M);
    o->a1 = (float)a[0]; o->a2 = (float)a[1]; o->a3 = (float)a[2];
    o->scale = (float)sc;
    o->inv1ma = (float)(1.0 / (1.0 - (a[0] + a[1] + a[2])));
    for (int i = 0; i < 9; ++i) o->M[i] = (float)M[i];
    const double A[9] = {a[0], a[1], a[2], 1, 0, 0, 0, 1, 0};  // companion matrix of the recursion
    double P[9];
    // the recursion is stable (spectral radius 0.20 at sigma 1, 0.69 at sigma 4), so A^(chunk * 2^j) underflows any fp32
    // significance after a few doublings: a Kogge-Stone step whose matrix is below 1e-13 cannot change the result
    auto fill = [&](int chunk, float (*dst)[9]) {
        int need = 0;
        for (int j = 0; j < 5; ++j) {
            mat3_pow(A, chunk << j, P);
            double mx = 0;
            for (int i = 0; i < 9; ++i) { dst[j][i] = (float)P[i]; mx = std::fmax(mx, std::fabs(P[i])); }
            if (mx > 1e-13) need = j + 1;
        }
        return need;
    };
    o->nsc = fill(K, o->P);
    o->nsr = fill(KRr, o->PR);
}

void iir_line_host(double* x, int n, double sigma, double iminus, double iplus) {
    double a[3], sc, M[9];
    iir_design(sigma, a, &sc, M);
    const double asum = a[0] + a[1] + a[2];
    const double um = iminus / (1.0 - asum);
    double u1 = um, u2 = um, u3 = um;
    for (int i = 0; i < n; ++i) {
        double u = x[i] + a[0] * u1 + a[1] * u2 + a[2] * u3;
        x[i] = u; u3 = u2; u2 = u1; u1 = u;
    }
    const double up = iplus / (1.0 - asum), vp = up / (1.0 - asum);
    const double d0 = x[n - 1] - up, d1 = x[n - 2] - up, d2 = x[n - 3] - up;
    double v1 = M[0] * d0 + M[1] * d1 + M[2] * d2 + vp;
    double v2 = M[3] * d0 + M[4] * d1 + M[5] * d2 + vp;
    double v3 = M[6] * d0 + M[7] * d1 + M[8] * d2 + vp;
    x[n - 1] = v1;
    for (int i = n - 2; i >= 0; --i) {
        double v = x[i] + a[0] * v1 + a[1] * v2 + a[2] * v3;
        x[i] = v; v3 = v2; v2 = v1; v1 = v;
    }
    for (int i = 0; i < n; ++i) x[i] *= sc;
}

}  // namespace sk

// ---------------------------------------------------------------------------------------------
// objects
// ---------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) return fail(SLAMKLT_E_CUDA, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct HostBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess) return fail(SLAMKLT_E_CUDA, "cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e));
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// ---------------------------------------------------------------------------------------------
// host worker pool (only used to repack Float64 host frames, see pack_u8_exact)
// ---------------------------------------------------------------------------------------------
class HostPool {
public:
    explicit HostPool(int n_threads) {
        for (int i = 0; i < n_threads; ++i) workers_.emplace_back([this] { run(); });
    }
    ~HostPool() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; ++epoch_; }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    int size() const { return (int)workers_.size() + 1; }
    // fn(i) for i in [0, n) on the workers; returns at once.  wait() joins in on what is left and returns when all items are done.
    void start(int n, std::function<void(int)> fn) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = std::move(fn); n_ = n; next_.store(0); left_.store(n);
            active_ = n > 0;
            ++epoch_;
        }
        cv_.notify_all();
    }
    void wait() {
        if (!active_) return;
        work();
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [this] { return left_.load() == 0; });
        active_ = false;
    }

private:
    void work() {
        for (;;) {
            const int i = next_.fetch_add(1);
            if (i >= n_) break;
            fn_(i);
            if (left_.fetch_sub(1) == 1) { std::lock_guard<std::mutex> lk(mu_); done_cv_.notify_all(); }
        }
    }
    void run() {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return epoch_ != seen; });
                seen = epoch_;
                if (stop_) return;
            }
            work();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    std::function<void(int)> fn_;
    int n_ = 0;
    bool active_ = false;
    std::atomic<int> next_{0}, left_{0};
    unsigned long long epoch_ = 0;
    bool stop_ = false;
};

// Lossless repacking of a Float64 frame that holds 8-bit data (every pixel exactly k/255, what Gray{Float64}.(load(png)) of the
// reference's example produces, example/kitty/main.jl:36-40): dst[i] = k.  Returns false at the first pixel that is not such a
// value; the caller then ships the frame as Float64.  The device reconstructs (double)k / 255.0, i.e. the identical Float64.
namespace sk { bool pack_u8_exact_impl(const double* src, uint8_t* dst, size_t n); }  // host_pack.cpp (AVX-512 / AVX2 / scalar)
static bool pack_u8_exact(const double* src, uint8_t* dst, size_t n) { return sk::pack_u8_exact_impl(src, dst, n); }

struct slamklt_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // H2D of the pipelined batch step
    cudaStream_t raw_stream = nullptr;   // H2D of the chunks that travel as plain Float64 while host threads repack the others
    double pack_Bps = 0.0, raw_Bps = 0.0;  // measured source bytes per second of the two upload engines (0 = not measured yet)
    double dbg_begin_ns = 0, dbg_pack_ns = 0, dbg_wait_ns = 0, dbg_end_ns = 0; long long dbg_steps = 0;  // SLAMKLT_VERBOSE: where a step's host time goes
    cudaStream_t d2h_stream = nullptr;   // D2H of the pipelined batch step
    PyrStreams pyr_streams{};            // build DAG: main + two side streams
    cudaStream_t lk_stream = nullptr;    // tracking of chunk k overlaps the build of chunk k+1
    std::vector<cudaEvent_t> pipe_ev;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_join = nullptr;
    std::mutex mu;
    unsigned long long* d_counters = nullptr;  // [0..1] executed window-iterations / iterations, then the work-counter ring
    unsigned work_idx = 0;
    uint64_t launches = 0, h2d = 0, d2h = 0;
    DevBuf staging, img64, pts, disp, outp, status, cell_out, cell_cnt, det_out, det_n, det_bin, cur, match, gtab;
    HostBuf h_out, h_status, h_misc, h_pack1;
    HostPool* pool = nullptr;  // created on first use
    std::map<std::pair<int, long long>, float*> norm_cache;  // (n, sigma bits) -> device 1/norm
    std::map<long long, double*> ytab_cache;                 // (first mask-blur tap bits) -> device table of tap-subset sums (detect)
    // per-kernel profiling (off by default)
    bool prof_on = false;
    Hook hook{nullptr, nullptr};
    std::vector<std::pair<std::string, cudaEvent_t>> prof_ev;
    std::vector<cudaEvent_t> ev_pool;
    std::map<std::string, std::pair<double, long long>> prof_acc;
    const Hook* hk() const { return prof_on ? &hook : nullptr; }
};

static int ensure_pool(slamklt_ctx* c) {
    if (!c->pool) {
        int avail = (int)std::thread::hardware_concurrency();
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof(set), &set) == 0) avail = CPU_COUNT(&set);
        const char* e = getenv("SLAMKLT_HOST_THREADS");
        int want = e ? atoi(e) : std::min(avail, 16);
        c->pool = new HostPool(std::max(want, 1) - 1);
    }
    return 0;
}

// work counters of the persistent tracking grid: a ring, one slot per launch (the launcher zeroes the slot on its stream)
static constexpr int WORK_RING = 1024;
static unsigned* work_slot(slamklt_ctx* c) { return reinterpret_cast<unsigned*>(c->d_counters + 2) + (c->work_idx++ % WORK_RING); }

static void prof_mark(void* user, const char* name) {
    slamklt_ctx* c = (slamklt_ctx*)user;
    cudaEvent_t e;
    if (!c->ev_pool.empty()) { e = c->ev_pool.back(); c->ev_pool.pop_back(); }
    else if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, c->stream);
    c->prof_ev.emplace_back(name, e);
}
static void prof_end(slamklt_ctx* c) { if (c->prof_on) prof_mark(c, "<end>"); }

struct slamklt_pyr {
    PyrGeom g;
    float* base = nullptr;
    void* d_maps = nullptr;  // device array of per-level tensor maps over `base` (lk_tma.cu); views use their batch's
    bool owns = false;
    bool built = false;
    int mode = 0;
    slamklt_batch* parent = nullptr;
    int logical_slot = 0;
};

struct slamklt_batch {
    PyrGeom g;
    float* base = nullptr;
    int n_frames = 0, n_slots = 0, slot0 = 0, max_pts = 0, n_pts = 0;
    int up_dtype = -1, up_ld = 0;
    void* d_maps = nullptr;  // device array of per-level tensor maps over the slot ring
    HostBuf h_pack;          // pinned staging of the 8-bit repacked host frames (slamklt_batch_step)
    TScratch ts{};           // scratch ring of the y-filtered product planes (kept in L2 by an access-policy window)
    int build_group = 0;     // frames per build group (0: whole batch at once, planes inside the frame blocks)
    DevBuf staging, staging8, img64, pts, outp, status, gtab;
    std::vector<slamklt_pyr*> views;
    bool primed = false;
    cudaEvent_t ev_lk_done = nullptr;  // last tracking kernel that read this batch's slots (recorded on the lk stream)
    bool lk_pending = false;
    cudaEvent_t ev_pack = nullptr;       // last copy out of h_pack queued by slamklt_batch_upload
    cudaEvent_t ev_step_done = nullptr;  // everything slamklt_batch_step_begin queued, result copies included
    bool step_pending = false;
    struct UpChunk { int f0, n, dtype; const void* ptr; };
    std::vector<UpChunk> up_chunks;      // up_dtype == -2 (a step that mixed repacked and plain chunks): where each chunk's frames sit
    bool quiesced = false;               // nothing of this batch is in flight on any stream (set by the calls that wait for it)
    std::vector<cudaEvent_t> raw_ev;     // timing events around the plain Float64 chunk copies of the last step (2 per chunk)
    std::vector<std::pair<int, size_t>> raw_meas;  // (chunk, bytes) of those copies: read back at the next step
};

// a batch whose last tracking kernel is still in flight on the lk stream must be waited for before the compute stream
// touches its slots, points or result buffers again
#define BATCH_WAIT_LK(c, b)                                                     \
    do {                                                                        \
        if ((b)->lk_pending) {                                                  \
            CK(cudaStreamWaitEvent((c)->stream, (b)->ev_lk_done, 0));           \
            (b)->lk_pending = false;                                            \
        }                                                                       \
    } while (0)

// LK reads up to 32 window columns without predicates; the last plane of an allocation needs that much slack
static size_t alloc_slack(const PyrGeom& g) { return (size_t)34 * std::max<size_t>(1100, (size_t)g.lv[0].pitch + 12); }

static int make_geom(int H, int W, int levels, PyrGeom* g) {
    if (H < 4 || W < 4) return fail(SLAMKLT_E_INVALID, "image %dx%d too small", H, W);
    if (levels < 0 || levels + 1 > MAX_LAYERS) return fail(SLAMKLT_E_INVALID, "pyramid_levels %d out of range [0,%d]", levels, MAX_LAYERS - 1);
    std::memset(g, 0, sizeof(*g));
    g->nl = levels + 1; g->H0 = H; g->W0 = W;
    size_t off = 0;
    int h = H, w = W;
    for (int l = 0; l < g->nl; ++l) {
        if (h < 4 || w < 4) return fail(SLAMKLT_E_INVALID, "level %d is %dx%d: recursive filter needs more than 3 samples per line", l, h, w);
        // levels with more than 1088 rows (level 0: 1280) or 2048 columns are built by the general kernels of pyramid.cu (level_tiled)
        if (h > 16384 || w > 16384) return fail(SLAMKLT_E_INVALID, "image %dx%d exceeds the supported 16384 x 16384", h, w);
        LevelGeom& L = g->lv[l];
        // one guard row and one guard column (kept zero) so that LK's weight-0 bilinear tap at H+1 / W+1 stays in bounds
        L.H = h; L.W = w; L.pitch = (h + 1 + 3) & ~3;
        L.plane_elems = (((size_t)L.pitch * (w + 1)) + 31) & ~(size_t)31;
        L.off = off;
        off += L.plane_elems * DP_COUNT;
        h = (h + 1) / 2; w = (w + 1) / 2;  // ceil(s/2), [3P] Images.gaussian_pyramid
    }
    g->frame_elems = off;
    g->pad_ = 0;
    return 0;
}

static float* pyr_frame_base(const slamklt_pyr* p) {
    if (p->parent) {
        const slamklt_batch* b = p->parent;
        return b->base + (size_t)((b->slot0 + p->logical_slot) % b->n_slots) * b->g.frame_elems;
    }
    return p->base;
}
static FrameSet fs_of(const slamklt_pyr* p) {
    if (p->parent) {
        const slamklt_batch* b = p->parent;
        return FrameSet{b->base, b->g.frame_elems, b->n_slots, (b->slot0 + p->logical_slot) % b->n_slots};
    }
    return FrameSet{p->base, p->g.frame_elems, 1, 0};
}
static const void* maps_of(const slamklt_pyr* p) { return p->parent ? p->parent->d_maps : p->d_maps; }

// Keep a scratch buffer resident in L2: persisting set-aside sized for it + an access-policy window on the streams whose kernels
// touch it (the build's main stream and the coarse-level side stream).  Best effort: without it the ring still works, its
// lines just compete with the streaming traffic.  SLAMKLT_NO_L2_WINDOW=1 turns it off.
static void set_l2_window(slamklt_ctx* c, void* base, size_t bytes) {
    if (getenv("SLAMKLT_NO_L2_WINDOW")) return;
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, c->device);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, c->device);
    if (max_persist <= 0 || max_window <= 0) return;
    const size_t set_aside = std::min((size_t)max_persist, bytes + (bytes >> 3));
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, set_aside) != cudaSuccess) { cudaGetLastError(); return; }
    cudaStreamAttrValue av{};
    av.accessPolicyWindow.base_ptr = base;
    av.accessPolicyWindow.num_bytes = std::min(bytes, (size_t)max_window);
    av.accessPolicyWindow.hitRatio = bytes <= set_aside ? 1.0f : (float)((double)set_aside / (double)bytes);
    av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaStream_t ss[] = {c->stream, c->pyr_streams.c};
    for (cudaStream_t st : ss)
        if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
    if (getenv("SLAMKLT_VERBOSE"))
        fprintf(stderr, "[slamklt] L2 window: %.1f MB ring, set-aside %.1f MB (max %.1f MB), window max %.1f MB\n", bytes / 1e6, set_aside / 1e6,
                max_persist / 1e6, max_window / 1e6);
}

// tensor maps of a frame ring for the TMA-staged tracking kernel, uploaded to 64-byte aligned device memory
static int make_maps(slamklt_ctx* c, const PyrGeom& g, float* base, int n_slots, void** d_maps) {
    std::vector<unsigned char> host(lk_tma_level_bytes() * MAX_LAYERS + 64);
    unsigned char* hp = (unsigned char*)(((uintptr_t)host.data() + 63) & ~(uintptr_t)63);
    char err[200];
    if (lk_tma_encode(g, base, n_slots, hp, err, sizeof(err))) return fail(SLAMKLT_E_CUDA, "%s", err);
    if (!*d_maps) CK(cudaMalloc(d_maps, lk_tma_level_bytes() * MAX_LAYERS));
    CK(cudaMemcpyAsync(*d_maps, hp, lk_tma_level_bytes() * MAX_LAYERS, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));  // `host` goes out of scope
    return 0;
}

// table of first-set-up structure tensors for the tracking kernel (SLAMKLT_LK_GTAB=0 computes them in the tracking warps)
static int set_gtab(DevBuf& buf, LKArgs* a) {
    static const bool on = !(getenv("SLAMKLT_LK_GTAB") && getenv("SLAMKLT_LK_GTAB")[0] == '0');
    a->gtab = nullptr; a->gtab_levels = 0;
    if (!on || !a->mapsA || !a->mapsB || 2 * a->window + 1 > 19) return 0;
    const int nlv = std::max(a->levels, a->mode == 2 ? a->levels3d : 0) + 1;
    int r = buf.ensure((size_t)a->n_frames * a->n_per_frame * nlv * 16);
    if (r) return r;
    a->gtab = buf.p; a->gtab_levels = nlv;
    return 0;
}
static FrameSet fs_of(const slamklt_batch* b) { return FrameSet{b->base, b->g.frame_elems, b->n_slots, b->slot0}; }

static int get_norms(slamklt_ctx* c, const PyrGeom& g, double sigma, const float** ny, const float** nx) {
    long long bits;
    std::memcpy(&bits, &sigma, sizeof(bits));
    for (int l = 0; l + 1 < g.nl; ++l) {
        for (int d = 0; d < 2; ++d) {
            const int n = d == 0 ? g.lv[l].H : g.lv[l].W;
            auto key = std::make_pair(n, bits);
            auto it = c->norm_cache.find(key);
            if (it == c->norm_cache.end()) {
                std::vector<double> ones(n, 1.0);
                iir_line_host(ones.data(), n, sigma, 0.0, 0.0);
                std::vector<float> inv(n);
                for (int i = 0; i < n; ++i) inv[i] = (float)(1.0 / ones[i]);
                float* dptr = nullptr;
                CK(cudaMalloc(&dptr, sizeof(float) * n));
                CK(cudaMemcpyAsync(dptr, inv.data(), sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
                CK(cudaStreamSynchronize(c->stream));
                it = c->norm_cache.emplace(key, dptr).first;
            }
            (d == 0 ? ny : nx)[l] = it->second;
        }
    }
    return 0;
}

static size_t dtype_size(int dtype) { return dtype == SLAMKLT_F64 ? 8 : dtype == SLAMKLT_F32 ? 4 : 1; }

// ---------------------------------------------------------------------------------------------
extern "C" {

const char* slamklt_last_error(void) { return g_err; }
int slamklt_version(void) { return SLAMKLT_VERSION; }

int slamklt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int slamklt_ctx_create(int device, slamklt_ctx** out) {
    if (!out) return fail(SLAMKLT_E_INVALID, "out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { cudaGetLastError(); return fail(SLAMKLT_E_NODEVICE, "no CUDA device visible (%s); libslamklt has no CPU fallback", e == cudaSuccess ? "count 0" : cudaGetErrorString(e)); }
    if (device < 0 || device >= n) return fail(SLAMKLT_E_INVALID, "device %d out of range [0,%d)", device, n);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(SLAMKLT_E_NODEVICE, "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
    CK(cudaSetDevice(device));
    slamklt_ctx* c = new slamklt_ctx();
    c->device = device;
    // stream priorities: pyramid-build streams above the tracking stream, so that the next batch's build kernels take the CTA
    // slots the tracking kernel frees at its tail first (measured: step 1.664 -> 1.654 ms; SLAMKLT_NO_PRIO=1 turns it off)
    int prio_lo = 0, prio_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    const bool prio = getenv("SLAMKLT_NO_PRIO") == nullptr;
    const int p_build = prio ? prio_hi : 0, p_lk = prio ? prio_lo : 0;
    CK(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, p_build));
    CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->raw_stream, cudaStreamNonBlocking));
    c->pyr_streams.main = c->stream;
    CK(cudaStreamCreateWithPriority(&c->lk_stream, cudaStreamNonBlocking, p_lk));
    CK(cudaStreamCreateWithPriority(&c->pyr_streams.b, cudaStreamNonBlocking, p_build));
    CK(cudaStreamCreateWithPriority(&c->pyr_streams.c, cudaStreamNonBlocking, p_build));
    for (auto& e : c->pyr_streams.ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->pyr_streams.parallel = getenv("SLAMKLT_SERIAL_BUILD") == nullptr;
    CK(cudaEventCreate(&c->ev0));
    CK(cudaEventCreate(&c->ev1));
    CK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    CK(cudaMalloc(&c->d_counters, 2 * sizeof(unsigned long long) + WORK_RING * sizeof(unsigned)));
    CK(cudaMemsetAsync(c->d_counters, 0, 2 * sizeof(unsigned long long) + WORK_RING * sizeof(unsigned), c->stream));
    *out = c;
    return 0;
}

int slamklt_ctx_destroy(slamklt_ctx* c) {
    if (!c) return 0;
    if (getenv("SLAMKLT_VERBOSE") && c->dbg_steps > 0)
        fprintf(stderr, "[slamklt] %lld steps: begin %.3f ms (of which waiting for the repack %.3f, repack busy %.3f), end %.3f ms per step\n", c->dbg_steps,
                c->dbg_begin_ns / c->dbg_steps * 1e-6, c->dbg_wait_ns / c->dbg_steps * 1e-6, c->dbg_pack_ns / c->dbg_steps * 1e-6, c->dbg_end_ns / c->dbg_steps * 1e-6);
    cudaSetDevice(c->device);
    // every stream of the context may still use the buffers freed below
    cudaStream_t all[] = {c->stream, c->lk_stream, c->copy_stream, c->raw_stream, c->d2h_stream, c->pyr_streams.b, c->pyr_streams.c};
    for (cudaStream_t st : all) if (st) cudaStreamSynchronize(st);
    for (auto& kv : c->norm_cache) cudaFree(kv.second);
    for (auto& kv : c->ytab_cache) cudaFree(kv.second);
    DevBuf* bufs[] = {&c->staging, &c->img64, &c->pts, &c->disp, &c->outp, &c->status, &c->cell_out, &c->cell_cnt, &c->det_out, &c->det_n, &c->det_bin, &c->cur, &c->match, &c->gtab};
    for (DevBuf* b : bufs) b->release();
    c->h_out.release(); c->h_status.release(); c->h_misc.release(); c->h_pack1.release();
    delete c->pool;
    for (auto& pe : c->prof_ev) cudaEventDestroy(pe.second);
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    cudaFree(c->d_counters);
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); cudaEventDestroy(c->ev_join);
    for (auto e : c->pipe_ev) cudaEventDestroy(e);
    cudaStreamDestroy(c->copy_stream);
    cudaStreamDestroy(c->raw_stream);
    cudaStreamDestroy(c->d2h_stream);
    for (auto& e : c->pyr_streams.ev) cudaEventDestroy(e);
    cudaStreamDestroy(c->lk_stream);
    cudaStreamDestroy(c->pyr_streams.b);
    cudaStreamDestroy(c->pyr_streams.c);
    cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

int slamklt_ctx_sync(slamklt_ctx* c) {
    if (!c) return fail(SLAMKLT_E_INVALID, "ctx is NULL");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaStreamSynchronize(c->lk_stream));
    CK(cudaStreamSynchronize(c->d2h_stream));
    return 0;
}

int slamklt_get_stats(slamklt_ctx* c, slamklt_stats* out, int reset) {
    if (!c || !out) return fail(SLAMKLT_E_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    unsigned long long h[2];
    CK(cudaMemcpyAsync(h, c->d_counters, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    out->kernel_launches = c->launches;
    out->lk_window_iters = h[0]; out->lk_iters = h[1];
    out->h2d_bytes = c->h2d; out->d2h_bytes = c->d2h;
    if (reset) CK(cudaMemsetAsync(c->d_counters, 0, sizeof(h), c->stream));
    return 0;
}

// measured rates of the two upload engines of slamklt_batch_step (source bytes per second; 0 = not measured yet)
int slamklt_upload_rates(slamklt_ctx* c, double* pack_Bps, double* raw_Bps) {
    if (!c || !pack_Bps || !raw_Bps) return fail(SLAMKLT_E_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(c->mu);
    *pack_Bps = c->pack_Bps; *raw_Bps = c->raw_Bps;
    return 0;
}

int slamklt_profile(slamklt_ctx* c, int enable) {
    if (!c) return fail(SLAMKLT_E_INVALID, "ctx is NULL");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    c->hook.fn = prof_mark; c->hook.user = c;
    c->prof_on = enable != 0;
    if (enable) { c->prof_acc.clear(); for (auto& pe : c->prof_ev) c->ev_pool.push_back(pe.second); c->prof_ev.clear(); }
    return 0;
}

int slamklt_profile_report(slamklt_ctx* c, char* buf, size_t cap) {
    if (!c || !buf || cap == 0) return fail(SLAMKLT_E_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i + 1 < c->prof_ev.size(); ++i) {
        if (c->prof_ev[i].first == "<end>") continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->prof_ev[i].second, c->prof_ev[i + 1].second) != cudaSuccess) { cudaGetLastError(); continue; }
        auto& acc = c->prof_acc[c->prof_ev[i].first];
        acc.first += ms; acc.second += 1;
    }
    for (auto& pe : c->prof_ev) c->ev_pool.push_back(pe.second);
    c->prof_ev.clear();
    size_t off = 0;
    buf[0] = 0;
    for (auto& kv : c->prof_acc) {
        int n = snprintf(buf + off, cap - off, "%s %lld %.6f\n", kv.first.c_str(), kv.second.second, kv.second.first);
        if (n < 0 || (size_t)n >= cap - off) break;
        off += (size_t)n;
    }
    return 0;
}

int slamklt_timer_start(slamklt_ctx* c) {
    if (!c) return fail(SLAMKLT_E_INVALID, "ctx is NULL");
    CK(cudaSetDevice(c->device));
    CK(cudaEventRecord(c->ev0, c->stream));
    return 0;
}

int slamklt_timer_stop(slamklt_ctx* c, float* ms) {
    if (!c || !ms) return fail(SLAMKLT_E_INVALID, "NULL argument");
    CK(cudaSetDevice(c->device));
    // tracking kernels may run on the side stream: the stopwatch ends when both streams are done
    CK(cudaEventRecord(c->ev_join, c->lk_stream));
    CK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaEventSynchronize(c->ev1));
    CK(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return 0;
}

// ---- pyramids --------------------------------------------------------------------------------
int slamklt_pyr_create(slamklt_ctx* c, int H, int W, int levels, slamklt_pyr** out) {
    if (!c || !out) return fail(SLAMKLT_E_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    PyrGeom g;
    int r = make_geom(H, W, levels, &g);
    if (r) return r;
    slamklt_pyr* p = new slamklt_pyr();
    p->g = g;
    cudaError_t e = cudaMalloc(&p->base, (g.frame_elems + alloc_slack(g)) * sizeof(float));
    if (e != cudaSuccess) { delete p; return fail(SLAMKLT_E_CUDA, "cudaMalloc pyramid failed: %s", cudaGetErrorString(e)); }
    cudaMemsetAsync(p->base, 0, (g.frame_elems + alloc_slack(g)) * sizeof(float), c->stream);
    p->owns = true;
    r = make_maps(c, g, p->base, 1, &p->d_maps);
    if (r) { cudaFree(p->base); delete p; return r; }
    *out = p;
    return 0;
}

int slamklt_pyr_destroy(slamklt_ctx* c, slamklt_pyr* p) {
    if (!p) return 0;
    if (!c) return fail(SLAMKLT_E_INVALID, "ctx is NULL");
    if (p->parent) return fail(SLAMKLT_E_INVALID, "pyramid is a batch slot view; destroy the batch instead");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    if (p->owns && p->base) cudaFree(p->base);
    if (p->d_maps) cudaFree(p->d_maps);
    delete p;
    return 0;
}

static int upload_frames(slamklt_ctx* c, DevBuf& staging, const void* img, int dtype, int ld, size_t frame_stride_bytes, int n_frames, int H, int W) {
    const size_t es = dtype_size(dtype);
    int r = staging.ensure((size_t)n_frames * H * W * es);
    if (r) return r;
    if (ld == H && (n_frames == 1 || frame_stride_bytes == (size_t)H * W * es)) {
        CK(cudaMemcpyAsync(staging.p, img, (size_t)n_frames * H * W * es, cudaMemcpyHostToDevice, c->stream));
    } else {
        for (int f = 0; f < n_frames; ++f)
            CK(cudaMemcpy2DAsync((char*)staging.p + (size_t)f * H * W * es, (size_t)H * es, (const char*)img + (size_t)f * frame_stride_bytes,
                                 (size_t)ld * es, (size_t)H * es, W, cudaMemcpyHostToDevice, c->stream));
    }
    c->h2d += (uint64_t)n_frames * H * W * es;
    return 0;
}

static int build_frames(slamklt_ctx* c, FrameSet fs, int f0, int n_frames, const PyrGeom& g, const void* staged, int dtype, double sigma, int mode,
                        double* img64, const TScratch* ts = nullptr, int group = 0) {
    if (!(sigma > 0)) return fail(SLAMKLT_E_INVALID, "sigma must be positive");
    const float* ny[MAX_LAYERS] = {nullptr};
    const float* nx[MAX_LAYERS] = {nullptr};
    if (mode == SLAMKLT_MODE_CTOR) {
        int r = get_norms(c, g, sigma, ny, nx);
        if (r) return r;
    }
    if (img64) {  // a Float64 copy is wanted as well (batched detect on non-Float64 frames): separate conversion pass
        c->launches += launch_convert(c->stream, staged, dtype, g.H0, (size_t)g.H0 * g.W0, fs, f0, n_frames, g, img64, c->hk());
        CKL();
    }
    // level 0 is converted on the fly by the fused column kernel (no separate conversion pass)
    if (ts && ts->base && group > 0) {
        // groups of a few frames: the scratch ring is rewritten by every group and stays in L2; the side streams rejoin the
        // main stream after the last group only, so the small coarse-level kernels of one group overlap the next group
        const size_t fbytes = (size_t)g.H0 * g.W0 * dtype_size(dtype);
        for (int g0 = 0; g0 < n_frames; g0 += group) {
            const int n = std::min(group, n_frames - g0);
            const void* raw = staged ? (const char*)staged + (size_t)g0 * fbytes : nullptr;
            c->launches += launch_pyramid(c->pyr_streams, fs, f0 + g0, n, g, sigma, mode, ny, nx, raw, dtype, c->hk(), ts, g0 + n >= n_frames);
            CKL();
        }
    } else {
        c->launches += launch_pyramid(c->pyr_streams, fs, f0, n_frames, g, sigma, mode, ny, nx, staged, dtype, c->hk());
        CKL();
    }
    prof_end(c);
    return 0;
}

// One Float64 frame that holds 8-bit data (Gray{Float64}.(load(png)), example/kitty/main.jl:36-40) is repacked to one byte per pixel by
// the worker pool and copied to c->staging as UInt8: lossless (the device rebuilds the identical Float64), and a pageable 3.7 MB frame
// otherwise goes through the driver's staged copy.  Measured on the bench host: update! of a pageable KITTI frame 0.30 -> 0.15 ms.
// *packed = false: nothing was copied (not Float64, too small, not 8-bit data, or switched off) and the caller uploads as usual.
static int upload_one_repacked(slamklt_ctx* c, const void* img, int dtype, int ld, int H, int W, bool* packed) {
    *packed = false;
    const size_t npx = (size_t)H * W;
    static const bool no_pack1 = getenv("SLAMKLT_NO_PACK1") != nullptr;
    if (dtype != SLAMKLT_F64 || ld != H || npx < (1u << 17) || no_pack1 || getenv("SLAMKLT_NO_PACK") != nullptr) return 0;
    int r;
    if ((r = c->h_pack1.ensure(npx))) return r;
    if ((r = c->staging.ensure(npx * 8))) return r;
    if ((r = ensure_pool(c))) return r;
    std::atomic<int> bad{0};
    uint8_t* hp = (uint8_t*)c->h_pack1.p;   // (free again: every earlier user of this buffer waited for its copy)
    const int cols = 16, blocks = (W + cols - 1) / cols;
    c->pool->start(blocks, [=, &bad](int item) {
        const int x0 = item * cols, x1 = std::min(W, x0 + cols);
        if (!pack_u8_exact((const double*)img + (size_t)x0 * H, hp + (size_t)x0 * H, (size_t)(x1 - x0) * H)) bad.store(1);
    });
    c->pool->wait();
    if (bad.load() != 0) return 0;
    CK(cudaMemcpyAsync(c->staging.p, hp, npx, cudaMemcpyHostToDevice, c->stream));
    c->h2d += npx;
    *packed = true;
    return 0;
}

int slamklt_pyr_build(slamklt_ctx* c, slamklt_pyr* p, const void* img, int dtype, int ld, double sigma, int mode) {
    if (!c || !p || !img) return fail(SLAMKLT_E_INVALID, "NULL argument");
    if (dtype < 0 || dtype > 2) return fail(SLAMKLT_E_INVALID, "unknown dtype %d", dtype);
    if (mode != SLAMKLT_MODE_UPDATE && mode != SLAMKLT_MODE_CTOR) return fail(SLAMKLT_E_INVALID, "unknown mode %d", mode);
    if (ld < p->g.H0) return fail(SLAMKLT_E_INVALID, "ld %d < H %d", ld, p->g.H0);
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    int r;
    const int H = p->g.H0, W = p->g.W0;
    bool packed = false;
    if ((r = upload_one_repacked(c, img, dtype, ld, H, W, &packed))) return r;
    if (packed) dtype = SLAMKLT_U8;
    if (!packed && (r = upload_frames(c, c->staging, img, dtype, ld, 0, 1, H, W))) return r;
    r = build_frames(c, fs_of(p), 0, 1, p->g, c->staging.p, dtype, sigma, mode, nullptr);
    if (r) return r;
    CK(cudaStreamSynchronize(c->stream));  // the caller may free img right after the call
    p->built = true; p->mode = mode;
    return 0;
}

int slamklt_pyr_copy(slamklt_ctx* c, slamklt_pyr* dst, const slamklt_pyr* src) {
    if (!c || !dst || !src) return fail(SLAMKLT_E_INVALID, "NULL argument");
    if (dst->g.H0 != src->g.H0 || dst->g.W0 != src->g.W0 || dst->g.nl != src->g.nl) return fail(SLAMKLT_E_INVALID, "pyramid shapes differ");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    // batch slot views keep no base pointer of their own, and their batch may still be read by a tracking kernel on the lk stream
    if (src->parent) BATCH_WAIT_LK(c, src->parent);
    if (dst->parent) BATCH_WAIT_LK(c, dst->parent);
    CK(cudaMemcpyAsync(pyr_frame_base(dst), pyr_frame_base(src), src->g.frame_elems * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    dst->built = src->built; dst->mode = src->mode;
    return 0;
}

int slamklt_pyr_clone(slamklt_ctx* c, const slamklt_pyr* src, slamklt_pyr** out) {
    if (!c || !src || !out) return fail(SLAMKLT_E_INVALID, "NULL argument");
    slamklt_pyr* p = nullptr;
    int r = slamklt_pyr_create(c, src->g.H0, src->g.W0, src->g.nl - 1, &p);
    if (r) return r;
    r = slamklt_pyr_copy(c, p, src);
    if (r) { slamklt_pyr_destroy(c, p); return r; }
    r = slamklt_ctx_sync(c);
    if (r) return r;
    *out = p;
    return 0;
}

int slamklt_pyr_swap(slamklt_ctx* c, slamklt_pyr* a, slamklt_pyr* b) {
    if (!c || !a || !b) return fail(SLAMKLT_E_INVALID, "NULL argument");
    if (a->parent || b->parent) return fail(SLAMKLT_E_INVALID, "cannot swap batch slot views");
    if (a->g.H0 != b->g.H0 || a->g.W0 != b->g.W0 || a->g.nl != b->g.nl) return fail(SLAMKLT_E_INVALID, "pyramid shapes differ");
    std::lock_guard<std::mutex> lk(c->mu);
    std::swap(a->base, b->base);
    std::swap(a->d_maps, b->d_maps);
    std::swap(a->built, b->built);
    std::swap(a->mode, b->mode);
    std::swap(a->owns, b->owns);
    return 0;
}

int slamklt_pyr_info(const slamklt_pyr* p, int* H, int* W, int* levels, int* built) {
    if (!p) return fail(SLAMKLT_E_INVALID, "pyr is NULL");
    if (H) *H = p->g.H0;
    if (W) *W = p->g.W0;
    if (levels) *levels = p->g.nl - 1;
    if (built) *built = p->built ? 1 : 0;
    return 0;
}

int slamklt_pyr_level_dims(const slamklt_pyr* p, int level, int* H, int* W) {
    if (!p) return fail(SLAMKLT_E_INVALID, "pyr is NULL");
    if (level < 0 || level >= p->g.nl) return fail(SLAMKLT_E_INVALID, "level %d out of range", level);
    if (H) *H = p->g.lv[level].H;
    if (W) *W = p->g.lv[level].W;
    return 0;
}

int slamklt_pyr_download(slamklt_ctx* c, const slamklt_pyr* p, int level, int plane, double* out) {
    if (!c || !p || !out) return fail(SLAMKLT_E_INVALID, "NULL argument");
    if (level < 0 || level >= p->g.nl) return fail(SLAMKLT_E_INVALID, "level %d out of range", level);
    int dp = DP_I, which = -1, comp = -1;  // which: smoothed product plane to recompute; comp: component of the gradient plane
    bool sat = false, prefix = false;  // prefix: raw device row-prefix plane, W + 1 columns
    switch (plane) {
        case SLAMKLT_PLANE_LAYER: dp = DP_I; break;
        case SLAMKLT_PLANE_IY: dp = DP_GRAD; comp = 0; break;
        case SLAMKLT_PLANE_IX: dp = DP_GRAD; comp = 1; break;
        case SLAMKLT_PLANE_IYY: which = 0; sat = true; break;
        case SLAMKLT_PLANE_IXX: which = 1; sat = true; break;
        case SLAMKLT_PLANE_IYX: which = 2; sat = true; break;
        case SLAMKLT_PLANE_SYY: which = 0; break;
        case SLAMKLT_PLANE_SXX: which = 1; break;
        case SLAMKLT_PLANE_SYX: which = 2; break;
        case SLAMKLT_PLANE_BLUR: dp = DP_BLUR; break;
        case SLAMKLT_PLANE_RYY: dp = DP_RYY; prefix = true; break;
        case SLAMKLT_PLANE_RXX: dp = DP_RXX; prefix = true; break;
        case SLAMKLT_PLANE_RYX: dp = DP_RYX; prefix = true; break;
        default: return fail(SLAMKLT_E_INVALID, "unknown plane %d", plane);
    }
    if ((which >= 0 || comp >= 0 || prefix) && !p->built) return fail(SLAMKLT_E_INVALID, "pyramid has no gradients (not built)");
    if (plane == SLAMKLT_PLANE_BLUR && p->parent && p->parent->n_frames > 1 && getenv("SLAMKLT_NO_FUSED_RESIZE") == nullptr)
        return fail(SLAMKLT_E_INVALID, "the blurred planes of a batch slot are not retained (blur and decimation are one kernel there); "
                                       "build a standalone pyramid to inspect them");
    if (which >= 0 && p->parent && p->parent->ts.base)
        return fail(SLAMKLT_E_INVALID, "the smoothed product planes of a batch slot are not retained (the batch builds them through a scratch ring); "
                                       "build a standalone pyramid to inspect them");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    if (p->parent) BATCH_WAIT_LK(c, p->parent);
    const LevelGeom& L = p->g.lv[level];
    const int ncol = L.W + (prefix ? 1 : 0);
    std::vector<float> tmp((size_t)L.H * ncol * (comp >= 0 ? 2 : 1));
    if (which >= 0) {
        // the pyramid keeps only the row-prefix form of the smoothed planes; rebuild the plain plane from the y-filtered scratch
        c->launches += launch_smoothed_plane(c->stream, fs_of(p), 0, p->g, level, which, c->hk());
        CKL();
        prof_end(c);
        dp = DP_TMP;
    }
    const float* src = pyr_frame_base(p) + plane_off(L, dp);
    const size_t rowb = (size_t)L.H * sizeof(float) * (comp >= 0 ? 2 : 1), pitchb = (size_t)L.pitch * sizeof(float) * (comp >= 0 ? 2 : 1);
    CK(cudaMemcpy2DAsync(tmp.data(), rowb, src, pitchb, rowb, ncol, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->d2h += tmp.size() * sizeof(float);
    const size_t npx = (size_t)L.H * ncol;
    if (comp >= 0) for (size_t i = 0; i < npx; ++i) out[i] = (double)tmp[2 * i + comp];
    else for (size_t i = 0; i < npx; ++i) out[i] = (double)tmp[i];
    if (sat) {  // integral_image!, lucas_kanade.jl:131-138, in Float64 on the host (parity access only)
        for (int x = 0; x < L.W; ++x) {
            double acc = 0;
            for (int y = 0; y < L.H; ++y) { acc += out[y + (size_t)x * L.H]; out[y + (size_t)x * L.H] = acc; }
        }
        for (int x = 1; x < L.W; ++x)
            for (int y = 0; y < L.H; ++y) out[y + (size_t)x * L.H] += out[y + (size_t)(x - 1) * L.H];
    }
    return 0;
}

// ---- Lucas-Kanade ----------------------------------------------------------------------------
static int fill_lk_levels(const PyrGeom& g, LKArgs* a) {
    a->nl = g.nl;
    for (int l = 0; l < g.nl; ++l) {
        const LevelGeom& L = g.lv[l];
        LKLevel& d = a->lv[l];
        d.H = L.H; d.W = L.W; d.pitch = L.pitch;
        d.oI = plane_off(L, DP_I); d.oG = plane_off(L, DP_GRAD);
        d.oRyy = plane_off(L, DP_RYY); d.oRxx = plane_off(L, DP_RXX); d.oRyx = plane_off(L, DP_RYX);
    }
    return 0;
}

static int check_lk(const slamklt_lk_params* p, int nlA, int nlB) {
    if (!p) return fail(SLAMKLT_E_INVALID, "params is NULL");
    // (up to 9: TMA-staged kernel; up to 11: cp.async patch kernel; up to 15: row-per-lane kernel; beyond: any-window kernel)
    if (p->window_size < 1 || p->window_size > 255) return fail(SLAMKLT_E_INVALID, "window_size %d outside [1,255]", p->window_size);
    if (p->iterations < 0) return fail(SLAMKLT_E_INVALID, "iterations < 0");
    if (p->pyramid_levels < 0) return fail(SLAMKLT_E_INVALID, "pyramid_levels < 0");
    if (!(nlA > p->pyramid_levels && nlB > p->pyramid_levels)) return fail(SLAMKLT_E_LAYERS, "Not enough layers in pyramids.");
    return 0;
}

static int run_lk_single(slamklt_ctx* c, const slamklt_pyr* A, const slamklt_pyr* B, const double* pts, const double* disp_in, int n,
                         const slamklt_lk_params* p, int mode) {
    if (A->g.H0 != B->g.H0 || A->g.W0 != B->g.W0) return fail(SLAMKLT_E_INVALID, "pyramid shapes differ");
    if (!A->built || !B->built) return fail(SLAMKLT_E_INVALID, "pyramid has no gradients (not built)");
    if (A->parent) BATCH_WAIT_LK(c, A->parent);
    if (B->parent) BATCH_WAIT_LK(c, B->parent);
    int r;
    if ((r = c->pts.ensure((size_t)n * 16))) return r;
    if ((r = c->disp.ensure((size_t)n * 16))) return r;
    if ((r = c->outp.ensure((size_t)n * 16))) return r;
    if ((r = c->status.ensure((size_t)n))) return r;
    CK(cudaMemcpyAsync(c->pts.p, pts, (size_t)n * 16, cudaMemcpyHostToDevice, c->stream));
    c->h2d += (uint64_t)n * 16;
    if (disp_in) {
        CK(cudaMemcpyAsync(c->disp.p, disp_in, (size_t)n * 16, cudaMemcpyHostToDevice, c->stream));
        c->h2d += (uint64_t)n * 16;
    }
    LKArgs a{};
    a.A = fs_of(A); a.B = fs_of(B); a.offA = 0; a.offB = 0;
    fill_lk_levels(A->g, &a);
    a.mode = mode;
    a.pts = (const double*)c->pts.p;
    a.disp_in = disp_in ? (const double*)c->disp.p : nullptr;
    a.disp_out = (double*)c->disp.p;
    a.out_pts = (double*)c->outp.p;
    a.status = (uint8_t*)c->status.p;
    a.n_per_frame = n; a.n_frames = 1;
    a.iterations = p->iterations; a.window = p->window_size; a.levels = p->pyramid_levels;
    a.eig_thr = p->eigenvalue_threshold; a.eps = p->epsilon; a.max_dist = p->max_distance;
    a.counters = c->d_counters; a.work = work_slot(c);
    a.mapsA = maps_of(A); a.mapsB = maps_of(B);
    if ((r = set_gtab(c->gtab, &a))) return r;
    c->launches += launch_lk(c->stream, a, c->hk());
    CKL();
    prof_end(c);
    return 0;
}

int slamklt_optflow(slamklt_ctx* c, const slamklt_pyr* A, const slamklt_pyr* B, const double* pts, double* disp, int n,
                    const slamklt_lk_params* p, uint8_t* status, int* n_good) {
    if (!c || !A || !B) return fail(SLAMKLT_E_INVALID, "NULL argument");
    int r = check_lk(p, A->g.nl, B->g.nl);
    if (r) return r;
    if (n < 0) return fail(SLAMKLT_E_INVALID, "n < 0");
    if (n == 0) { if (n_good) *n_good = 0; return 0; }
    if (!pts || !disp || !status) return fail(SLAMKLT_E_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    r = run_lk_single(c, A, B, pts, disp, n, p, 0);
    if (r) return r;
    CK(cudaMemcpyAsync(disp, c->disp.p, (size_t)n * 16, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(status, c->status.p, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->d2h += (uint64_t)n * 17;
    int good = 0;
    for (int i = 0; i < n; ++i) good += status[i] ? 1 : 0;
    if (n_good) *n_good = good;
    return 0;
}

int slamklt_fb_track(slamklt_ctx* c, const slamklt_pyr* A, const slamklt_pyr* B, const double* pts, const double* disp, int n,
                     const slamklt_lk_params* p, double* out_pts, uint8_t* status) {
    if (!c || !A || !B) return fail(SLAMKLT_E_INVALID, "NULL argument");
    int r = check_lk(p, A->g.nl, B->g.nl);
    if (r) return r;
    if (n < 0) return fail(SLAMKLT_E_INVALID, "n < 0");
    if (n == 0) return 0;  // isempty(keypoints) && return, tracker.jl:24
    if (!pts || !out_pts || !status) return fail(SLAMKLT_E_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    r = run_lk_single(c, A, B, pts, disp, n, p, 1);
    if (r) return r;
    if ((r = c->h_out.ensure((size_t)n * 16))) return r;
    CK(cudaMemcpyAsync(c->h_out.p, c->outp.p, (size_t)n * 16, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(status, c->status.p, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->d2h += (uint64_t)n * 17;
    const double* h = (const double*)c->h_out.p;
    for (int i = 0; i < n; ++i)
        if (status[i] & 2) { out_pts[2 * i] = h[2 * i]; out_pts[2 * i + 1] = h[2 * i + 1]; }  // tracker.jl:43 writes only where forward ok
    return 0;
}

int slamklt_flow_matching(slamklt_ctx* c, const slamklt_pyr* A, const slamklt_pyr* B, const double* pts, const double* prior,
                          const uint8_t* has_prior, int n, const slamklt_lk_params* p, int levels_3d, double* out_pts, uint8_t* status) {
    if (!c || !A || !B) return fail(SLAMKLT_E_INVALID, "NULL argument");
    int r = check_lk(p, A->g.nl, B->g.nl);
    if (r) return r;
    if (levels_3d < 0 || !(A->g.nl > levels_3d && B->g.nl > levels_3d)) return fail(SLAMKLT_E_LAYERS, "Not enough layers in pyramids.");
    if (n < 0) return fail(SLAMKLT_E_INVALID, "n < 0");
    if (n == 0) return 0;
    if (!pts || !prior || !has_prior || !out_pts || !status) return fail(SLAMKLT_E_INVALID, "NULL argument");
    if (A->g.H0 != B->g.H0 || A->g.W0 != B->g.W0) return fail(SLAMKLT_E_INVALID, "pyramid shapes differ");
    if (!A->built || !B->built) return fail(SLAMKLT_E_INVALID, "pyramid has no gradients (not built)");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    if (A->parent) BATCH_WAIT_LK(c, A->parent);
    if (B->parent) BATCH_WAIT_LK(c, B->parent);
    if ((r = c->pts.ensure((size_t)n * 16))) return r;
    if ((r = c->disp.ensure((size_t)n * 16))) return r;
    if ((r = c->outp.ensure((size_t)n * 16))) return r;
    if ((r = c->status.ensure((size_t)n))) return r;
    if ((r = c->cur.ensure((size_t)n))) return r;
    CK(cudaMemcpyAsync(c->pts.p, pts, (size_t)n * 16, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->disp.p, prior, (size_t)n * 16, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->cur.p, has_prior, (size_t)n, cudaMemcpyHostToDevice, c->stream));
    c->h2d += (uint64_t)n * 33;
    LKArgs a{};
    a.A = fs_of(A); a.B = fs_of(B); a.offA = 0; a.offB = 0;
    fill_lk_levels(A->g, &a);
    a.mode = 2;
    a.pts = (const double*)c->pts.p; a.disp_in = (const double*)c->disp.p; a.disp_out = nullptr;
    a.out_pts = (double*)c->outp.p; a.status = (uint8_t*)c->status.p;
    a.has_prior = (const uint8_t*)c->cur.p; a.levels3d = levels_3d;
    a.n_per_frame = n; a.n_frames = 1;
    a.iterations = p->iterations; a.window = p->window_size; a.levels = p->pyramid_levels;
    a.eig_thr = p->eigenvalue_threshold; a.eps = p->epsilon; a.max_dist = p->max_distance;
    a.counters = c->d_counters; a.work = work_slot(c);
    a.mapsA = maps_of(A); a.mapsB = maps_of(B);
    if ((r = set_gtab(c->gtab, &a))) return r;
    c->launches += launch_lk(c->stream, a, c->hk());
    CKL();
    prof_end(c);
    if ((r = c->h_out.ensure((size_t)n * 16))) return r;
    CK(cudaMemcpyAsync(c->h_out.p, c->outp.p, (size_t)n * 16, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(status, c->status.p, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->d2h += (uint64_t)n * 17;
    const double* h = (const double*)c->h_out.p;
    for (int i = 0; i < n; ++i)
        if (status[i] & 2) { out_pts[2 * i] = h[2 * i]; out_pts[2 * i + 1] = h[2 * i + 1]; }
    return 0;
}

static MatchCam cam_dev(const slamklt_camera* c) { return MatchCam{c->fx, c->fy, c->cx, c->cy, c->k1, c->k2, c->p1, c->p2}; }

int slamklt_optical_flow_matching(slamklt_ctx* c, const slamklt_pyr* A, const slamklt_pyr* B, const double* pix, const uint8_t* is_3d,
                                  const double* world, const double* undist, int n, const double* cw, const slamklt_camera* cam,
                                  const slamklt_camera* rcam, const slamklt_matching_params* mp, double* out_pix, double* out_und,
                                  double* out_pos, uint8_t* status) {
    if (!c || !A || !B || !mp) return fail(SLAMKLT_E_INVALID, "NULL argument");
    const slamklt_lk_params* p = &mp->lk;
    int r = check_lk(p, A->g.nl, B->g.nl);
    if (r) return r;
    const int l3 = mp->pyramid_levels_3d;
    if (l3 < 0 || !(A->g.nl > l3 && B->g.nl > l3)) return fail(SLAMKLT_E_LAYERS, "Not enough layers in pyramids.");
    if (n < 0) return fail(SLAMKLT_E_INVALID, "n < 0");
    if (n == 0) return 0;  // map_manager.jl:536
    if (!pix || !is_3d || !world || !cw || !cam || !out_pix || !out_und || !out_pos || !status)
        return fail(SLAMKLT_E_INVALID, "NULL argument");
    if (mp->stereo && (!rcam || !undist)) return fail(SLAMKLT_E_INVALID, "stereo matching needs right_cam and undist_yx");
    if (A->g.H0 != B->g.H0 || A->g.W0 != B->g.W0) return fail(SLAMKLT_E_INVALID, "pyramid shapes differ");
    if (!A->built || !B->built) return fail(SLAMKLT_E_INVALID, "pyramid has no gradients (not built)");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    if (A->parent) BATCH_WAIT_LK(c, A->parent);
    if (B->parent) BATCH_WAIT_LK(c, B->parent);
    // device block: [pix 2n | world 3n | undist 2n | disp 2n | tracked 2n | out_pix 2n | out_und 2n | out_pos 3n] doubles, then
    // [is_3d n | flag n | status n] bytes
    const size_t N = (size_t)n;
    // sections start at multiples of 16 bytes (the tracking kernel reads points and priors as double2): M = N rounded up to even
    const size_t M = (N + 1) & ~(size_t)1;
    if ((r = c->match.ensure(M * 18 * 8 + N * 3))) return r;
    double* d = (double*)c->match.p;
    double *d_pix = d, *d_world = d + 2 * M, *d_undist = d + 5 * M, *d_disp = d + 7 * M, *d_trk = d + 9 * M, *d_opix = d + 11 * M,
           *d_ound = d + 13 * M, *d_opos = d + 15 * M;
    uint8_t* d_is3d = (uint8_t*)(d + 18 * M);
    uint8_t *d_flag = d_is3d + N, *d_status = d_flag + N;
    CK(cudaMemcpyAsync(d_pix, pix, N * 16, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_world, world, N * 24, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_is3d, is_3d, N, cudaMemcpyHostToDevice, c->stream));
    c->h2d += N * 41;
    if (mp->stereo) { CK(cudaMemcpyAsync(d_undist, undist, N * 16, cudaMemcpyHostToDevice, c->stream)); c->h2d += N * 16; }

    MatchArgs m{};
    m.n = n; m.stereo = mp->stereo ? 1 : 0;
    m.cam = cam_dev(cam);
    m.rcam = cam_dev(mp->stereo ? rcam : cam);
    if (mp->stereo) {
        // Ti0 * cw (frame.jl:464-468: `right_camera.Ti0 * cw * point` associates to the left), each element a left-to-right sum
        for (int j = 0; j < 4; ++j)
            for (int i = 0; i < 4; ++i) {
                volatile double acc = rcam->Ti0[i] * cw[4 * j];
                for (int k = 1; k < 4; ++k) { volatile double t = rcam->Ti0[i + 4 * k] * cw[k + 4 * j]; acc = acc + t; }
                m.T[i + 4 * j] = acc;
            }
        m.bound_h = (double)rcam->height; m.bound_w = (double)rcam->width;
    } else {
        for (int k = 0; k < 16; ++k) m.T[k] = cw[k];
        m.bound_h = (double)cam->height; m.bound_w = (double)cam->width;
    }
    m.scale = 1.0 / std::ldexp(1.0, l3);  // map_manager.jl:466
    m.epipolar = mp->epipolar_error;
    m.pix = d_pix; m.is_3d = d_is3d; m.world = d_world; m.undist = d_undist; m.flag = d_flag; m.disp = d_disp;
    m.tracked = d_trk; m.status = d_status; m.out_pix = d_opix; m.out_und = d_ound; m.out_pos = d_opos;
    mark(c->hk(), "k_match_prior");
    launch_match_prior(c->stream, m);

    LKArgs a{};
    a.A = fs_of(A); a.B = fs_of(B); a.offA = 0; a.offB = 0;
    fill_lk_levels(A->g, &a);
    a.mode = 2;
    a.pts = d_pix; a.disp_in = d_disp; a.disp_out = nullptr;
    a.out_pts = d_trk; a.status = d_status;
    a.has_prior = d_flag; a.levels3d = l3;
    a.n_per_frame = n; a.n_frames = 1;
    a.iterations = p->iterations; a.window = p->window_size; a.levels = p->pyramid_levels;
    a.eig_thr = p->eigenvalue_threshold; a.eps = p->epsilon; a.max_dist = p->max_distance;
    a.counters = c->d_counters; a.work = work_slot(c);
    a.mapsA = maps_of(A); a.mapsB = maps_of(B);
    if ((r = set_gtab(c->gtab, &a))) return r;
    c->launches += launch_lk(c->stream, a, c->hk());
    mark(c->hk(), "k_match_update");
    launch_match_update(c->stream, m);
    c->launches += 2;
    CKL();
    prof_end(c);
    CK(cudaMemcpyAsync(out_pix, d_opix, N * 16, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(out_und, d_ound, N * 16, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(out_pos, d_opos, N * 24, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(status, d_status, N, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->d2h += N * 57;
    return 0;
}

int slamklt_triangulate_stereo(slamklt_ctx* c, const double* und, const double* rund, int n, const slamklt_camera* cam,
                               const slamklt_camera* rcam, const double* wc, double max_error, double* out_world, uint8_t* status) {
    if (!c || !cam || !rcam || !wc) return fail(SLAMKLT_E_INVALID, "NULL argument");
    if (n < 0) return fail(SLAMKLT_E_INVALID, "n < 0");
    if (n == 0) return 0;  // "No stereo keypoints to triangulate", mapper.jl:146-149
    if (!und || !rund || !out_world || !status) return fail(SLAMKLT_E_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    const size_t N = (size_t)n;
    int r;
    if ((r = c->match.ensure(N * 7 * 8 + N))) return r;  // [und 2n | rund 2n | world 3n] doubles, then status bytes
    double* d = (double*)c->match.p;
    CK(cudaMemcpyAsync(d, und, N * 16, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d + 2 * N, rund, N * 16, cudaMemcpyHostToDevice, c->stream));
    c->h2d += N * 32;
    TriArgs a{};
    a.n = n; a.cam = cam_dev(cam); a.rcam = cam_dev(rcam);
    for (int k = 0; k < 16; ++k) { a.Ti0[k] = rcam->Ti0[k]; a.wc[k] = wc[k]; }
    a.max_error = max_error;
    a.und = d; a.rund = d + 2 * N; a.world = d + 4 * N; a.status = (uint8_t*)(d + 7 * N);
    mark(c->hk(), "k_triangulate_stereo");
    launch_triangulate_stereo(c->stream, a);
    c->launches += 1;
    CKL();
    prof_end(c);
    CK(cudaMemcpyAsync(out_world, a.world, N * 24, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(status, a.status, N, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->d2h += N * 25;
    return 0;
}

// ---- BRIEF describe + Hamming matching -----------------------------------------------
int slamklt_describe(slamklt_ctx* c, const void* img, int dtype, int H, int W, int ld, const int64_t* kps_yx, int n, const int32_t* pairs,
                     int n_bits, int window, double sigma, uint32_t* out_desc, uint8_t* out_valid) {
    if (!c || !img) return fail(SLAMKLT_E_INVALID, "NULL argument");
    if (dtype < 0 || dtype > 2) return fail(SLAMKLT_E_INVALID, "unknown dtype %d", dtype);
    if (H < 3 || W < 3 || ld < H || n < 0) return fail(SLAMKLT_E_INVALID, "bad shape");
    if (n_bits <= 0 || n_bits % 32 != 0 || n_bits > 1024) return fail(SLAMKLT_E_INVALID, "descriptor size %d must be a multiple of 32 (<= 1024)", n_bits);
    if (window < 1 || window > 9) return fail(SLAMKLT_E_INVALID, "window %d outside [1,9]", window);
    if (!(sigma > 0) || sigma > 4.0) return fail(SLAMKLT_E_INVALID, "sigma outside (0,4]");
    if (n == 0) return 0;
    if (!kps_yx || !pairs || !out_desc || !out_valid) return fail(SLAMKLT_E_INVALID, "NULL argument");
    const int lim = (window + 1) / 2;  // ceil(Int, window / 2)
    for (int b = 0; b < 4 * n_bits; ++b)
        if (pairs[b] < -lim || pairs[b] > lim) return fail(SLAMKLT_E_INVALID, "sampling offset %d outside the window", pairs[b]);
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    int r;
    if ((r = upload_frames(c, c->staging, img, dtype, ld, 0, 1, H, W))) return r;
    const double* d_img;
    if (dtype == SLAMKLT_F64) d_img = (const double*)c->staging.p;
    else {
        if ((r = c->img64.ensure((size_t)H * W * 8))) return r;
        PyrGeom g;
        if ((r = make_geom(H, W, 0, &g))) return r;
        if ((r = c->outp.ensure(g.frame_elems * sizeof(float)))) return r;
        FrameSet fs{(float*)c->outp.p, g.frame_elems, 1, 0};
        c->launches += launch_convert(c->stream, c->staging.p, dtype, H, (size_t)H * W, fs, 0, 1, g, (double*)c->img64.p, c->hk());
        CKL();
        d_img = (const double*)c->img64.p;
    }
    const size_t N = (size_t)n, words = (size_t)n_bits / 32;
    // device block: [kps 2n i64 | pairs 4*n_bits i32 | desc n*words u32 | valid n u8]
    const size_t o_pairs = N * 16, o_desc = o_pairs + (size_t)n_bits * 16, o_valid = o_desc + N * words * 4;
    if ((r = c->match.ensure(o_valid + N))) return r;
    char* d = (char*)c->match.p;
    CK(cudaMemcpyAsync(d, kps_yx, N * 16, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d + o_pairs, pairs, (size_t)n_bits * 16, cudaMemcpyHostToDevice, c->stream));
    c->h2d += N * 16 + (size_t)n_bits * 16;
    BriefArgs a{};
    a.img = d_img; a.H = H; a.W = W; a.ld = H; a.n = n;
    a.kps = (const long long*)d; a.pairs = (const int4*)(d + o_pairs);
    a.n_bits = n_bits; a.lim = lim;
    a.hw = 2 * (int)std::ceil(sigma);  // Kernel.gaussian(sigma): length 4*ceil(sigma)+1 [3P]
    double sum = 0;
    for (int i = -a.hw; i <= a.hw; ++i) { a.kw[i + a.hw] = std::exp(-(double)i * i / (2 * sigma * sigma)); sum += a.kw[i + a.hw]; }
    for (int i = 0; i <= 2 * a.hw; ++i) a.kw[i] /= sum;
    a.desc = (unsigned*)(d + o_desc); a.valid = (uint8_t*)(d + o_valid);
    mark(c->hk(), "k_brief");
    launch_brief(c->stream, a);
    c->launches += 1;
    CKL();
    prof_end(c);
    CK(cudaMemcpyAsync(out_desc, a.desc, N * words * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(out_valid, a.valid, N, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->d2h += N * (words * 4 + 1);
    return 0;
}

int slamklt_find_best_match(slamklt_ctx* c, const uint32_t* desc, int n_desc, int words, const int32_t* set_off, int n_sets,
                            const int32_t* target_set, const int32_t* cand_off, const int32_t* cand, int n_targets, int max_distance,
                            int32_t* best_pos, int32_t* best_dist, int32_t* second_dist) {
    if (!c) return fail(SLAMKLT_E_INVALID, "ctx is NULL");
    if (n_targets < 0 || n_desc < 0 || n_sets < 0 || words < 1) return fail(SLAMKLT_E_INVALID, "bad shape");
    if (n_targets == 0) return 0;
    if (!desc || !set_off || !target_set || !cand_off || !cand || !best_pos || !best_dist || !second_dist) return fail(SLAMKLT_E_INVALID, "NULL argument");
    if (set_off[0] != 0 || set_off[n_sets] != n_desc || cand_off[0] != 0) return fail(SLAMKLT_E_INVALID, "offset arrays must start at 0 and end at the row count");
    const int n_cand = cand_off[n_targets];
    for (int t = 0; t < n_targets; ++t)
        if (target_set[t] < 0 || target_set[t] >= n_sets || cand_off[t + 1] < cand_off[t]) return fail(SLAMKLT_E_INVALID, "bad target %d", t);
    for (int k = 0; k < n_cand; ++k)
        if (cand[k] < 0 || cand[k] >= n_sets) return fail(SLAMKLT_E_INVALID, "candidate set id out of range");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    const size_t b_desc = (size_t)n_desc * words * 4, b_set = ((size_t)n_sets + 1) * 4, b_t = (size_t)n_targets * 4, b_co = b_t + 4, b_c = (size_t)std::max(n_cand, 1) * 4;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~(size_t)15; return o; };
    const size_t o_desc = take(b_desc), o_set = take(b_set), o_ts = take(b_t), o_co = take(b_co), o_c = take(b_c), o_bp = take(b_t), o_bd = take(b_t), o_sd = take(b_t);
    int r;
    if ((r = c->match.ensure(off))) return r;
    char* d = (char*)c->match.p;
    CK(cudaMemcpyAsync(d + o_desc, desc, b_desc, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d + o_set, set_off, b_set, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d + o_ts, target_set, b_t, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d + o_co, cand_off, b_co, cudaMemcpyHostToDevice, c->stream));
    if (n_cand > 0) CK(cudaMemcpyAsync(d + o_c, cand, (size_t)n_cand * 4, cudaMemcpyHostToDevice, c->stream));
    c->h2d += b_desc + b_set + b_t + b_co + (size_t)n_cand * 4;
    HammingArgs a{};
    a.desc = (const unsigned*)(d + o_desc); a.set_off = (const int*)(d + o_set); a.words = words; a.n_targets = n_targets;
    a.max_distance = max_distance; a.target_set = (const int*)(d + o_ts); a.cand_off = (const int*)(d + o_co); a.cand = (const int*)(d + o_c);
    a.best_pos = (int*)(d + o_bp); a.best_dist = (int*)(d + o_bd); a.second_dist = (int*)(d + o_sd);
    mark(c->hk(), "k_best_match");
    launch_best_match(c->stream, a);
    c->launches += 1;
    CKL();
    prof_end(c);
    CK(cudaMemcpyAsync(best_pos, a.best_pos, b_t, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(best_dist, a.best_dist, b_t, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(second_dist, a.second_dist, b_t, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->d2h += 3 * b_t;
    return 0;
}

// ---- extractor -------------------------------------------------------------------------------
static int fill_det(const slamklt_detect_params* p, int H, int W, int n_cur, DetArgs* a) {
    if (!p) return fail(SLAMKLT_E_INVALID, "params is NULL");
    if (p->cell_size < 3 || p->cell_size > 64) return fail(SLAMKLT_E_INVALID, "cell_size %d outside [3,64]", p->cell_size);
    if (p->grid_h < 1 || p->grid_w < 1) return fail(SLAMKLT_E_INVALID, "grid_resolution must be positive");
    if (p->radius < 0) return fail(SLAMKLT_E_INVALID, "radius < 0");
    std::memset(a, 0, sizeof(*a));
    a->H = H; a->W = W; a->n_cur = n_cur;
    a->radius = p->radius; a->grid_h = p->grid_h; a->grid_w = p->grid_w; a->cs = p->cell_size;
    const int n_cells = p->grid_h * p->grid_w;
    const long long n_detect = (long long)p->max_points - n_cur;
    a->k_cell = (int)((n_detect + n_cells - 1) / n_cells);  // ceil(Int, n_detect / n_cells), extractor.jl:76
    a->slots = a->k_cell < p->cell_size * p->cell_size ? a->k_cell : p->cell_size * p->cell_size;
    if (a->slots < 1) a->slots = 1;
    a->min_resp = p->min_response;
    int hw = 0;
    if (n_cur > 0 && p->sigma_mask > 0 && !(std::fabs(p->sigma_mask) < 1e-12)) {
        hw = 2 * (int)std::ceil(p->sigma_mask);  // Kernel.gaussian(sigma): length 4*ceil(sigma)+1 [3P]
        if (hw > 16) return fail(SLAMKLT_E_INVALID, "sigma_mask %g too large (max 8)", p->sigma_mask);
        double s = 0;
        for (int i = -hw; i <= hw; ++i) { a->kw[i + hw] = std::exp(-(double)i * i / (2 * p->sigma_mask * p->sigma_mask)); s += a->kw[i + hw]; }
        for (int i = 0; i <= 2 * hw; ++i) a->kw[i] /= s;
    }
    a->hw = hw;
    if (detect_smem_bytes(a->cs, a->hw) > 220 * 1024) return fail(SLAMKLT_E_INVALID, "cell_size/sigma_mask need too much shared memory");
    return 0;
}

static int prep_detect(slamklt_ctx* c, DetArgs& a);

// src / src_dtype: the staged frames; Float64 always, other types only when the register-tiled kernel covers the request
static int run_detect(slamklt_ctx* c, DetArgs& a, const void* src, int src_dtype, int n_frames, const double* cur_host, int cap, int64_t* out_yx, int* n_out) {
    const int n_cells = a.grid_h * a.grid_w;
    int r;
    a.img = src_dtype == SLAMKLT_F64 ? (const double*)src : nullptr;
    a.src = src; a.src_dtype = src_dtype;
    a.n_frames = n_frames; a.cap = cap;
    if ((r = c->cell_out.ensure((size_t)n_frames * n_cells * a.slots * 16))) return r;
    if ((r = c->cell_cnt.ensure((size_t)n_frames * n_cells * 4))) return r;
    if ((r = c->det_out.ensure((size_t)n_frames * cap * 16 + 16))) return r;
    if ((r = c->det_n.ensure((size_t)n_frames * 4))) return r;
    if (a.n_cur > 0) {
        if ((r = c->cur.ensure((size_t)n_frames * a.n_cur * 16))) return r;
        CK(cudaMemcpyAsync(c->cur.p, cur_host, (size_t)n_frames * a.n_cur * 16, cudaMemcpyHostToDevice, c->stream));
        c->h2d += (uint64_t)n_frames * a.n_cur * 16;
        a.cur = (const double*)c->cur.p;
    }
    a.cell_out = (int64_t*)c->cell_out.p; a.cell_cnt = (int*)c->cell_cnt.p;
    a.out = (int64_t*)c->det_out.p; a.n_out = (int*)c->det_n.p;
    if (a.n_cur > 0 && detect2_supported(a)) {  // per (frame, cell row) lists of the current points within reach
        if ((r = c->det_bin.ensure((size_t)n_frames * a.grid_h * ((size_t)a.n_cur * 8 + 4)))) return r;
        a.bin_pts = (int2*)c->det_bin.p;
        a.bin_cnt = (int*)((char*)c->det_bin.p + (size_t)n_frames * a.grid_h * a.n_cur * 8);
    }
    c->launches += launch_detect(c->stream, a, c->hk());
    CKL();
    prof_end(c);
    // one copy for the counts and one for the whole [frame][cap][2] block (entries past n_out[f] are unspecified)
    CK(cudaMemcpyAsync(n_out, c->det_n.p, (size_t)n_frames * 4, cudaMemcpyDeviceToHost, c->stream));
    if (cap > 0) CK(cudaMemcpyAsync(out_yx, c->det_out.p, (size_t)n_frames * cap * 16, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->d2h += (uint64_t)n_frames * (4 + (uint64_t)cap * 16);
    return 0;
}

// Tables the register-tiled detect kernel needs (before detect2_supported is asked).
// ImageDraw's filled circle (see detect.cu in_disc): largest sy >= 0 with (sy / r)^2 + (dx / r)^2 < 1 in Float64, or -1
static int disc_half_height_host(int dx, int radius) {
    int best = -1;
    for (int sy = 0; sy <= radius; ++sy) {
        const volatile double vy = (double)sy / (double)radius, vx = (double)dx / (double)radius;
        const volatile double yy = vy * vy, xx = vx * vx;   // (volatile: the products are rounded before the sum, no contraction)
        if (yy + xx < 1.0) best = sy; else break;
    }
    return best;
}

static int prep_detect(slamklt_ctx* c, DetArgs& a) {
    a.sy_valid = 0;
    if (a.n_cur > 0 && a.radius >= 1 && 2 * a.radius + 1 <= 64) {
        for (int i = 0; i < 2 * a.radius + 1; ++i) a.sy[i] = (signed char)disc_half_height_host(i - a.radius, a.radius);
        a.sy_valid = 1;
    }
    if (a.n_cur > 0 && a.hw == 6) {
        // table of the 2^13 tap-subset sums of the mask blur's y pass (detect.cu, k_detect_cells2): entry `pat` adds the taps whose
        // bit is set in tap order with the same Float64 additions the tap loop performs (adding 0.0 for a clear bit changes nothing)
        long long key;
        std::memcpy(&key, &a.kw[0], sizeof(key));
        auto it = c->ytab_cache.find(key);
        if (it == c->ytab_cache.end()) {
            std::vector<double> tab(1u << 13);
            for (unsigned pat = 0; pat < (1u << 13); ++pat) {
                double acc = 0.0;
                for (int t = 0; t < 13; ++t) acc += ((pat >> t) & 1u) ? a.kw[t] : 0.0;
                tab[pat] = acc;
            }
            double* d = nullptr;
            CK(cudaMalloc(&d, tab.size() * sizeof(double)));
            CK(cudaMemcpyAsync(d, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            it = c->ytab_cache.emplace(key, d).first;
        }
        a.ytab = it->second;
    }
    return 0;
}

int slamklt_detect(slamklt_ctx* c, const void* img, int dtype, int H, int W, int ld, const double* cur, int n_cur,
                   const slamklt_detect_params* p, int64_t* out_yx, int cap, int* n_out) {
    if (!c || !img || !n_out || (!out_yx && cap > 0)) return fail(SLAMKLT_E_INVALID, "NULL argument");
    if (dtype < 0 || dtype > 2) return fail(SLAMKLT_E_INVALID, "unknown dtype %d", dtype);
    if (H < 3 || W < 3 || ld < H || n_cur < 0 || cap < 0) return fail(SLAMKLT_E_INVALID, "bad shape");
    if (n_cur > 0 && !cur) return fail(SLAMKLT_E_INVALID, "cur_pts is NULL");
    if (!p) return fail(SLAMKLT_E_INVALID, "params is NULL");
    *n_out = 0;
    if (n_cur >= p->max_points) return 0;  // extractor.jl:64
    DetArgs a;
    int r = fill_det(p, H, W, n_cur, &a);
    if (r) return r;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    if ((r = prep_detect(c, a))) return r;
    {
        // (the register-tiled kernel reads UInt8 frames directly: the repacked upload applies whenever it covers the request)
        bool packed = false;
        if (detect2_supported(a) && (r = upload_one_repacked(c, img, dtype, ld, H, W, &packed))) return r;
        if (packed) dtype = SLAMKLT_U8;
        else if ((r = upload_frames(c, c->staging, img, dtype, ld, 0, 1, H, W))) return r;
    }
    const void* d_img;
    int d_type = dtype;
    if (dtype == SLAMKLT_F64 || detect2_supported(a)) d_img = c->staging.p;  // (the register-tiled kernel converts on load)
    else {
        d_type = SLAMKLT_F64;
        if ((r = c->img64.ensure((size_t)H * W * 8))) return r;
        // reuse the convert kernel through a throw-away geometry: only the f64 copy is consumed
        PyrGeom g;
        if ((r = make_geom(H, W, 0, &g))) return r;
        DevBuf& scratch = c->outp;  // any buffer large enough for one level-0 frame
        if ((r = scratch.ensure(g.frame_elems * sizeof(float)))) return r;
        FrameSet fs{(float*)scratch.p, g.frame_elems, 1, 0};
        c->launches += launch_convert(c->stream, c->staging.p, dtype, H, (size_t)H * W, fs, 0, 1, g, (double*)c->img64.p, c->hk());
        CKL();
        d_img = c->img64.p;
    }
    r = run_detect(c, a, d_img, d_type, 1, cur, cap, out_yx, n_out);
    if (r) return r;
    if (*n_out > cap) return fail(SLAMKLT_E_CAPACITY, "detected %d keypoints but cap is %d", *n_out, cap);
    return 0;
}

// ---- batch -----------------------------------------------------------------------------------
int slamklt_batch_create(slamklt_ctx* c, int H, int W, int levels, int n_frames, int max_pts, slamklt_batch** out) {
    if (!c || !out) return fail(SLAMKLT_E_INVALID, "NULL argument");
    if (n_frames < 1 || max_pts < 0) return fail(SLAMKLT_E_INVALID, "bad batch shape");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    PyrGeom g;
    int r = make_geom(H, W, levels, &g);
    if (r) return r;
    slamklt_batch* b = new slamklt_batch();
    b->g = g; b->n_frames = n_frames; b->n_slots = n_frames + 1; b->slot0 = 0; b->max_pts = max_pts;
    cudaError_t e = cudaMalloc(&b->base, (g.frame_elems * b->n_slots + alloc_slack(g)) * sizeof(float));
    if (e != cudaSuccess) { delete b; return fail(SLAMKLT_E_CUDA, "cudaMalloc batch (%zu bytes) failed: %s", g.frame_elems * sizeof(float) * b->n_slots, cudaGetErrorString(e)); }
    cudaMemsetAsync(b->base, 0, (g.frame_elems * b->n_slots + alloc_slack(g)) * sizeof(float), c->stream);
    if ((r = b->pts.ensure((size_t)n_frames * max_pts * 16 + 16))) return r;
    if ((r = b->outp.ensure((size_t)n_frames * max_pts * 16 + 16))) return r;
    if ((r = b->status.ensure((size_t)n_frames * max_pts + 16))) return r;
    b->views.resize(b->n_slots, nullptr);
    CK(cudaEventCreateWithFlags(&b->ev_lk_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&b->ev_step_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&b->ev_pack, cudaEventDisableTiming));
    if ((r = make_maps(c, g, b->base, b->n_slots, &b->d_maps))) return r;
    // Scratch ring for the y-filtered product planes: SLAMKLT_BUILD_GROUP frames per build group (default 0 = off).  Measured on
    // B200 (64 KITTI frames, ncu --cache-control none): with groups of 8 and the L2 window the T planes never reach HBM (DRAM
    // traffic of a build 2.46 -> 1.73 GB) but the build gets slower, 0.75 -> 1.45 ms: 128 launches instead of 16, and a group's
    // column kernel is a single 84 %-full round of warps (55 us for 8 frames against 32 us for its share of the 64-frame launch).
    {
        static const int want = [] { const char* e = getenv("SLAMKLT_BUILD_GROUP"); return e ? atoi(e) : 0; }();
        if (want > 0 && n_frames >= 2 * want) {
            size_t off = 0;
            for (int l = 0; l < g.nl; ++l) { b->ts.off[l] = off; off += 3 * g.lv[l].plane_elems; }
            b->ts.stride = off; b->ts.ring = want;
            const size_t bytes = off * want * sizeof(float);
            cudaError_t e2 = cudaMalloc(&b->ts.base, bytes + alloc_slack(g) * sizeof(float));
            if (e2 != cudaSuccess) { b->ts.base = nullptr; cudaGetLastError(); }
            else {
                cudaMemsetAsync(b->ts.base, 0, bytes + alloc_slack(g) * sizeof(float), c->stream);
                b->build_group = want;
                set_l2_window(c, b->ts.base, bytes);
            }
        }
    }
    *out = b;
    return 0;
}

int slamklt_batch_destroy(slamklt_ctx* c, slamklt_batch* b) {
    if (!b) return 0;
    if (!c) return fail(SLAMKLT_E_INVALID, "ctx is NULL");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaStreamSynchronize(c->lk_stream));
    CK(cudaStreamSynchronize(c->copy_stream));
    CK(cudaStreamSynchronize(c->raw_stream));
    CK(cudaStreamSynchronize(c->d2h_stream));
    for (auto* v : b->views) delete v;
    if (b->ev_lk_done) cudaEventDestroy(b->ev_lk_done);
    if (b->ev_step_done) cudaEventDestroy(b->ev_step_done);
    if (b->ev_pack) cudaEventDestroy(b->ev_pack);
    for (auto e : b->raw_ev) cudaEventDestroy(e);
    b->staging.release(); b->staging8.release(); b->img64.release(); b->pts.release(); b->outp.release(); b->status.release(); b->gtab.release();
    b->h_pack.release();
    if (b->d_maps) cudaFree(b->d_maps);
    if (b->ts.base) cudaFree(b->ts.base);
    cudaFree(b->base);
    delete b;
    return 0;
}

int slamklt_batch_prime(slamklt_ctx* c, slamklt_batch* b, const void* img, int dtype, int ld, double sigma, int mode) {
    if (!c || !b || !img) return fail(SLAMKLT_E_INVALID, "NULL argument");
    if (dtype < 0 || dtype > 2) return fail(SLAMKLT_E_INVALID, "unknown dtype %d", dtype);
    if (ld < b->g.H0) return fail(SLAMKLT_E_INVALID, "ld < H");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    BATCH_WAIT_LK(c, b);
    int r = upload_frames(c, c->staging, img, dtype, ld, 0, 1, b->g.H0, b->g.W0);
    if (r) return r;
    r = build_frames(c, fs_of(b), 0, 1, b->g, c->staging.p, dtype, sigma, mode, nullptr);
    if (r) return r;
    CK(cudaStreamSynchronize(c->stream));
    b->primed = true;
    b->quiesced = !b->step_pending;
    return 0;
}

int slamklt_batch_upload(slamklt_ctx* c, slamklt_batch* b, const void* imgs, int dtype, int ld, size_t frame_stride_bytes,
                         const double* pts, int n_pts) {
    if (!c || !b || !imgs) return fail(SLAMKLT_E_INVALID, "NULL argument");
    if (dtype < 0 || dtype > 2) return fail(SLAMKLT_E_INVALID, "unknown dtype %d", dtype);
    if (ld < b->g.H0) return fail(SLAMKLT_E_INVALID, "ld < H");
    if (n_pts < 0 || n_pts > b->max_pts) return fail(SLAMKLT_E_INVALID, "n_pts %d outside [0,%d]", n_pts, b->max_pts);
    if (n_pts > 0 && !pts) return fail(SLAMKLT_E_INVALID, "pts is NULL");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    BATCH_WAIT_LK(c, b);
    b->quiesced = false;
    int r;
    // Float64 frames that hold 8-bit data travel as one byte per pixel here too (see batch_pipeline): lossless, the device rebuilds
    // the identical Float64.  The staging buffer may still be read by work queued earlier on the compute stream, and the repacked
    // bytes land at different offsets than plain frames would: order the copy behind that work (it runs on the compute stream).
    const int H = b->g.H0, W = b->g.W0, nf = b->n_frames;
    const size_t npx = (size_t)H * W;
    bool packed = false;
    if (dtype == SLAMKLT_F64 && ld == H && (frame_stride_bytes == npx * 8 || nf == 1) && (size_t)nf * npx >= (1u << 17) &&
        getenv("SLAMKLT_NO_PACK") == nullptr) {
        if ((r = b->h_pack.ensure((size_t)nf * npx))) return r;
        if ((r = b->staging.ensure((size_t)nf * npx * 8))) return r;
        if ((r = ensure_pool(c))) return r;
        CK(cudaStreamSynchronize(c->copy_stream));  // (an earlier step or upload may still be copying out of the repack buffer)
        CK(cudaEventSynchronize(b->ev_pack));
        std::atomic<int> bad{0};
        uint8_t* hp = (uint8_t*)b->h_pack.p;
        const int cols = 64, blocks = (W + cols - 1) / cols;
        c->pool->start(nf * blocks, [=, &bad](int item) {
            const int f = item / blocks, x0 = (item - f * blocks) * cols, x1 = std::min(W, x0 + cols);
            const double* sp = (const double*)imgs + (size_t)f * npx + (size_t)x0 * H;
            if (!pack_u8_exact(sp, hp + (size_t)f * npx + (size_t)x0 * H, (size_t)(x1 - x0) * H)) bad.store(1);
        });
        c->pool->wait();
        if (bad.load() == 0) {
            CK(cudaMemcpyAsync(b->staging.p, hp, (size_t)nf * npx, cudaMemcpyHostToDevice, c->stream));
            CK(cudaEventRecord(b->ev_pack, c->stream));   // the repack buffer is free again once this copy has run
            c->h2d += (uint64_t)nf * npx;
            dtype = SLAMKLT_U8;
            packed = true;
        }
    }
    if (!packed && (r = upload_frames(c, b->staging, imgs, dtype, ld, frame_stride_bytes, b->n_frames, b->g.H0, b->g.W0))) return r;
    if (n_pts > 0) {
        CK(cudaMemcpyAsync(b->pts.p, pts, (size_t)b->n_frames * n_pts * 16, cudaMemcpyHostToDevice, c->stream));
        c->h2d += (uint64_t)b->n_frames * n_pts * 16;
    }
    b->n_pts = n_pts; b->up_dtype = dtype; b->up_ld = ld;
    return 0;
}

int slamklt_batch_build(slamklt_ctx* c, slamklt_batch* b, double sigma, int mode) {
    if (!c || !b) return fail(SLAMKLT_E_INVALID, "NULL argument");
    if (b->up_dtype < 0) return fail(SLAMKLT_E_INVALID, b->up_dtype == -2 ? "the frames of the last step sit on the device in two formats: upload them again (slamklt_batch_upload)" : "no frames uploaded");
    if (mode != SLAMKLT_MODE_UPDATE && mode != SLAMKLT_MODE_CTOR) return fail(SLAMKLT_E_INVALID, "unknown mode %d", mode);
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    BATCH_WAIT_LK(c, b);
    b->quiesced = false;
    return build_frames(c, fs_of(b), 1, b->n_frames, b->g, b->staging.p, b->up_dtype, sigma, mode, nullptr, &b->ts, b->build_group);
}

int slamklt_batch_track(slamklt_ctx* c, slamklt_batch* b, const slamklt_lk_params* p) {
    if (!c || !b) return fail(SLAMKLT_E_INVALID, "NULL argument");
    int r = check_lk(p, b->g.nl, b->g.nl);
    if (r) return r;
    if (!b->primed) return fail(SLAMKLT_E_INVALID, "batch slot 0 was never built (call slamklt_batch_prime)");
    if (b->n_pts == 0) return 0;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    BATCH_WAIT_LK(c, b);
    b->quiesced = false;
    LKArgs a{};
    a.A = fs_of(b); a.B = fs_of(b); a.offA = 0; a.offB = 1;
    fill_lk_levels(b->g, &a);
    a.mode = 1;
    a.pts = (const double*)b->pts.p; a.disp_in = nullptr; a.disp_out = nullptr;
    a.out_pts = (double*)b->outp.p; a.status = (uint8_t*)b->status.p;
    a.n_per_frame = b->n_pts; a.n_frames = b->n_frames;
    a.iterations = p->iterations; a.window = p->window_size; a.levels = p->pyramid_levels;
    a.eig_thr = p->eigenvalue_threshold; a.eps = p->epsilon; a.max_dist = p->max_distance;
    a.counters = c->d_counters; a.work = work_slot(c);
    a.mapsA = b->d_maps; a.mapsB = b->d_maps;
    if ((r = set_gtab(b->gtab, &a))) return r;
    c->launches += launch_lk(c->stream, a, c->hk());
    CKL();
    prof_end(c);
    return 0;
}

// Stereo matching of two batches (mapper.jl:51-60 for n_frames keyframes at once): pair i tracks `to`'s uploaded points from
// frame i of `from` to frame i of `to`; results land in `to`'s result buffers (slamklt_batch_download(to, ...)).
int slamklt_batch_track_cross(slamklt_ctx* c, slamklt_batch* from, slamklt_batch* to, const slamklt_lk_params* p) {
    if (!c || !from || !to) return fail(SLAMKLT_E_INVALID, "NULL argument");
    int r = check_lk(p, from->g.nl, to->g.nl);
    if (r) return r;
    if (from->g.H0 != to->g.H0 || from->g.W0 != to->g.W0 || from->g.nl != to->g.nl || from->n_frames != to->n_frames)
        return fail(SLAMKLT_E_INVALID, "batch shapes differ");
    if (to->n_pts == 0) return 0;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    BATCH_WAIT_LK(c, from);
    BATCH_WAIT_LK(c, to);
    from->quiesced = false; to->quiesced = false;
    LKArgs a{};
    a.A = fs_of(from); a.B = fs_of(to); a.offA = 1; a.offB = 1;
    fill_lk_levels(to->g, &a);
    a.mode = 1;
    a.pts = (const double*)to->pts.p; a.disp_in = nullptr; a.disp_out = nullptr;
    a.out_pts = (double*)to->outp.p; a.status = (uint8_t*)to->status.p;
    a.n_per_frame = to->n_pts; a.n_frames = to->n_frames;
    a.iterations = p->iterations; a.window = p->window_size; a.levels = p->pyramid_levels;
    a.eig_thr = p->eigenvalue_threshold; a.eps = p->epsilon; a.max_dist = p->max_distance;
    a.counters = c->d_counters; a.work = work_slot(c);
    a.mapsA = from->d_maps; a.mapsB = to->d_maps;
    if ((r = set_gtab(to->gtab, &a))) return r;
    c->launches += launch_lk(c->stream, a, c->hk());
    CKL();
    prof_end(c);
    return 0;
}

int slamklt_batch_download(slamklt_ctx* c, slamklt_batch* b, double* out_pts, uint8_t* status) {
    if (!c || !b) return fail(SLAMKLT_E_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    if (b->lk_pending) { CK(cudaStreamWaitEvent(c->stream, b->ev_lk_done, 0)); b->lk_pending = false; }
    const size_t n = (size_t)b->n_frames * b->n_pts;
    if (n > 0) {
        if (out_pts) CK(cudaMemcpyAsync(out_pts, b->outp.p, n * 16, cudaMemcpyDeviceToHost, c->stream));
        if (status) CK(cudaMemcpyAsync(status, b->status.p, n, cudaMemcpyDeviceToHost, c->stream));
        c->d2h += (out_pts ? n * 16 : 0) + (status ? n : 0);
    }
    CK(cudaStreamSynchronize(c->stream));
    b->quiesced = !b->step_pending;
    return 0;
}

int slamklt_batch_rotate(slamklt_ctx* c, slamklt_batch* b) {
    if (!c || !b) return fail(SLAMKLT_E_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(c->mu);
    b->slot0 = (b->slot0 + b->n_frames) % b->n_slots;  // last frame of this batch becomes slot 0
    return 0;
}

// Chunked pipeline shared by slamklt_batch_step (host buffers) and slamklt_batch_process (device-resident frames):
// H2D of chunk k+1 (copy streams) || pyramid build of chunk k (main + side streams) || tracking of chunk k-1 (lk stream)
// || D2H of chunk k-2 (d2h stream).  imgs == nullptr: frames already sit in the batch staging buffer.
// Nothing here waits for the device: the caller synchronises (slamklt_batch_step_end / slamklt_batch_download).
//
// Float64 host frames reach the device through TWO engines at once (round 2): host worker threads repack chunks that hold 8-bit
// data to one byte per pixel (pack_u8_exact: lossless, the device rebuilds the identical Float64) while the copy engine ships
// other chunks as plain Float64 -- both pull from host memory concurrently, so the step is bound by the sum of their rates
// instead of by the slower of "all repacked" / "all plain".  Which chunk goes which way is a greedy schedule over the two
// measured rates (source bytes per second of the worker pool, bytes per second of the plain copies), refreshed every step.
enum { UP_PLAIN = 0, UP_PACK = 1, UP_RAW = 2 };

static long long now_ns() { return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static int batch_pipeline(slamklt_ctx* c, slamklt_batch* b, const void* imgs, int dtype, int ld, size_t frame_stride_bytes,
                          const double* pts, int n_pts, double sigma, int mode, const slamklt_lk_params* p, double* out_pts,
                          uint8_t* status, int want_chunks, bool throughput = false) {
    int r;
    const int H = b->g.H0, W = b->g.W0, nf = b->n_frames;
    const size_t es = dtype_size(dtype), fbytes = (size_t)H * W * es;
    if ((r = b->staging.ensure((size_t)nf * fbytes))) return r;
    const size_t npx = (size_t)H * W;
    const bool contiguous = (ld == H) && (frame_stride_bytes == fbytes || nf == 1);
    // ---- chunks (frame ranges, each built + tracked by its own launches) and how each one's frames reach the device
    struct Chunk { int f0, n; char how; };
    std::vector<Chunk> chunks;
    auto uniform = [&](int want, char how) {
        const int len = nf >= 2 * want ? (nf + want - 1) / want : nf;
        for (int f = 0; f < nf; f += len) chunks.push_back(Chunk{f, std::min(len, nf - f), how});
    };
    bool try_pack = imgs && dtype == SLAMKLT_F64 && ld == H && (size_t)nf * H * W >= (1u << 20) && getenv("SLAMKLT_NO_PACK") == nullptr;
    if (try_pack) {
        if ((r = b->h_pack.ensure((size_t)nf * H * W))) return r;
        if ((r = ensure_pool(c))) return r;
        CK(cudaEventSynchronize(b->ev_pack));   // (slamklt_batch_upload may still be copying out of the repack buffer)
        // rate of the plain copies of this batch's previous step (its events have completed: the step was waited for)
        if (!b->raw_meas.empty()) {
            double ms_sum = 0.0, bytes = 0.0;
            for (auto& m : b->raw_meas) {
                float ms = 0.f;
                if (cudaEventElapsedTime(&ms, b->raw_ev[2 * m.first], b->raw_ev[2 * m.first + 1]) == cudaSuccess && ms > 0.f) { ms_sum += ms; bytes += (double)m.second; }
                else cudaGetLastError();
            }
            if (ms_sum > 0.0) { const double rate = bytes / (ms_sum * 1e-3); c->raw_Bps = c->raw_Bps > 0 ? 0.5 * (c->raw_Bps + rate) : rate; }
            b->raw_meas.clear();
        }
        // the plain-copy engine only helps when it can read the caller's buffer directly (page-locked memory)
        cudaPointerAttributes at;
        bool pinned = cudaPointerGetAttributes(&at, imgs) == cudaSuccess && at.type == cudaMemoryTypeHost;
        if (!pinned) cudaGetLastError();
        const bool two_engines = pinned && contiguous && nf >= 8 && getenv("SLAMKLT_NO_HYBRID") == nullptr;
        const char* forced = getenv("SLAMKLT_UPLOAD_PLAN");  // test / experiment knob, e.g. "PRPPRPPR" (P = repacked, R = plain), cycled over the chunks
        if (two_engines && forced && *forced) {
            uniform(want_chunks, UP_PACK);
            const size_t fl = strlen(forced);
            for (size_t k = 0; k < chunks.size(); ++k) if (forced[k % fl] == 'R' || forced[k % fl] == 'r') chunks[k].how = UP_RAW;
        } else if (two_engines && throughput) {
            // Throughput mode (another batch's kernels hide this step's uploads): the first `a` frames travel as plain Float64 on
            // the copy engine while the workers repack the rest, a chosen so that both finish together:
            //   a*fb/rw + (nf-a)*fb/8/rw  =  (nf-a)*fb/pk      (the repacked bytes cross the same link)
            // Few, large chunks: the kernels run on full grids (measured on B200: 8 chunks of 8 frames cost 2.3 ms of GPU time per
            // step against 1.5 ms for 2 chunks).
            const double pk = c->pack_Bps > 0 ? c->pack_Bps : 60e9, rw = c->raw_Bps > 0 ? c->raw_Bps : 45e9;
            const double share = std::max(0.0, (1.0 / pk - 0.125 / rw) / (0.875 / rw + 1.0 / pk));
            int n_raw = std::min(nf - 1, (int)std::lround(share * nf));
            if (c->raw_Bps == 0.0 && n_raw == 0) n_raw = 1;  // (first step: get the plain-copy rate measured)
            // chunks stay in frame order (pair i tracks slot i -> slot i + 1, so frame i - 1 must be built before frame i is
            // tracked): the plain part first -- its copy starts at once -- then the repacked part
            if (n_raw > 0) chunks.push_back(Chunk{0, n_raw, UP_RAW});
            chunks.push_back(Chunk{n_raw, nf - n_raw, UP_PACK});
        } else {
            uniform(throughput ? std::min(want_chunks, 2) : want_chunks, UP_PACK);
        }
    } else {
        uniform(want_chunks, UP_PLAIN);
    }
    const int nchunks = (int)chunks.size();
    // Repacked frames sit compactly (1 B/px) at frame_index * npx, plain ones at frame_index * 8 * npx: in one buffer a plain chunk
    // would overlap the repacked chunks behind it, so a two-engine plan keeps its repacked frames in a buffer of their own.
    bool any_raw = false;
    for (const Chunk& ch : chunks) any_raw |= ch.how == UP_RAW;
    if (any_raw && (r = b->staging8.ensure((size_t)nf * npx))) return r;
    char* const pk_base = any_raw ? (char*)b->staging8.p : (char*)b->staging.p;
    while ((int)c->pipe_ev.size() < 3 * nchunks + 2) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->pipe_ev.push_back(e);
    }
    cudaEvent_t* evH2D = c->pipe_ev.data();
    cudaEvent_t* evBuilt = c->pipe_ev.data() + nchunks;
    cudaEvent_t* evLK = c->pipe_ev.data() + 2 * nchunks;
    cudaEvent_t evStart = c->pipe_ev[3 * nchunks], evPts = c->pipe_ev[3 * nchunks + 1];
    const bool side_lk = !c->prof_on;  // the per-kernel profiler times one serial stream
    cudaStream_t lks = side_lk ? c->lk_stream : c->stream;
    // A batch with nothing in flight (slamklt_batch_step_end / _download / _prime returned) can be written at once: its copies need
    // not queue behind ANOTHER batch's pyramid builds on the compute stream.  Otherwise order the copy streams after everything
    // already queued on the compute stream (staging / slot reuse).
    const bool fresh = b->quiesced;
    b->quiesced = false;
    if (b->lk_pending) {  // the previous tracking kernel of THIS batch still reads its slots / points / result buffers
        CK(cudaStreamWaitEvent(c->stream, b->ev_lk_done, 0));
        b->lk_pending = false;
    }
    if (!fresh) {
        CK(cudaEventRecord(evStart, c->stream));
        CK(cudaStreamWaitEvent(c->copy_stream, evStart, 0));
        CK(cudaStreamWaitEvent(c->raw_stream, evStart, 0));
        if (side_lk) CK(cudaStreamWaitEvent(c->lk_stream, evStart, 0));
    }
    int n_packed = 0, n_plain = 0;
    std::atomic<int> pack_bad{0};
    std::atomic<long long> pack_t1{0};
    long long pack_t0 = 0;
    double pack_ns = 0.0, pack_bytes = 0.0;
    // repack chunk k on the workers (asynchronously: the calling thread meanwhile queues the GPU work of the chunk before)
    auto start_pack = [&](int k) {
        const int f0 = chunks[k].f0, n = chunks[k].n;
        uint8_t* hp = (uint8_t*)b->h_pack.p + (size_t)f0 * npx;
        const int cols = 64, blocks = (W + cols - 1) / cols;
        pack_t0 = now_ns();
        pack_t1.store(pack_t0);
        c->pool->start(n * blocks, [=, &pack_bad, &pack_t1](int item) {
            const int f = item / blocks, x0 = (item - f * blocks) * cols, x1 = std::min(W, x0 + cols);
            const double* sp = (const double*)((const char*)imgs + (size_t)(f0 + f) * frame_stride_bytes) + (size_t)x0 * H;
            if (!pack_u8_exact(sp, hp + (size_t)f * npx + (size_t)x0 * H, (size_t)(x1 - x0) * H)) pack_bad.store(1);
            const long long t = now_ns();
            long long seen = pack_t1.load(std::memory_order_relaxed);
            while (t > seen && !pack_t1.compare_exchange_weak(seen, t, std::memory_order_relaxed)) {}
        });
    };
    auto next_packed = [&](int after) { for (int k = after + 1; k < nchunks; ++k) if (chunks[k].how == UP_PACK) return k; return -1; };
    struct PoolJoin { HostPool* p; ~PoolJoin() { if (p) p->wait(); } } pool_join{try_pack ? c->pool : nullptr};  // no job outlives this call
    if (try_pack) { const int k0 = next_packed(-1); if (k0 >= 0) start_pack(k0); }
    // plain copy of chunk k (whole frames as they are) on stream st; timed when it is one of the hybrid plan's copies
    auto copy_plain = [&](int k, cudaStream_t st, bool timed) -> int {
        const int f0 = chunks[k].f0, n = chunks[k].n, f1 = f0 + n;
        char* dst = (char*)b->staging.p + (size_t)f0 * fbytes;
        if (timed) {
            const int slot = (int)b->raw_meas.size();
            while ((int)b->raw_ev.size() < 2 * (slot + 1)) { cudaEvent_t e; CK(cudaEventCreate(&e)); b->raw_ev.push_back(e); }
            CK(cudaEventRecord(b->raw_ev[2 * slot], st));
            b->raw_meas.emplace_back(slot, (size_t)n * fbytes);
        }
        if (contiguous) {
            CK(cudaMemcpyAsync(dst, (const char*)imgs + (size_t)f0 * fbytes, (size_t)n * fbytes, cudaMemcpyHostToDevice, st));
        } else {
            for (int f = f0; f < f1; ++f)
                CK(cudaMemcpy2DAsync((char*)b->staging.p + (size_t)f * fbytes, (size_t)H * es, (const char*)imgs + (size_t)f * frame_stride_bytes,
                                     (size_t)ld * es, (size_t)H * es, W, cudaMemcpyHostToDevice, st));
        }
        if (timed) CK(cudaEventRecord(b->raw_ev[2 * b->raw_meas.back().first + 1], st));
        CK(cudaEventRecord(evH2D[k], st));
        c->h2d += (uint64_t)n * fbytes;
        return 0;
    };
    // the plain copies of the plan are queued a few chunks ahead of the chunk the host is working on: far enough to keep the copy
    // engine busy while the calling thread helps repacking, near enough that a repacked chunk's small copy is not stuck behind them
    static const int raw_look = [] { const char* e = getenv("SLAMKLT_RAW_LOOKAHEAD"); return e ? atoi(e) : 3; }();
    int raw_next = 0;
    auto queue_raw_upto = [&](int last) -> int {
        for (; raw_next <= last && raw_next < nchunks; ++raw_next)
            if (chunks[raw_next].how == UP_RAW) { int rr = copy_plain(raw_next, c->raw_stream, true); if (rr) return rr; }
        return 0;
    };
    if (imgs) {
        if (n_pts > 0) {
            CK(cudaMemcpyAsync(b->pts.p, pts, (size_t)nf * n_pts * 16, cudaMemcpyHostToDevice, c->copy_stream));
            CK(cudaEventRecord(evPts, c->copy_stream));
            CK(cudaStreamWaitEvent(lks, evPts, 0));  // (chunk 0's frames may travel on the other copy stream)
            c->h2d += (uint64_t)nf * n_pts * 16;
        }
        b->n_pts = n_pts; b->up_ld = ld;
    }
    LKArgs a{};
    a.A = fs_of(b); a.B = fs_of(b);
    fill_lk_levels(b->g, &a);
    a.mode = 1;
    a.iterations = p->iterations; a.window = p->window_size; a.levels = p->pyramid_levels;
    a.eig_thr = p->eigenvalue_threshold; a.eps = p->epsilon; a.max_dist = p->max_distance;
    a.counters = c->d_counters; a.work = work_slot(c);
    a.n_per_frame = n_pts;
    a.mapsA = b->d_maps; a.mapsB = b->d_maps;
    a.n_frames = nf;
    if ((r = set_gtab(b->gtab, &a))) return r;
    char* const gtab0 = (char*)a.gtab;
    bool any_d2h = false;
    for (int k = 0; k < nchunks; ++k) {
        const int f0 = chunks[k].f0, n = chunks[k].n;
        int cd = dtype;                                                         // dtype this chunk reaches the device in
        const char* src = (const char*)b->staging.p + (size_t)f0 * fbytes;
        if (imgs) {
            if ((r = queue_raw_upto(throughput ? nchunks : k + raw_look))) return r;
            if (chunks[k].how == UP_PACK) {
                const long long tw0 = now_ns();
                c->pool->wait();
                c->dbg_wait_ns += (double)(now_ns() - tw0); c->dbg_pack_ns += (double)(pack_t1.load() - pack_t0);
                pack_ns += (double)(pack_t1.load() - pack_t0); pack_bytes += (double)n * fbytes;
                if (pack_bad.load() != 0) {
                    // not 8-bit data: this chunk and every chunk still planned for repacking travel as plain Float64
                    for (int j = k; j < nchunks; ++j) if (chunks[j].how == UP_PACK) chunks[j].how = UP_PLAIN;
                    try_pack = false;
                } else {
                    const uint8_t* hp = (const uint8_t*)b->h_pack.p + (size_t)f0 * npx;
                    char* dst = pk_base + (size_t)f0 * npx;
                    CK(cudaMemcpyAsync(dst, hp, (size_t)n * npx, cudaMemcpyHostToDevice, c->copy_stream));
                    CK(cudaEventRecord(evH2D[k], c->copy_stream));
                    c->h2d += (uint64_t)n * npx;
                    cd = SLAMKLT_U8; src = dst;
                    ++n_packed;
                    const int kn = next_packed(k);
                    if (kn >= 0) start_pack(kn);
                }
            }
            if (chunks[k].how == UP_PLAIN) { if ((r = copy_plain(k, c->copy_stream, false))) return r; }
            if (chunks[k].how != UP_PACK) ++n_plain;
            CK(cudaStreamWaitEvent(c->stream, evH2D[k], 0));
        }
        if ((r = build_frames(c, fs_of(b), 1 + f0, n, b->g, src, cd, sigma, mode, nullptr, &b->ts, b->build_group))) return r;
        if (n_pts > 0) {
            if (side_lk) { CK(cudaEventRecord(evBuilt[k], c->stream)); CK(cudaStreamWaitEvent(lks, evBuilt[k], 0)); }
            a.offA = f0; a.offB = f0 + 1; a.n_frames = n;
            a.pts = (const double*)b->pts.p + (size_t)f0 * n_pts * 2;
            a.out_pts = (double*)b->outp.p + (size_t)f0 * n_pts * 2;
            a.status = (uint8_t*)b->status.p + (size_t)f0 * n_pts;
            if (gtab0) a.gtab = gtab0 + (size_t)f0 * n_pts * a.gtab_levels * 16;
            if (nchunks > 1) a.work = work_slot(c);  // chunks may overlap on the lk stream's tail: one counter each
            c->launches += launch_lk(lks, a, c->hk());
            CKL();
            prof_end(c);
            CK(cudaEventRecord(evLK[k], lks));
            if (out_pts && status) {
                CK(cudaStreamWaitEvent(c->d2h_stream, evLK[k], 0));
                CK(cudaMemcpyAsync(out_pts + (size_t)f0 * n_pts * 2, a.out_pts, (size_t)n * n_pts * 16, cudaMemcpyDeviceToHost, c->d2h_stream));
                CK(cudaMemcpyAsync(status + (size_t)f0 * n_pts, a.status, (size_t)n * n_pts, cudaMemcpyDeviceToHost, c->d2h_stream));
                c->d2h += (uint64_t)n * n_pts * 17;
                any_d2h = true;
            }
        }
    }
    if (pack_ns > 0.0) { const double rate = pack_bytes / (pack_ns * 1e-9); c->pack_Bps = c->pack_Bps > 0 ? 0.5 * (c->pack_Bps + rate) : rate; }
    // what the staging buffers now hold (slamklt_batch_detect reads them): one dtype, or -2 = per-chunk list after a mixed step
    if (imgs) {
        b->up_dtype = n_plain == 0 ? SLAMKLT_U8 : (n_packed == 0 ? dtype : -2);
        b->up_chunks.clear();
        if (b->up_dtype == -2)
            for (const Chunk& ch : chunks)
                b->up_chunks.push_back(slamklt_batch::UpChunk{ch.f0, ch.n, ch.how == UP_PACK ? SLAMKLT_U8 : dtype,
                                                              ch.how == UP_PACK ? (const void*)(pk_base + (size_t)ch.f0 * npx)
                                                                                : (const void*)((const char*)b->staging.p + (size_t)ch.f0 * fbytes)});
    }
    // Only work that touches THIS batch again has to wait for its tracking kernels (see lk_pending above and
    // batch_wait_lk): another batch may build on the compute stream while this one is still being tracked.
    if (n_pts > 0 && side_lk) { CK(cudaEventRecord(b->ev_lk_done, lks)); b->lk_pending = true; }
    // one event that covers everything queued above: result copies after tracking after builds after uploads
    if (any_d2h) CK(cudaEventRecord(b->ev_step_done, c->d2h_stream));
    else if (n_pts > 0) CK(cudaEventRecord(b->ev_step_done, lks));
    else CK(cudaEventRecord(b->ev_step_done, c->stream));
    return 0;
}

static int check_step_args(slamklt_ctx* c, slamklt_batch* b, const void* imgs, int dtype, int ld, const double* pts, int n_pts, int mode,
                           const slamklt_lk_params* p, double* out_pts, uint8_t* status) {
    if (!c || !b || !imgs) return fail(SLAMKLT_E_INVALID, "NULL argument");
    if (dtype < 0 || dtype > 2) return fail(SLAMKLT_E_INVALID, "unknown dtype %d", dtype);
    if (ld < b->g.H0) return fail(SLAMKLT_E_INVALID, "ld < H");
    if (n_pts < 0 || n_pts > b->max_pts) return fail(SLAMKLT_E_INVALID, "n_pts %d outside [0,%d]", n_pts, b->max_pts);
    if (n_pts > 0 && (!pts || !out_pts || !status)) return fail(SLAMKLT_E_INVALID, "NULL argument");
    if (mode != SLAMKLT_MODE_UPDATE && mode != SLAMKLT_MODE_CTOR) return fail(SLAMKLT_E_INVALID, "unknown mode %d", mode);
    int r = check_lk(p, b->g.nl, b->g.nl);
    if (r) return r;
    if (!b->primed) return fail(SLAMKLT_E_INVALID, "batch slot 0 was never built (call slamklt_batch_prime)");
    if (b->step_pending) return fail(SLAMKLT_E_INVALID, "the batch has a step in flight (call slamklt_batch_step_end first)");
    return 0;
}

// First half of slamklt_batch_step: queue upload + build + track + download of one step and return without waiting for the
// device.  imgs, pts_yx, out_pts_yx and status must stay valid (and untouched) until slamklt_batch_step_end.  Several batches
// of one context may have a step in flight: their copies, builds and tracking kernels overlap (a stream consumer keeps two
// batches going so that the uploads of one hide behind the kernels of the other).
static int step_begin(slamklt_ctx* c, slamklt_batch* b, const void* imgs, int dtype, int ld, size_t frame_stride_bytes, const double* pts,
                      int n_pts, double sigma, int mode, const slamklt_lk_params* p, double* out_pts, uint8_t* status, bool throughput) {
    int r = check_step_args(c, b, imgs, dtype, ld, pts, n_pts, mode, p, out_pts, status);
    if (r) return r;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    // more chunks hide more of the copy but run the kernels on smaller grids: 8 for 8-byte pixels, fewer for light copies
    static const int forced = [] { const char* e = getenv("SLAMKLT_STEP_CHUNKS"); return e ? atoi(e) : 0; }();
    int want = dtype == SLAMKLT_F64 ? 8 : (dtype == SLAMKLT_F32 ? 4 : 2);
    // throughput mode: the caller keeps several batches in flight, so another batch's kernels hide this step's uploads and large
    // chunks (full grids) win; measured on B200, UInt8 frames: 1.55 ms per step with one chunk against 1.63 with two
    if (throughput && dtype != SLAMKLT_F64) want = 1;
    if (forced > 0) want = forced;
    const long long t0 = now_ns();
    if ((r = batch_pipeline(c, b, imgs, dtype, ld, frame_stride_bytes, pts, n_pts, sigma, mode, p, out_pts, status, want, throughput))) return r;
    c->dbg_begin_ns += (double)(now_ns() - t0); c->dbg_steps += 1;
    b->step_pending = true;
    return 0;
}

int slamklt_batch_step_begin(slamklt_ctx* c, slamklt_batch* b, const void* imgs, int dtype, int ld, size_t frame_stride_bytes,
                             const double* pts, int n_pts, double sigma, int mode, const slamklt_lk_params* p, double* out_pts, uint8_t* status) {
    return step_begin(c, b, imgs, dtype, ld, frame_stride_bytes, pts, n_pts, sigma, mode, p, out_pts, status, true);
}

// Second half: wait until the step's results sit in the caller's buffers, then rotate (the last frame becomes slot 0).
int slamklt_batch_step_end(slamklt_ctx* c, slamklt_batch* b) {
    if (!c || !b) return fail(SLAMKLT_E_INVALID, "NULL argument");
    if (!b->step_pending) return fail(SLAMKLT_E_INVALID, "no step in flight on this batch");
    {
        std::lock_guard<std::mutex> lk(c->mu);
        CK(cudaSetDevice(c->device));
        const long long t0 = now_ns();
        CK(cudaEventSynchronize(b->ev_step_done));
        c->dbg_end_ns += (double)(now_ns() - t0);
        b->step_pending = false;
        b->lk_pending = false;   // ev_step_done was recorded after the last tracking kernel
        b->quiesced = true;
    }
    return slamklt_batch_rotate(c, b);
}

// Whole step through host buffers: upload + build + track + download + rotate, pipelined in chunks of frames.
int slamklt_batch_step(slamklt_ctx* c, slamklt_batch* b, const void* imgs, int dtype, int ld, size_t frame_stride_bytes,
                       const double* pts, int n_pts, double sigma, int mode, const slamklt_lk_params* p, double* out_pts, uint8_t* status) {
    // latency mode: nothing else hides this step's uploads, so it is cut into more chunks that pipeline against each other
    int r = step_begin(c, b, imgs, dtype, ld, frame_stride_bytes, pts, n_pts, sigma, mode, p, out_pts, status, false);
    if (r) return r;
    return slamklt_batch_step_end(c, b);
}

// Device-resident step: frames and points were uploaded before (slamklt_batch_upload); build every pyramid and track every
// pair in one asynchronous call.  Results are fetched with slamklt_batch_download (which synchronises).
int slamklt_batch_process(slamklt_ctx* c, slamklt_batch* b, double sigma, int mode, const slamklt_lk_params* p) {
    if (!c || !b) return fail(SLAMKLT_E_INVALID, "NULL argument");
    if (b->up_dtype < 0) return fail(SLAMKLT_E_INVALID, b->up_dtype == -2 ? "the frames of the last step sit on the device in two formats: upload them again (slamklt_batch_upload)" : "no frames uploaded");
    if (mode != SLAMKLT_MODE_UPDATE && mode != SLAMKLT_MODE_CTOR) return fail(SLAMKLT_E_INVALID, "unknown mode %d", mode);
    int r = check_lk(p, b->g.nl, b->g.nl);
    if (r) return r;
    if (!b->primed) return fail(SLAMKLT_E_INVALID, "batch slot 0 was never built (call slamklt_batch_prime)");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    // measured on B200: chunking a device-resident batch only shrinks the grids (2.57 vs 2.30 ms per 64 frames), so one chunk
    static const int chunks = [] { const char* e = getenv("SLAMKLT_PROCESS_CHUNKS"); return e ? std::max(1, atoi(e)) : 1; }();  // experiment knob
    return batch_pipeline(c, b, nullptr, b->up_dtype, b->up_ld, 0, nullptr, b->n_pts, sigma, mode, p, nullptr, nullptr, chunks);
}

int slamklt_batch_slot(slamklt_batch* b, int slot, slamklt_pyr** out) {
    if (!b || !out) return fail(SLAMKLT_E_INVALID, "NULL argument");
    if (slot < 0 || slot >= b->n_slots) return fail(SLAMKLT_E_INVALID, "slot %d out of range", slot);
    slamklt_pyr*& v = b->views[slot];
    if (!v) { v = new slamklt_pyr(); v->g = b->g; v->parent = b; v->logical_slot = slot; v->owns = false; }
    v->built = true;
    b->quiesced = false;  // the view may be handed to calls that queue work on the compute stream
    *out = v;
    return 0;
}

int slamklt_batch_detect(slamklt_ctx* c, slamklt_batch* b, const double* cur, int n_cur, const slamklt_detect_params* p,
                         int64_t* out_yx, int cap, int* n_out) {
    if (!c || !b || !n_out || (!out_yx && cap > 0)) return fail(SLAMKLT_E_INVALID, "NULL argument");
    if (b->up_dtype == -1) return fail(SLAMKLT_E_INVALID, "no frames uploaded");
    if (n_cur < 0 || cap < 0) return fail(SLAMKLT_E_INVALID, "bad shape");
    if (n_cur > 0 && !cur) return fail(SLAMKLT_E_INVALID, "cur_pts is NULL");
    if (!p) return fail(SLAMKLT_E_INVALID, "params is NULL");
    for (int f = 0; f < b->n_frames; ++f) n_out[f] = 0;
    if (n_cur >= p->max_points) return 0;
    DetArgs a;
    int r = fill_det(p, b->g.H0, b->g.W0, n_cur, &a);
    if (r) return r;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    BATCH_WAIT_LK(c, b);
    b->quiesced = false;
    if ((r = prep_detect(c, a))) return r;
    const void* d_img;
    int d_type = b->up_dtype;
    if (b->up_dtype == -2) {
        // the last step shipped some chunks repacked and some plain: one Float64 copy, converted chunk by chunk
        d_type = SLAMKLT_F64;
        const size_t fpx = (size_t)b->g.H0 * b->g.W0;
        if ((r = b->img64.ensure((size_t)b->n_frames * fpx * 8))) return r;
        for (const auto& ch : b->up_chunks) {
            c->launches += launch_convert(c->stream, ch.ptr, ch.dtype, b->g.H0, fpx, fs_of(b), 1 + ch.f0, ch.n, b->g, (double*)b->img64.p + (size_t)ch.f0 * fpx, c->hk());
            CKL();
        }
        d_img = b->img64.p;
    } else if (b->up_dtype == SLAMKLT_F64 || detect2_supported(a)) d_img = b->staging.p;  // (the register-tiled kernel converts on load)
    else {
        d_type = SLAMKLT_F64;
        if ((r = b->img64.ensure((size_t)b->n_frames * b->g.H0 * b->g.W0 * 8))) return r;
        c->launches += launch_convert(c->stream, b->staging.p, b->up_dtype, b->g.H0, (size_t)b->g.H0 * b->g.W0, fs_of(b), 1, b->n_frames, b->g, (double*)b->img64.p, c->hk());
        CKL();
        d_img = b->img64.p;
    }
    r = run_detect(c, a, d_img, d_type, b->n_frames, cur, cap, out_yx, n_out);
    if (r) return r;
    for (int f = 0; f < b->n_frames; ++f)
        if (n_out[f] > cap) return fail(SLAMKLT_E_CAPACITY, "frame %d: detected %d keypoints but cap is %d", f, n_out[f], cap);
    return 0;
}

int slamklt_host_alloc(size_t bytes, void** out) {
    if (!out) return fail(SLAMKLT_E_INVALID, "out is NULL");
    CK(cudaMallocHost(out, bytes));
    return 0;
}

int slamklt_host_free(void* p) {
    if (p) CK(cudaFreeHost(p));
    return 0;
}

}  // extern "C"

