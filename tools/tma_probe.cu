// Stand-alone probe of 3-D tensor-map TMA loads on sm_100a: which descriptor placements / box shapes the hardware accepts.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe tools/tma_probe.cu ; run on the GPU box.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ unsigned s_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int BR, int BC>
__global__ void k_probe(const CUtensorMap* gmap, const __grid_constant__ CUtensorMap pmap, int use_param, int c0, int c1, int c2, float* out) {
    __shared__ __align__(128) float tile[BC][BR];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(&bar)), "r"((unsigned)(BR * BC * 4)) : "memory");
        const CUtensorMap* m = use_param ? &pmap : gmap;
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(s_u32(&tile[0][0])),
                     "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(s_u32(&bar))
                     : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra W_%=;\n\t}" ::"r"(s_u32(&bar)), "r"(0u)
        : "memory");
    for (int i = threadIdx.x; i < BR * BC; i += 32) out[i] = (&tile[0][0])[i];
}

template <int BR, int BC>
static int run(EncodeTiledFn enc, const char* name, float* dbase, int rows, int cols, int slots, size_t frame_elems, int use_param, int c0, int c1, int c2,
               const std::vector<float>& host) {
    CUtensorMap m;
    const cuuint64_t dims[3] = {(cuuint64_t)rows, (cuuint64_t)cols, (cuuint64_t)slots};
    const cuuint64_t strides[2] = {(cuuint64_t)rows * 4, (cuuint64_t)frame_elems * 4};
    const cuuint32_t box[3] = {BR, BC, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dbase, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%-40s encode failed %d\n", name, (int)r); return 1; }
    CUtensorMap* dm;
    cudaMalloc(&dm, sizeof(m));
    cudaMemcpy(dm, &m, sizeof(m), cudaMemcpyHostToDevice);
    float* dout;
    cudaMalloc(&dout, BR * BC * 4);
    cudaMemset(dout, 0xff, BR * BC * 4);
    k_probe<BR, BC><<<1, 32>>>(dm, m, use_param, c0, c1, c2, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-40s KERNEL ERROR: %s\n", name, cudaGetErrorString(e)); return 2; }
    std::vector<float> o(BR * BC);
    cudaMemcpy(o.data(), dout, BR * BC * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int j = 0; j < BC; ++j)
        for (int i = 0; i < BR; ++i) {
            const int y = c0 + i, x = c1 + j;
            float want = 0.f;
            if (y >= 0 && y < rows && x >= 0 && x < cols && c2 >= 0 && c2 < slots) want = host[(size_t)c2 * frame_elems + (size_t)x * rows + y];
            if (o[j * BR + i] != want) ++bad;
        }
    printf("%-40s ok, mismatches %d of %d\n", name, bad, BR * BC);
    cudaFree(dm); cudaFree(dout);
    return bad ? 3 : 0;
}

int main(int argc, char** argv) {
    const int which = argc > 1 ? atoi(argv[1]) : -1;  // one case per process: a faulting kernel poisons the context
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no entry point\n"); return 1; }
    EncodeTiledFn enc = (EncodeTiledFn)p;
    const int rows = 76, cols = 132, slots = 3;
    const size_t frame_elems = ((size_t)rows * cols + 31) / 32 * 32 + 64;
    std::vector<float> host(frame_elems * slots);
    for (size_t i = 0; i < host.size(); ++i) host[i] = (float)(i % 100003) * 0.5f + 1.f;
    float* d;
    cudaMalloc(&d, host.size() * 4);
    cudaMemcpy(d, host.data(), host.size() * 4, cudaMemcpyHostToDevice);
    int rc = 0;
    if (which == 0 || which < 0) rc |= run<40, 29>(enc, "param desc, box 40x29 interior", d, rows, cols, slots, frame_elems, 1, 10, 20, 1, host);
    if (which == 1 || which < 0) rc |= run<40, 29>(enc, "gmem  desc, box 40x29 interior", d, rows, cols, slots, frame_elems, 0, 10, 20, 1, host);
    if (which == 2 || which < 0) rc |= run<40, 29>(enc, "gmem  desc, box 40x29 negative coords", d, rows, cols, slots, frame_elems, 0, -7, -3, 2, host);
    if (which == 3 || which < 0) rc |= run<40, 29>(enc, "gmem  desc, box 40x29 past the end", d, rows, cols, slots, frame_elems, 0, 60, 120, 0, host);
    if (which == 4 || which < 0) rc |= run<24, 20>(enc, "gmem  desc, box 24x20 unaligned", d, rows, cols, slots, frame_elems, 0, 13, 17, 1, host);
    if (which == 5 || which < 0) rc |= run<48, 19>(enc, "gmem  desc, box 48x19", d, rows, cols, slots, frame_elems, 0, 5, 7, 1, host);
    // box taller than the tensor
    if (which == 6 || which < 0) rc |= run<40, 29>(enc, "gmem  desc, box 40 > 20 rows", d, 20, 34, slots, 20 * 34 + 40, 0, -3, 2, 1, host);
    if (which == 7 || which < 0) rc |= run<40, 29>(enc, "param desc, box 40 > 20 rows", d, 20, 34, slots, 20 * 34 + 40, 1, -3, 2, 1, host);
    if (which == 8) rc |= run<40, 29>(enc, "gmem desc, c0=12 slot 1", d, rows, cols, slots, frame_elems, 0, 12, 20, 1, host);
    if (which == 9) rc |= run<40, 29>(enc, "gmem desc, c0=10 slot 0", d, rows, cols, slots, frame_elems, 0, 10, 20, 0, host);
    if (which == 10) rc |= run<40, 29>(enc, "gmem desc, c0=-8 c1=-3 slot 2", d, rows, cols, slots, frame_elems, 0, -8, -3, 2, host);
    if (which == 11) rc |= run<40, 29>(enc, "param desc, c0=10 slot 0", d, rows, cols, slots, frame_elems, 1, 10, 20, 0, host);
    printf("probe rc %d\n", rc);
    return 0;
}
