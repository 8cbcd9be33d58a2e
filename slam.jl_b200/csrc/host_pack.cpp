// Host-side lossless repacking of Float64 frames that hold 8-bit data (compiled by g++ directly: the AVX intrinsics headers are
// not for nvcc's front end).  See pack_u8_exact in api.cu for the contract.
#include <immintrin.h>
#include <stddef.h>
#include <stdint.h>

namespace sk {

static double g_table[256];
static bool g_table_ready = [] {
    for (int k = 0; k < 256; ++k) g_table[k] = (double)k / 255.0;
    return true;
}();

static bool pack_scalar(const double* src, uint8_t* dst, size_t n) {
    unsigned bad = 0;
    for (size_t i = 0; i < n; ++i) {
        const double v = src[i];
        int k = (int)(v * 255.0 + 0.5);
        k = k < 0 ? 0 : (k > 255 ? 255 : k);
        bad |= (g_table[k] != v);  // (NaN never equals: counted as bad)
        dst[i] = (uint8_t)k;
    }
    return bad == 0;
}

__attribute__((target("avx512f,avx512bw,avx512vl,avx512dq"))) static bool pack_avx512(const double* src, uint8_t* dst, size_t n) {
    const __m512d c255 = _mm512_set1_pd(255.0), half = _mm512_set1_pd(0.5);
    const __m256i lo = _mm256_setzero_si256(), hi = _mm256_set1_epi32(255);
    __mmask8 bad = 0;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        const __m512d v = _mm512_loadu_pd(src + i);
        __m256i k = _mm512_cvttpd_epi32(_mm512_fmadd_pd(v, c255, half));  // values are in [0, 1]: the fused rounding cannot move floor()
        k = _mm256_min_epi32(_mm256_max_epi32(k, lo), hi);
        const __m512d t = _mm512_i32gather_pd(k, g_table, 8);
        bad |= _mm512_cmp_pd_mask(t, v, _CMP_NEQ_UQ);
        _mm_storel_epi64((__m128i*)(dst + i), _mm256_cvtepi32_epi8(k));
    }
    return bad == 0 && pack_scalar(src + i, dst + i, n - i);
}

__attribute__((target("avx2,fma"))) static bool pack_avx2(const double* src, uint8_t* dst, size_t n) {
    const __m256d c255 = _mm256_set1_pd(255.0), half = _mm256_set1_pd(0.5);
    const __m128i lo = _mm_setzero_si128(), hi = _mm_set1_epi32(255);
    int bad = 0;
    size_t i = 0;
    for (; i + 4 <= n; i += 4) {
        const __m256d v = _mm256_loadu_pd(src + i);
        __m128i k = _mm256_cvttpd_epi32(_mm256_add_pd(_mm256_mul_pd(v, c255), half));
        k = _mm_min_epi32(_mm_max_epi32(k, lo), hi);
        const __m256d t = _mm256_i32gather_pd(g_table, k, 8);
        bad |= _mm256_movemask_pd(_mm256_cmp_pd(t, v, _CMP_NEQ_UQ));
        const __m128i b = _mm_shuffle_epi8(k, _mm_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1));
        *(int32_t*)(dst + i) = _mm_cvtsi128_si32(b);
    }
    return bad == 0 && pack_scalar(src + i, dst + i, n - i);
}

// dst[i] = k for src[i] == k/255 exactly; false at the first pixel (block) that is not such a value
bool pack_u8_exact_impl(const double* src, uint8_t* dst, size_t n) {
    (void)g_table_ready;
    static const int level = [] {
        __builtin_cpu_init();
        if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") &&
            __builtin_cpu_supports("avx512dq"))
            return 2;
        if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma")) return 1;
        return 0;
    }();
    return level == 2 ? pack_avx512(src, dst, n) : (level == 1 ? pack_avx2(src, dst, n) : pack_scalar(src, dst, n));
}

}  // namespace sk
