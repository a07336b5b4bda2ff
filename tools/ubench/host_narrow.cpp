#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <thread>
#include <vector>
#include <immintrin.h>
// int64 -> int32, returns OR of all values (sign / range check by the caller)
__attribute__((target("avx2"))) static uint64_t narrow_avx2(int32_t *dst, const int64_t *src, size_t n)
{
    __m256i acc = _mm256_setzero_si256();
    const __m256i idx = _mm256_setr_epi32(0, 2, 4, 6, 0, 0, 0, 0);
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        __m256i a = _mm256_loadu_si256((const __m256i *)(src + i));
        __m256i b = _mm256_loadu_si256((const __m256i *)(src + i + 4));
        acc = _mm256_or_si256(acc, _mm256_or_si256(a, b));
        __m128i lo = _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(a, idx));
        __m128i hi = _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(b, idx));
        _mm_stream_si128((__m128i *)(dst + i), lo);
        _mm_stream_si128((__m128i *)(dst + i + 4), hi);
    }
    uint64_t o[4];
    _mm256_storeu_si256((__m256i *)o, acc);
    uint64_t r = o[0] | o[1] | o[2] | o[3];
    for (; i < n; ++i) { r |= (uint64_t)src[i]; dst[i] = (int32_t)src[i]; }
    return r;
}
int main(int argc, char **argv)
{
    const int T = argc > 1 ? atoi(argv[1]) : 8;
    const size_t n = (size_t)256 << 20 >> 3;   // 256 MB of int64
    int64_t *src = (int64_t *)aligned_alloc(64, n * 8);
    int32_t *dst = (int32_t *)aligned_alloc(64, n * 4);
    for (size_t i = 0; i < n; ++i) src[i] = i & 1023;
    memset(dst, 0, n * 4);
    for (int rep = 0; rep < 4; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        std::vector<uint64_t> ors(T);
        for (int t = 0; t < T; ++t) th.emplace_back([&, t] { size_t lo = n * t / T & ~(size_t)7, hi = t + 1 == T ? n : (n * (t + 1) / T & ~(size_t)7); ors[t] = narrow_avx2(dst + lo, src + lo, hi - lo); });
        for (auto &x : th) x.join();
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("threads %d: %.2f ms, %.1f GB/s of int64 read\n", T, s * 1e3, n * 8 / s / 1e9);
    }
    return 0;
}
