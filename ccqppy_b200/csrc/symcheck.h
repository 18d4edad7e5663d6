// Host-side test behind the symmetric upload (upload.cu): does every entry below the block diagonal equal its mirror image?
// Pure C++ (no CUDA), so that it can be compiled and timed on its own.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include <sched.h>
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define CCQP_SYMCHECK_AVX2 1
#else
#define CCQP_SYMCHECK_AVX2 0
#endif

namespace ccqp {
namespace {

constexpr int kChkTile = 64;        // 64 x 64 doubles = 32 KB per tile

#if CCQP_SYMCHECK_AVX2
// One 4 x 4 block: rows of L against the transposed rows of U; returns the OR of (bit differences | NaN masks).
__attribute__((target("avx2"))) inline __m256d block4_diff(const double* Lp, const double* Up, long long lda) {
    const __m256d u0 = _mm256_loadu_pd(Up), u1 = _mm256_loadu_pd(Up + lda), u2 = _mm256_loadu_pd(Up + 2 * lda),
                  u3 = _mm256_loadu_pd(Up + 3 * lda);
    const __m256d t0 = _mm256_unpacklo_pd(u0, u1), t1 = _mm256_unpackhi_pd(u0, u1), t2 = _mm256_unpacklo_pd(u2, u3),
                  t3 = _mm256_unpackhi_pd(u2, u3);
    const __m256d c0 = _mm256_permute2f128_pd(t0, t2, 0x20), c1 = _mm256_permute2f128_pd(t1, t3, 0x20),
                  c2 = _mm256_permute2f128_pd(t0, t2, 0x31), c3 = _mm256_permute2f128_pd(t1, t3, 0x31);   // columns of U
    const __m256d l0 = _mm256_loadu_pd(Lp), l1 = _mm256_loadu_pd(Lp + lda), l2 = _mm256_loadu_pd(Lp + 2 * lda),
                  l3 = _mm256_loadu_pd(Lp + 3 * lda);
    __m256d d = _mm256_or_pd(_mm256_or_pd(_mm256_xor_pd(l0, c0), _mm256_xor_pd(l1, c1)),
                             _mm256_or_pd(_mm256_xor_pd(l2, c2), _mm256_xor_pd(l3, c3)));
    const __m256d nan = _mm256_or_pd(_mm256_or_pd(_mm256_cmp_pd(l0, l0, _CMP_UNORD_Q), _mm256_cmp_pd(l1, l1, _CMP_UNORD_Q)),
                                     _mm256_or_pd(_mm256_cmp_pd(l2, l2, _CMP_UNORD_Q), _mm256_cmp_pd(l3, l3, _CMP_UNORD_Q)));
    return _mm256_or_pd(d, nan);
}
// L = A[i0 .. i0+nc)[j0 .. j0+nr) against U = A[j0 .. j0+nr)[i0 .. i0+nc), nc and nr multiples of 4
__attribute__((target("avx2"))) inline bool tile_differs_avx2(const double* A, long long lda, long long i0, long long j0, int nc, int nr) {
    __m256d acc = _mm256_setzero_pd();
    for (int a0 = 0; a0 < nc; a0 += 4) {
        const double* Lrow = A + (i0 + a0) * lda + j0;
        const double* Ucol = A + j0 * lda + (i0 + a0);
        for (int b0 = 0; b0 < nr; b0 += 4) acc = _mm256_or_pd(acc, block4_diff(Lrow + b0, Ucol + (long long)b0 * lda, lda));
    }
    return !_mm256_testz_si256(_mm256_castpd_si256(acc), _mm256_castpd_si256(acc));
}
#endif

int host_threads_available() {
    cpu_set_t set;
    int c = 0;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) c = CPU_COUNT(&set);
    if (c <= 0) c = (int)std::thread::hardware_concurrency();
    return std::max(1, std::min(c, 64));
}

// true iff A[i][j] == A[j][i] (bitwise-equal values; NaN never is) for all pairs with j < blk * (i / blk); blk a multiple of kChkTile
bool host_lower_blocks_mirror_upper(const double* A, long long n, long long lda, int blk, int threads) {
    // work items: pairs of 64 x 64 tiles (ti, tj), tj's columns below the block start of ti's rows
    const long long nt = (n + kChkTile - 1) / kChkTile;
    std::atomic<long long> next(0);
    std::atomic<bool> differs(false);
#if CCQP_SYMCHECK_AVX2
    const char* scalar_env = getenv("CCQP_SYMCHECK_SCALAR");           // test hook: the portable path on an AVX2 machine
    const bool use_avx2 = __builtin_cpu_supports("avx2") && !(scalar_env && atoi(scalar_env) != 0);
#endif
    auto worker = [&]() {
        for (;;) {
            const long long ti = next.fetch_add(1, std::memory_order_relaxed);      // one row of tiles per grab
            if (ti >= nt || differs.load(std::memory_order_relaxed)) return;
            const long long i0 = ti * kChkTile, i1 = std::min(n, i0 + kChkTile);
            const long long jend = (i0 / blk) * blk;                       // columns [0, jend) are mirrored
            for (long long j0 = 0; j0 < jend; j0 += kChkTile) {
                const long long j1 = std::min(jend, j0 + kChkTile);
                const int nr = (int)(j1 - j0), nc = (int)(i1 - i0);
                // L = A[i0.., j0..] (nc x nr, below the diagonal) against U = A[j0.., i0..] (nr x nc), in 8 x 8 blocks: the 8 + 8 cache
                // lines of a block are each used completely and exactly once, so nothing has to stay cached and the power-of-two
                // row stride (every row of U in the same cache set) does not matter.
                // Equality is on bits (a -0.0 / +0.0 pair counts as different: the mirror would change the sign bit); a NaN mirrored
                // by the same NaN passes the bit test and must not: mx collects the largest |value| pattern, NaNs exceed +inf's.
#if CCQP_SYMCHECK_AVX2
                if (use_avx2 && !(nr & 3) && !(nc & 3)) {
                    if (tile_differs_avx2(A, lda, i0, j0, nc, nr)) { differs.store(true, std::memory_order_relaxed); return; }
                    continue;
                }
#endif
                unsigned long long diff = 0, mx = 0;
                const unsigned long long kAbs = 0x7fffffffffffffffULL, kInf = 0x7ff0000000000000ULL;
                for (int a0 = 0; a0 < nc; a0 += 8) {
                    const int na = std::min(8, nc - a0);
                    for (int b0 = 0; b0 < nr; b0 += 8) {
                        const int nb = std::min(8, nr - b0);
                        const double* Lp = A + (i0 + a0) * lda + (j0 + b0);
                        const double* Up = A + (j0 + b0) * lda + (i0 + a0);
                        if (na == 8 && nb == 8) {
                            unsigned long long l[8][8];
                            for (int a = 0; a < 8; ++a) std::memcpy(l[a], Lp + a * lda, 64);
                            for (int b = 0; b < 8; ++b) {
                                unsigned long long u[8];
                                std::memcpy(u, Up + b * lda, 64);
                                for (int a = 0; a < 8; ++a) { diff |= l[a][b] ^ u[a]; mx = std::max(mx, u[a] & kAbs); }
                            }
                        } else {
                            for (int a = 0; a < na; ++a)
                                for (int b = 0; b < nb; ++b) {
                                    unsigned long long x, y;
                                    std::memcpy(&x, Lp + a * lda + b, 8);
                                    std::memcpy(&y, Up + b * lda + a, 8);
                                    diff |= x ^ y; mx = std::max(mx, y & kAbs);
                                }
                        }
                    }
                }
                if (mx > kInf) diff = 1;
                if (diff) { differs.store(true, std::memory_order_relaxed); return; }
            }
        }
    };
    std::vector<std::thread> pool;
    const int T = (int)std::max<long long>(1, std::min<long long>(threads, nt));
    for (int t = 1; t < T; ++t) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
    return !differs.load();
}

}  // namespace
}  // namespace ccqp
