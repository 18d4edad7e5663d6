// Measurement kernels behind ccqp_fp64_peak() / ccqp_microbench(): the denominators and latencies the
// batched solver kernel is judged and modelled with ("measure a DFMA peak
// micro-benchmark first").  Nothing here is on the product path.
//
//   fp64_peak_kernel : every thread runs 8 independent DFMA chains; with enough warps per SM this is
//                      the FP64 pipe's throughput (the fp64 roofline term of the batched mode),
//                      with 12 warps per SM it is what the batched kernel's occupancy can issue.
//   probe_kernel     : one warp (or one 64-thread CTA), clock64() around a dependent chain of one
//                      instruction kind: cycles per DFMA / DADD / DMUL / SHFL.64+DADD / IEEE division /
//                      sqrt / LDS.128 / STS+bar+LDS / bar.sync.
#pragma once
#include "common.cuh"

namespace ccqp {

__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double x, double y) {
    double a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = (double)(threadIdx.x + j) * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = fma(a[j], x, y);
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j];
    if (s == 123.456) out[0] = s;          // never true in practice; keeps the chains alive
}
constexpr int kPeakFmaPerIter = 64;         // per thread per loop trip

enum ProbeKind : int {
    PROBE_DFMA = 0, PROBE_DADD, PROBE_DMUL, PROBE_SHFL_DADD, PROBE_DIV, PROBE_SQRT, PROBE_LDS128_BCAST, PROBE_LDS128_DISTINCT,
    PROBE_STS_BAR_LDS, PROBE_BAR, PROBE_DSETP_SEL, PROBE_DMMA, PROBE_DMMA_X8, PROBE_WARPSUM_DMMA, PROBE_WARPSUM_SHFL,
    PROBE_DMMA_X8_DFMA_X8, PROBE_DFMA_X8_REUSE, PROBE_DFMA_X8_2RF, PROBE_DFMA_MATVEC64, PROBE_COUNT
};

// D(8x8) = A(8x4) B(4x8) + C on the FP64 tensor path (SASS DMMA.884): lane l holds A[l/4][l%4], B[l%4][l/4],
// C/D[l/4][2(l%4)], [2(l%4)+1]
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b, double c0, double c1) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                 : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
// sum of v over the 32 lanes of a warp in 2 DMMAs + 1 DADD (A = ones): quad sums, pair-of-quad sums, total
__device__ __forceinline__ double warp_sum_dmma(double v) {
    double d0, d1, e0, e1;
    dmma884(d0, d1, 1.0, v, 0.0, 0.0);
    dmma884(e0, e1, 1.0, d0 + d1, 0.0, 0.0);
    return e0;
}

// cycles[kind] = clock64() ticks of `reps` dependent operations executed by warp 0 of a 64-thread CTA
__global__ void __launch_bounds__(64) probe_kernel(long long* cycles, double* sink, int reps, double x, double y) {
    __shared__ __align__(16) double sh[128];
    const int t = threadIdx.x;
    sh[t] = x + t; sh[64 + t] = y;
    __syncthreads();
    double v = x + 1e-9 * t, w = y;
    long long t0, t1;
    auto tick = [] { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)::"memory"); return c; };

#define PROBE(kind, body)                                      \
    __syncthreads();                                           \
    t0 = tick();                                               \
    for (int r = 0; r < reps; ++r) { body; }                   \
    t1 = tick();                                               \
    if (t == 0) cycles[kind] = t1 - t0;                        \
    __syncthreads();

    PROBE(PROBE_DFMA, v = fma(v, w, w))
    PROBE(PROBE_DADD, v = v + w)
    PROBE(PROBE_DMUL, v = v * w)
    PROBE(PROBE_SHFL_DADD, v = v + __shfl_xor_sync(0xffffffffu, v, 1))
    v = fabs(v) + 1.0;
    PROBE(PROBE_DIV, v = w / v + 1.5)
    PROBE(PROBE_SQRT, v = sqrt(v) + 2.0)
    {
        const uint32_t base = smem_u32(sh);
        double a, b = 0.0;
        PROBE(PROBE_LDS128_BCAST, { const uint32_t ad = base + ((__double2loint(v) & 1) << 4);
                                     asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "r"(ad) : "memory"); v += a; })
        PROBE(PROBE_LDS128_DISTINCT, { const uint32_t ad = base + (((t & 31) + (__double2loint(v) & 1)) << 4);
                                        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "r"(ad) : "memory"); v += a; })
    }
    PROBE(PROBE_STS_BAR_LDS, { sh[t] = v; __syncthreads(); v = sh[(t + 1) & 63] + 1.0; __syncthreads(); })
    PROBE(PROBE_BAR, __syncthreads())
    PROBE(PROBE_DSETP_SEL, v = (v < w) ? w : v + 1.0)
    double keep = v;                          // every probe's result reaches the sink (nothing is dead code)
    v = 1.0 + 1e-9 * t;
    { double d0, d1; PROBE(PROBE_DMMA, { dmma884(d0, d1, 0.25, v, 0.0, 0.0); v = d0; }) }
    {
        double d[8][2];
        PROBE(PROBE_DMMA_X8, {
_Pragma("unroll")
            for (int u = 0; u < 8; ++u) dmma884(d[u][0], d[u][1], 0.25, v + u, 0.0, 0.0);
            v = d[0][0] * 1e-3;
_Pragma("unroll")
            for (int u = 1; u < 8; ++u) v += d[u][1] * 1e-9;
        })
    }
    keep += v;
    v = 1.0 + 1e-9 * t;
    PROBE(PROBE_WARPSUM_DMMA, v = warp_sum_dmma(v) * 0.03125)
    PROBE(PROBE_WARPSUM_SHFL, {
        v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2); v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16); v *= 0.03125; })
    {
        double d[8][2], f[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) f[u] = v + u;
        PROBE(PROBE_DMMA_X8_DFMA_X8, {          // do DMMA and DFMA share an issue path?  8 + 8 independent per trip
_Pragma("unroll")
            for (int u = 0; u < 8; ++u) { dmma884(d[u][0], d[u][1], 0.25, f[u], 0.0, 0.0); f[u] = fma(f[u], w, w); }
            v = d[0][0] * 1e-3;
_Pragma("unroll")
            for (int u = 1; u < 8; ++u) v += d[u][1] * 1e-9;
            f[0] += v * 1e-9;
        })
#pragma unroll
        for (int u = 0; u < 8; ++u) keep += f[u];
    }
    // Issue rate of INDEPENDENT DFMAs as a function of where the operands come from (one warp per SM sub-partition):
    //   x8_reuse  : f[u] = fma(f[u], w, w)         one fresh register operand per instruction (fp64_peak_kernel's pattern)
    //   x8_2rf    : f[u] = fma(g[u], w, f[u])      two (multiplicand and accumulator), 8 + 8 registers
    //   matvec64  : s[i] = fma(a[i][j], x[j], s[i]) the register-resident 8 x 8 block of the batched kernels: two fresh operands, the
    //               multiplicand never repeats within a trip (64 registers)
    {
        double f[8], g8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { f[u] = v * 1e-3 + u; g8[u] = w + 1e-3 * u; }
        PROBE(PROBE_DFMA_X8_REUSE, {
_Pragma("unroll")
            for (int u = 0; u < 8; ++u) f[u] = fma(f[u], w, w); })
        PROBE(PROBE_DFMA_X8_2RF, {
_Pragma("unroll")
            for (int u = 0; u < 8; ++u) f[u] = fma(g8[u], w, f[u]); })
        double a[8][8], xx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            xx[i] = 1e-6 * (w + i);
#pragma unroll
            for (int j = 0; j < 8; ++j) a[i][j] = 1e-3 * (x + i) + 1e-6 * (t + j);
        }
        PROBE(PROBE_DFMA_MATVEC64, {
_Pragma("unroll")
            for (int j = 0; j < 8; ++j)
_Pragma("unroll")
                for (int i = 0; i < 8; ++i) f[i] = fma(a[i][j], xx[j], f[i]); })
#pragma unroll
        for (int u = 0; u < 8; ++u) keep += f[u] + g8[u];
    }
    v += keep;
#undef PROBE
    sink[t] = v;
}

}  // namespace ccqp
