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
    PROBE_STS_BAR_LDS, PROBE_BAR, PROBE_DSETP_SEL, PROBE_COUNT
};

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
        double a, b;
        PROBE(PROBE_LDS128_BCAST, { const uint32_t ad = base + ((__double2loint(v) & 1) << 4);
                                     asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "r"(ad) : "memory"); v += a; })
        PROBE(PROBE_LDS128_DISTINCT, { const uint32_t ad = base + (((t & 31) + (__double2loint(v) & 1)) << 4);
                                        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "r"(ad) : "memory"); v += a; })
    }
    PROBE(PROBE_STS_BAR_LDS, { sh[t] = v; __syncthreads(); v = sh[(t + 1) & 63] + 1.0; __syncthreads(); })
    PROBE(PROBE_BAR, __syncthreads())
    PROBE(PROBE_DSETP_SEL, v = (v < w) ? w : v + 1.0)
#undef PROBE
    sink[t] = v;
}

}  // namespace ccqp
