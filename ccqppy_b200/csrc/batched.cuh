// Batched mode: thousands of small independent box-constrained QPs (n <= 64), one CTA per problem
// at a time, the whole solver loop of the reference inside one persistent kernel (no host round
// trips).  Problem i is CCQPSolverX(tol,max_mv).solve(A[i], b[i], x0[i], BoxProjOp(n, lb[i], ub[i]))
// (solvers.py:94/220/393/583/719/878 with solution_spaces.py:280-366).
//
// Data path per problem (n = 64: 32 KB of A, 2 KB of vectors in, 512 B out):
//   HBM --TMA row copies (cp.async.bulk, 64 x 512 B, one mbarrier)--> padded shared-memory tile
//       --LDS.128--> registers: thread t keeps row t of A (64 doubles) for the whole solve.
//   The tile is free as soon as the rows are in registers, so the NEXT problem's copy is issued
//   immediately and lands while the current problem iterates (single buffer, full overlap).
//   Each mat-vec is 64 FMAs per thread against x broadcast from shared memory (LDS.128, one
//   wavefront per warp).  Dot products: warp shuffle tree, then the two warps exchange through
//   shared memory.  Everything else (projection, axpy, step lengths, stopping tests) is per-thread
//   register arithmetic.  Problems are handed out through an atomic counter (iteration counts
//   differ per problem); results do not depend on the schedule.
//
// Bounds: HBM bytes/problem = 8 n^2 + 32 n (+8 n for x0, + uniforms read by SPG);
//         fp64 flops/problem = 2 n^2 * (mat-vecs executed).
#pragma once
#include <functional>
#include <string>

#include "common.cuh"
#include "../../include/ccqp_b200.h"

namespace ccqp {

constexpr int kBN = 64;                 // max unknowns per problem = threads per CTA
constexpr int kBStride = kBN + 2;       // padded tile row (doubles): 528 B, conflict-free LDS.128
constexpr int kBWindow = 64;

struct BatchedOut {
    double residual;
    int mv, gemv, iters, draws;
    int converged, status;
};

struct BatchedCtx {
    const double* A;        // [batch][n][n]
    const double* b;        // [batch][n]
    const double* x0;       // [batch][n] or null
    const double* lb;
    const double* ub;
    const double* uniforms; // [batch][n_uniforms]
    long long n_uniforms;
    double* x_out;          // [batch][n]
    BatchedOut* out;        // [batch]
    unsigned* counter;      // work queue head
    int batch, n;
    int tma_ok;             // rows can be moved with 16-byte bulk copies
    double tol, max_mv, step, tau, sig1, sig2;
    int m;
};

struct BatchedSmem {
    double tile[kBN * kBStride];
    double xs[kBN];
    double red[2][2][4];    // [parity][warp][slot]
    uint64_t mbar;
    int next;
};

// sum of up to 4 values over the 64 threads; result in every thread; ONE __syncthreads
template <int K>
__device__ __forceinline__ void cta64_sum(double (&a)[K], BatchedSmem& sm, int& parity) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) a[k] = warp_sum(a[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) sm.red[parity][warp][k] = a[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) a[k] = sm.red[parity][0][k] + sm.red[parity][1][k];
    parity ^= 1;
}

// y_t = sum_j a[j] * xs[j]   (xs complete in shared memory; 4 interleaved FMA chains)
__device__ __forceinline__ double row_dot(const double (&a)[kBN], const double* xs) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int j = 0; j < kBN; j += 4) {
        const double2 u = *reinterpret_cast<const double2*>(xs + j);
        const double2 v = *reinterpret_cast<const double2*>(xs + j + 2);
        s0 = fma(a[j], u.x, s0);
        s1 = fma(a[j + 1], u.y, s1);
        s2 = fma(a[j + 2], v.x, s2);
        s3 = fma(a[j + 3], v.y, s3);
    }
    return (s0 + s1) + (s2 + s3);
}

__device__ __forceinline__ double clampd(double t, double lo, double hi) { return t < lo ? lo : (t > hi ? hi : t); }

// publish v as the mat-vec input and return (A v)_t
__device__ __forceinline__ double matvec(const double (&a)[kBN], BatchedSmem& sm, double v) {
    __syncthreads();            // previous readers of xs are done
    sm.xs[threadIdx.x] = v;
    __syncthreads();
    return row_dot(a, sm.xs);
}

struct BState {                 // per-thread view of one problem (thread t <-> unknown t)
    double b, lo, hi, x0;
    double cs;                  // 1/(3 n gd)
    bool act;                   // t < n
};

template <int SOLVER>
__device__ __forceinline__ void solve_one(const BatchedCtx& c, const double (&a)[kBN], BatchedSmem& sm, const BState& s,
                                          const double* uni, double& xsol, BatchedOut& o) {
    int par = 0;
    int mv = 0, gemv = 0, iters = 0, draws = 0, status = 0;
    double res = NAN;
    const double tol = c.tol, maxmv = c.max_mv;
    auto resid2 = [&](double x, double g) { const double d = s.cs * (x - clampd(x - kGd * g, s.lo, s.hi)); return d * d; };

    if constexpr (SOLVER == CCQP_SOLVER_PGD || SOLVER == CCQP_SOLVER_BBPGD || SOLVER == CCQP_SOLVER_BBPGDF) {
        // solvers.py:114-170, 606-669, 741-819
        double x = s.x0, xm = s.x0, g, gm, xmin = s.x0, gmin = s.x0, resmin = INFINITY;
        gm = matvec(a, sm, xm) + s.b; gemv++; mv = 1;
        double r1[1] = {resid2(xm, gm)};
        cta64_sum<1>(r1, sm, par);
        res = sqrt(r1[0]);
        if (res >= tol) {
            double step = c.step;
            if (SOLVER != CCQP_SOLVER_PGD) {
                const double ag = matvec(a, sm, gm); gemv++;          // not counted (:635)
                double q[2] = {gm * gm, gm * ag};
                cta64_sum<2>(q, sm, par);
                step = q[0] / q[1];
            }
            for (;;) {
                x = clampd(xm - step * gm, s.lo, s.hi);
                g = matvec(a, sm, x) + s.b; gemv++; mv++;
                if ((double)mv >= maxmv) break;
                const double sx = x - xm, sy = g - gm;
                double q[3] = {resid2(x, g), sx * sx, sx * sy};
                cta64_sum<3>(q, sm, par);
                res = sqrt(q[0]);
                iters++;
                if (res < tol) break;
                if (SOLVER == CCQP_SOLVER_BBPGDF) {                   // :793-800
                    if (res < resmin) { resmin = res; xmin = x; gmin = g; }
                    if (step < 10 * kEps) {
                        x = clampd(xmin - kGd * gmin, s.lo, s.hi);
                        const double sx2 = x - xm;
                        double q2[2] = {sx2 * sx2, sx2 * sy};
                        cta64_sum<2>(q2, sm, par);
                        q[1] = q2[0]; q[2] = q2[1];
                    }
                }
                if (SOLVER != CCQP_SOLVER_PGD) step = q[1] / (q[2] + 10 * kEps);
                xm = x; gm = g;
            }
        }
        xsol = x;
    } else if constexpr (SOLVER == CCQP_SOLVER_SPG) {
        // solvers.py:906-975
        double x = s.x0;
        double g = matvec(a, sm, x) + s.b; gemv++;
        const double ag = matvec(a, sm, g); gemv++;
        double q0[3] = {g * x, g * g, g * ag};
        cta64_sum<3>(q0, sm, par);
        double f = q0[0];
        double alpha = q0[1] / q0[2];
        mv = 2;
        double window[kBWindow];
        int wcount = 1, whead = 0;
        window[0] = f;
        double dd_rep = NAN;
        for (;;) {
            const double d = clampd(x - alpha * g, s.lo, s.hi) - x;
            const double ad = matvec(a, sm, d); gemv++; mv++;
            if ((double)mv >= maxmv) break;
            double q[3] = {d * d, d * ad, d * g};
            cta64_sum<3>(q, sm, par);
            const double dd = q[0], dAd = q[1], dg = q[2];
            dd_rep = dd;
            if (sqrt(dd) <= tol) break;
            double fmax = window[0];
            for (int j = 1; j < wcount; ++j) fmax = fmax > window[j] ? fmax : window[j];
            const double xi = (fmax - f) / dAd;
            const double beta = -dg / dAd;
            const double bhat = c.tau * beta + sqrt((c.tau * c.tau) * (beta * beta) + 2 * xi);
            const double hi = (c.sig2 < bhat) ? c.sig2 : bhat;        // Python min(bhat, sig2)
            if (hi != hi) { status = CCQP_ERR_RANGE; break; }
            if (draws >= c.n_uniforms) { status = CCQP_ERR_UNIFORMS_EXHAUSTED; break; }
            const double bk = c.sig1 + (hi - c.sig1) * uni[draws];
            draws++;
            x += bk * d;
            g += bk * ad;
            f += bk * bk * dg + 0.5 * (bk * bk) * dAd;                // :963 as written
            if (wcount < c.m) window[wcount++] = f;
            else { window[whead] = f; whead = (whead + 1) % c.m; }
            alpha = dd / dAd;
            iters++;
        }
        res = sqrt(dd_rep);
        xsol = x;
    } else if constexpr (SOLVER == CCQP_SOLVER_APGD || SOLVER == CCQP_SOLVER_APGD_AR) {
        // solvers.py:242-343, 415-533
        constexpr bool AR = SOLVER == CCQP_SOLVER_APGD_AR;
        double x = s.x0, y = s.x0, xp = s.x0, xhat = 1.0, axp = 0.0;
        const double d0 = s.act ? (s.x0 - 1.0) : 0.0;
        const double ad0 = matvec(a, sm, d0); gemv++; mv = 1;
        double q0[2] = {ad0 * ad0, d0 * d0};
        cta64_sum<2>(q0, sm, par);
        double L = sqrt(q0[0]) / sqrt(q0[1]);
        double t = 1.0 / L, theta = 1.0, resmin = INFINITY;
        for (;;) {
            const double ay = matvec(a, sm, y); gemv++; mv++;
            if ((double)mv >= maxmv) break;
            const double g = ay + s.b;
            xp = clampd(y - t * g, s.lo, s.hi);
            double r12[2] = {y * ay, y * s.b};
            bool have12 = false;
            double rt1 = 0.0, rt2 = 0.0;
            for (;;) {
                axp = matvec(a, sm, xp); gemv++; mv++;
                const bool lim = (double)mv >= maxmv;
                const double df = xp - y;
                if (!have12) {       // fold the two outer sums into the first inner reduction
                    double q[6] = {xp * axp, xp * s.b, g * df, df * df, r12[0], r12[1]};
                    // 6 slots: two rounds of <=4
                    double qa[4] = {q[0], q[1], q[2], q[3]};
                    double qb[2] = {q[4], q[5]};
                    cta64_sum<4>(qa, sm, par);
                    cta64_sum<2>(qb, sm, par);
                    rt1 = qb[0] * 0.5; rt2 = qb[1];
                    have12 = true;
                    if (lim) break;
                    if ((qa[0] * 0.5 + qa[1]) <= (rt1 + rt2 + qa[2] + 0.5 * L * qa[3])) break;
                } else {
                    double qa[4] = {xp * axp, xp * s.b, g * df, df * df};
                    cta64_sum<4>(qa, sm, par);
                    if (lim) break;
                    if ((qa[0] * 0.5 + qa[1]) <= (rt1 + rt2 + qa[2] + 0.5 * L * qa[3])) break;
                }
                L *= 2;
                t = 1.0 / L;
                xp = clampd(y - t * g, s.lo, s.hi);
            }
            double theta_n = 0.5 * (-theta * theta + theta * sqrt(4 + theta * theta));
            const double beta = theta * (1 - theta) / (theta * theta + theta_n);
            double yn = (1 + beta) * xp - beta * x;
            double q[2] = {resid2(xp, axp + s.b), AR ? g * (xp - x) : 0.0};
            cta64_sum<2>(q, sm, par);
            res = sqrt(q[0]);
            iters++;
            if (AR && res < resmin) { resmin = res; xhat = xp; }
            if (res < tol) break;
            if (AR && q[1] > 0) { yn = xp; theta_n = 1; }
            L *= 0.9;
            t = 1.0 / L;
            y = yn;
            { const double tmp = x; x = xp; xp = tmp; }   // buffer swap (:332-334)
            theta = theta_n;
        }
        xsol = AR ? xhat : xp;
    }
    o.residual = res;
    o.mv = mv; o.gemv = gemv; o.iters = iters; o.draws = draws;
    o.converged = ((double)mv < maxmv) ? 1 : 0;
    o.status = status;
}

template <int SOLVER>
__global__ void __launch_bounds__(kBN, 4) batched_kernel(const BatchedCtx c) {
    __shared__ __align__(128) BatchedSmem sm;
    const int t = threadIdx.x, n = c.n;
    unsigned phase = 0;
    if (t == 0) { mbar_init(&sm.mbar, 1); mbar_fence_init(); }
    __syncthreads();

    auto issue_load = [&](int prob) {   // rows of A[prob] -> padded tile
        if (c.tma_ok) {
            if (t == 0) mbar_expect_tx(&sm.mbar, (uint32_t)(n * n * 8));
            if (t < n) bulk_g2s(sm.tile + t * kBStride, c.A + ((size_t)prob * n + t) * n, (uint32_t)(n * 8), &sm.mbar);
        }
    };
    int cur;
    if (t == 0) sm.next = (int)atomicAdd(c.counter, 1u);
    __syncthreads();
    cur = sm.next;
    if (cur < c.batch) { fence_proxy_async(); issue_load(cur); }
    while (cur < c.batch) {
        double a[kBN];
        if (c.tma_ok) {
            mbar_wait(&sm.mbar, phase);
            phase ^= 1u;
        } else {
            __syncthreads();
            const double* Ap = c.A + (size_t)cur * n * n;
            for (int idx = t; idx < n * n; idx += kBN) sm.tile[(idx / n) * kBStride + (idx % n)] = Ap[idx];
            __syncthreads();
        }
#pragma unroll
        for (int j = 0; j < kBN; j += 2) {
            if (t < n && j + 1 < n) {
                const double2 v = *reinterpret_cast<const double2*>(sm.tile + t * kBStride + j);
                a[j] = v.x; a[j + 1] = v.y;
            } else if (t < n && j < n) {
                a[j] = sm.tile[t * kBStride + j]; a[j + 1] = 0.0;
            } else { a[j] = 0.0; a[j + 1] = 0.0; }
        }
        BState s;
        s.act = t < n;
        const size_t vo = (size_t)cur * n + t;
        s.b = s.act ? c.b[vo] : 0.0;
        s.lo = s.act ? c.lb[vo] : 0.0;
        s.hi = s.act ? c.ub[vo] : 0.0;
        s.x0 = (s.act && c.x0) ? c.x0[vo] : 0.0;
        s.cs = 1.0 / (3 * (double)n * kGd);
        __syncthreads();                       // every thread has its row: the tile is free
        if (t == 0) sm.next = (int)atomicAdd(c.counter, 1u);
        __syncthreads();
        const int nxt = sm.next;
        if (nxt < c.batch) { fence_proxy_async(); issue_load(nxt); }   // lands while we iterate

        double xsol = 0.0;
        BatchedOut o;
        solve_one<SOLVER>(c, a, sm, s, c.uniforms ? c.uniforms + (size_t)cur * c.n_uniforms : nullptr, xsol, o);
        if (s.act) c.x_out[vo] = xsol;
        if (t == 0) c.out[cur] = o;
        cur = nxt;
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <int SOLVER>
inline cudaError_t launch_batched(const BatchedCtx& c, int sm_count, cudaStream_t stream) {
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, batched_kernel<SOLVER>, kBN, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)sm_count * per_sm;
    if (grid > c.batch) grid = c.batch;
    batched_kernel<SOLVER><<<(unsigned)grid, kBN, 0, stream>>>(c);
    return cudaGetLastError();
}

// Returns a ccqp_status.  `alloc(bytes)` returns a device workspace of at least that size.
inline int batched_solve(cudaStream_t stream, int sm_count, void*, size_t, int solver, const ccqp_params& prm,
                         long long batch, long long n, const double* A, const double* b, const double* x0,
                         const double* lb, const double* ub, const double* uniforms, long long n_uniforms, double* x_out,
                         int memtype, ccqp_result* results, ccqp_result* summary, cudaEvent_t ev0, cudaEvent_t ev1,
                         int* launches, std::string& err, const std::function<void*(size_t)>& alloc) {
#define BCU(call)                                                                                  \
    do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e__); return CCQP_ERR_CUDA; } } while (0)
    if (n > kBN) return CCQP_ERR_UNSUPPORTED;
    if (solver == CCQP_SOLVER_MPRGP) return CCQP_ERR_UNSUPPORTED;
    if (batch >= (1LL << 31) - 1024) return CCQP_ERR_INVALID_ARG;
    const bool host = memtype == CCQP_MEM_HOST;
    if (solver == CCQP_SOLVER_SPG && (n_uniforms < 0 || (n_uniforms > 0 && !uniforms))) return CCQP_ERR_INVALID_ARG;
    const size_t szA = (size_t)batch * n * n * 8, szV = (size_t)batch * n * 8;
    const size_t szU = (solver == CCQP_SOLVER_SPG) ? (size_t)batch * n_uniforms * 8 : 0;
    const size_t szO = (size_t)batch * sizeof(BatchedOut);
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    size_t total = 256 + al(szO);
    if (host) total += al(szA) + 4 * al(szV) + al(szV) + al(szU);
    unsigned char* ws = static_cast<unsigned char*>(alloc(total));
    if (!ws) { err = "workspace allocation failed"; return CCQP_ERR_CUDA; }
    size_t off = 0;
    auto take = [&](size_t bytes) { unsigned char* p = ws + off; off += al(bytes); return p; };
    unsigned* counter = reinterpret_cast<unsigned*>(take(256));
    BatchedOut* dout = reinterpret_cast<BatchedOut*>(take(szO));
    BatchedCtx c;
    std::memset(&c, 0, sizeof(c));
    if (host) {
        double* dA = reinterpret_cast<double*>(take(szA));
        double* db = reinterpret_cast<double*>(take(szV));
        double* dlb = reinterpret_cast<double*>(take(szV));
        double* dub = reinterpret_cast<double*>(take(szV));
        double* dx0 = reinterpret_cast<double*>(take(szV));
        double* dxo = reinterpret_cast<double*>(take(szV));
        double* du = szU ? reinterpret_cast<double*>(take(szU)) : nullptr;
        BCU(cudaMemcpyAsync(dA, A, szA, cudaMemcpyHostToDevice, stream));
        BCU(cudaMemcpyAsync(db, b, szV, cudaMemcpyHostToDevice, stream));
        BCU(cudaMemcpyAsync(dlb, lb, szV, cudaMemcpyHostToDevice, stream));
        BCU(cudaMemcpyAsync(dub, ub, szV, cudaMemcpyHostToDevice, stream));
        if (x0) BCU(cudaMemcpyAsync(dx0, x0, szV, cudaMemcpyHostToDevice, stream));
        if (du) BCU(cudaMemcpyAsync(du, uniforms, szU, cudaMemcpyHostToDevice, stream));
        c.A = dA; c.b = db; c.lb = dlb; c.ub = dub; c.x0 = x0 ? dx0 : nullptr; c.uniforms = du; c.x_out = dxo;
    } else {
        c.A = A; c.b = b; c.lb = lb; c.ub = ub; c.x0 = x0; c.uniforms = (solver == CCQP_SOLVER_SPG) ? uniforms : nullptr;
        c.x_out = x_out;
    }
    c.n_uniforms = (solver == CCQP_SOLVER_SPG) ? n_uniforms : 0;
    c.out = dout; c.counter = counter;
    c.batch = (int)batch; c.n = (int)n;
    c.tma_ok = ((n * 8) % 16 == 0 && (reinterpret_cast<uintptr_t>(c.A) & 15) == 0) ? 1 : 0;
    c.tol = prm.tol; c.max_mv = prm.max_mv; c.step = prm.step_size;
    c.tau = prm.tau; c.sig1 = prm.sigma1; c.sig2 = prm.sigma2; c.m = prm.m;
    BCU(cudaMemsetAsync(counter, 0, 256, stream));
    BCU(cudaEventRecord(ev0, stream));
    cudaError_t le;
    switch (solver) {
        case CCQP_SOLVER_PGD: le = launch_batched<CCQP_SOLVER_PGD>(c, sm_count, stream); break;
        case CCQP_SOLVER_APGD: le = launch_batched<CCQP_SOLVER_APGD>(c, sm_count, stream); break;
        case CCQP_SOLVER_APGD_AR: le = launch_batched<CCQP_SOLVER_APGD_AR>(c, sm_count, stream); break;
        case CCQP_SOLVER_BBPGD: le = launch_batched<CCQP_SOLVER_BBPGD>(c, sm_count, stream); break;
        case CCQP_SOLVER_BBPGDF: le = launch_batched<CCQP_SOLVER_BBPGDF>(c, sm_count, stream); break;
        case CCQP_SOLVER_SPG: le = launch_batched<CCQP_SOLVER_SPG>(c, sm_count, stream); break;
        default: return CCQP_ERR_INVALID_ARG;
    }
    BCU(le);
    *launches = 1;
    BCU(cudaEventRecord(ev1, stream));
    if (host) BCU(cudaMemcpyAsync(x_out, c.x_out, szV, cudaMemcpyDeviceToHost, stream));
    std::vector<BatchedOut> hout((size_t)batch);
    BCU(cudaMemcpyAsync(hout.data(), dout, szO, cudaMemcpyDeviceToHost, stream));
    BCU(cudaStreamSynchronize(stream));
    float ms = 0.f;
    BCU(cudaEventElapsedTime(&ms, ev0, ev1));
    long long tot_mv = 0, tot_gemv = 0, tot_it = 0, tot_dr = 0, nconv = 0;
    int first_status = 0;
    for (long long i = 0; i < batch; ++i) {
        const BatchedOut& o = hout[(size_t)i];
        if (results) {
            ccqp_result& r = results[i];
            std::memset(&r, 0, sizeof(r));
            r.residual = o.residual; r.mv_count = o.mv; r.gemv_count = o.gemv; r.iterations = o.iters;
            r.uniforms_used = o.draws; r.converged = o.converged; r.status = o.status;
            r.hbm_bytes = 8.0 * n * n + 8.0 * n * (x0 ? 5 : 4);
        }
        tot_mv += o.mv; tot_gemv += o.gemv; tot_it += o.iters; tot_dr += o.draws; nconv += o.converged;
        if (o.status && !first_status) first_status = o.status;
    }
    if (summary) {
        std::memset(summary, 0, sizeof(*summary));
        summary->gpu_seconds = ms * 1e-3;
        summary->mv_count = tot_mv; summary->gemv_count = tot_gemv; summary->iterations = tot_it;
        summary->uniforms_used = tot_dr; summary->converged = (nconv == batch) ? 1 : 0;
        summary->status = first_status;
        summary->hbm_bytes = (double)batch * (8.0 * n * n + 8.0 * n * (x0 ? 5 : 4)) + 8.0 * (double)tot_dr;
        summary->kernel_launches = 1;
        summary->residual = NAN;
    }
    return CCQP_OK;
#undef BCU
}

}  // namespace ccqp
