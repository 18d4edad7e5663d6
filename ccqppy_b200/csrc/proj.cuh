// Device form of the feasible set and the projection / normal-vector device functions.
//
// Host side (capi.cu) turns the C-ABI block table (one block == one leaf operator of the
// reference, solution_spaces.py:77-560) into
//   * per-element arrays lo[n], hi[n], ekind[n] for the elementwise kinds
//     (Identity/Lower/Upper/Box), so the hot path clamps without touching the block table;
//   * the list of "norm" blocks (Sphere / Cone / SOC), split into small ones (dim <= kSmallDim,
//     one thread per block, sequential FMA chain = the 3-element ddot order of the reference's
//     BLAS) and big ones (one CTA per block);
//   * the full block table, used by normal_vector, whose feasibility test is per leaf operator.
//
// A "projection phase" is expressed with two functors so it can be fused with whatever produces
// the argument and consumes the result:
//     arg(i)              -> t_i, the i-th entry of the vector being projected (pure, may be
//                            evaluated twice for norm blocks)
//     sink(i, t_i, p_i)   -> called exactly once per element with p = P(t)
#pragma once
#include "common.cuh"

namespace ccqp {

enum : int { kIdentity = 0, kLower = 1, kUpper = 2, kBox = 3, kSphere = 4, kConeRef = 5, kSoc = 6 };
enum : uint8_t { kElemNorm = 7 };   // ekind of elements owned by a norm block
constexpr int kSmallDim = 8;

struct ProjTable {
    int n;
    const double* lo;        // [n]  -inf where there is no lower bound
    const double* hi;        // [n]  +inf where there is no upper bound
    const uint8_t* ekind;    // [n]  kIdentity/kLower/kUpper/kBox or kElemNorm
    // every leaf block, in order
    int nblk;
    const int* bkind;
    const int* boff;
    const int* bdim;
    const double* bpar;      // first parameter of norm blocks (radius / mu); unused otherwise
    // norm blocks
    int nsmall;
    const int* small_ids;
    int nbig;
    const int* big_ids;
    int has_cone_ref;
    int all_elementwise;     // no norm blocks at all
    // element range and block ranges owned by this rank (multi-GPU row shard); single GPU: all
    int e0, e1;
};

__device__ __forceinline__ double clamp_elem(int kind, double t, double lo, double hi) {
    // compare/select form of solution_spaces.py:200-201, :276-277, :363-366 (value-identical
    // for finite inputs up to the sign of zero)
    if (kind == kLower) return t < lo ? lo : t;
    if (kind == kUpper) return t > hi ? hi : t;
    if (kind == kBox) return t < lo ? lo : (t > hi ? hi : t);
    return t;
}

// Scale rule of a norm block given its reductions.  Returns P(t)_j for entry j of the block.
//   sphere : r = |t|;                 r > R ? (R*t)/r : t                  solution_spaces.py:431-435
//   coneref: r = |t| (whole block);   see solution_spaces.py:484-492
//   soc    : tn = |t[:-1]|, z = t[-1] (extension)
struct NormRule {
    int kind;
    double par;
    double r;      // sphere/coneref: norm of whole block ; soc: norm of t[:-1]
    double last;   // t[dim-1]
    int mode;      // 0: identity, 1: zero, 2: scale
    double s;      // coneref / soc scalar
    __device__ __forceinline__ void finish() {
        if (kind == kSphere) {
            mode = (r > par) ? 2 : 0;
        } else if (kind == kConeRef) {
            if (par * last >= r) mode = 0;
            else if (-last / par >= r) mode = 1;
            else { mode = 2; s = (last + par * r) / (par * par + 1.0); }
        } else {   // soc
            if (r <= par * last) mode = 0;
            else if (par * r <= -last) mode = 1;
            else { mode = 2; s = (par * r + last) / (par * par + 1.0); }
        }
    }
    __device__ __forceinline__ double apply(double t, bool is_last) const {
        if (mode == 0) return t;
        if (mode == 1) return 0.0;
        if (kind == kSphere) return par * t / r;
        if (kind == kConeRef) return is_last ? s * (-par) : s * (t / r);
        return is_last ? s : (par * s) * t / r;
    }
};

// One projection pass over the elements this CTA is responsible for.
//   gtid/gstride : global thread id / total thread count of the grid
//   scratch      : 64 doubles of shared memory (only touched when big norm blocks exist)
// Every thread of every CTA must call it (it contains __syncthreads() when nbig > 0).
// kDoElem / kDoSmall = false skip the elementwise kinds / the small norm blocks (MPRGP's bisection
// handles those itself).
template <bool kDoElem = true, bool kDoSmall = true, class Arg, class Sink>
__device__ __forceinline__ void project_pass(const ProjTable& T, int gtid, int gstride, double* scratch,
                                             Arg arg, Sink sink) {
    if (kDoElem) for (int i = T.e0 + gtid; i < T.e1; i += gstride) {
        const int k = T.ekind[i];
        if (k == kElemNorm) continue;
        const double t = arg(i);
        sink(i, t, clamp_elem(k, t, T.lo[i], T.hi[i]));
    }
    if (T.all_elementwise) return;
    if (kDoSmall) for (int s = gtid; s < T.nsmall; s += gstride) {
        const int b = T.small_ids[s];
        const int off = T.boff[b], dim = T.bdim[b];
        if (off < T.e0 || off >= T.e1) continue;
        double t[kSmallDim];
        NormRule R;
        R.kind = T.bkind[b]; R.par = T.bpar[b];
        const int nsq = (R.kind == kSoc) ? dim - 1 : dim;
        double ss = 0.0;
#pragma unroll
        for (int j = 0; j < kSmallDim; ++j) {
            if (j < dim) {
                t[j] = arg(off + j);
                if (j < nsq) ss = (j == 0) ? t[j] * t[j] : fma(t[j], t[j], ss);
            }
        }
        R.r = sqrt(ss);
        R.last = 0.0;
#pragma unroll
        for (int j = 0; j < kSmallDim; ++j) if (j == dim - 1) R.last = t[j];
        R.finish();
#pragma unroll
        for (int j = 0; j < kSmallDim; ++j)
            if (j < dim) sink(off + j, t[j], R.apply(t[j], j == dim - 1));
    }
    // CTA index / CTA count of this rank's grid, derived from the thread ids the caller numbers the grid with
    for (int s = gtid / (int)blockDim.x; s < T.nbig; s += gstride / (int)blockDim.x) {
        const int b = T.big_ids[s];
        const int off = T.boff[b], dim = T.bdim[b];
        if (off < T.e0 || off >= T.e1) continue;   // CTA-uniform
        NormRule R;
        R.kind = T.bkind[b]; R.par = T.bpar[b];
        const int nsq = (R.kind == kSoc) ? dim - 1 : dim;
        // the last entry travels through the reduction too (exact: it is the only non-zero term),
        // so no thread re-reads an element after another thread's sink may have rewritten it
        double acc[2] = {0.0, 0.0};
        for (int j = threadIdx.x; j < dim; j += blockDim.x) {
            const double t = arg(off + j);
            if (j < nsq) acc[0] = fma(t, t, acc[0]);
            if (j == dim - 1) acc[1] = t;
        }
        cta_sum<2>(acc, scratch);
        R.r = sqrt(acc[0]);
        R.last = acc[1];
        R.finish();
        for (int j = threadIdx.x; j < dim; j += blockDim.x) {
            const double t = arg(off + j);
            sink(off + j, t, R.apply(t, j == dim - 1));
        }
    }
}

// Does this rank own block b (multi-GPU)?  Blocks never straddle a shard boundary: the host
// aligns shard boundaries to block boundaries (capi.cu) or refuses the table.
__device__ __forceinline__ bool owns_block(const ProjTable& T, int b) {
    return T.boff[b] >= T.e0 && T.boff[b] < T.e1;
}

// normal_vector(x) of the whole table into out[] (solution_spaces.py:146-160, :222-236, :306-322,
// :389-403, :512-525).  Per leaf operator: zero unless |x - P(x)| isclose 0 over the WHOLE leaf;
// then the kind's boundary rule.  One CTA per leaf with dim > kSmallDim, one thread otherwise.
// Only evaluated on MPRGP's rare infeasible-iterate path and by the ccqp_normal test hook, so
// it favours clarity over speed.  xv(i) reads x_i.
template <class XF>
__device__ void normal_pass(const ProjTable& T, int gtid, int gstride, double* scratch, XF xv, double* out) {
    // thread-per-leaf for small leaves
    for (int b = gtid; b < T.nblk; b += gstride) {
        const int dim = T.bdim[b];
        if (dim > kSmallDim || !owns_block(T, b)) continue;
        const int off = T.boff[b], kind = T.bkind[b];
        double x[kSmallDim], p[kSmallDim];
        NormRule R;
        R.kind = kind; R.par = T.bpar[b];
        const bool nb = kind >= kSphere;
        const int nsq = (kind == kSoc) ? dim - 1 : dim;
        double ss = 0.0;
        for (int j = 0; j < dim; ++j) {
            x[j] = xv(off + j);
            if (nb && j < nsq) ss = (j == 0) ? x[j] * x[j] : fma(x[j], x[j], ss);
        }
        if (nb) { R.r = sqrt(ss); R.last = x[dim - 1]; R.finish(); }
        double d2 = 0.0;
        for (int j = 0; j < dim; ++j) {
            p[j] = nb ? R.apply(x[j], j == dim - 1) : clamp_elem(kind, x[j], T.lo[off + j], T.hi[off + j]);
            const double d = x[j] - p[j];
            d2 = (j == 0) ? d * d : fma(d, d, d2);
        }
        const bool feas = is_close(sqrt(d2), 0.0);
        double pn2 = 0.0, pt2 = 0.0;
        for (int j = 0; j < dim; ++j) {
            pn2 = (j == 0) ? p[j] * p[j] : fma(p[j], p[j], pn2);
            if (j < dim - 1) pt2 = (j == 0) ? p[j] * p[j] : fma(p[j], p[j], pt2);
        }
        const double pn = sqrt(pn2), pt = sqrt(pt2);
        for (int j = 0; j < dim; ++j) {
            double v = 0.0;
            if (feas) {
                if (kind == kLower) v = is_close(p[j], T.lo[off + j]) ? -1.0 : 0.0;
                else if (kind == kUpper) v = is_close(p[j], T.hi[off + j]) ? 1.0 : 0.0;
                else if (kind == kBox) v = is_close(p[j], T.hi[off + j]) ? 1.0 : (is_close(p[j], T.lo[off + j]) ? -1.0 : 0.0);
                else if (kind == kSphere) v = is_close(pn, R.par) ? p[j] / pn : 0.0;
                else if (kind == kSoc) {
                    if (pt > 0.0 && is_close(pt, R.par * p[dim - 1]))
                        v = ((j == dim - 1) ? -R.par : p[j] / pt) / sqrt(1.0 + R.par * R.par);
                }
            }
            out[off + j] = v;
        }
    }
    // CTA-per-leaf for large leaves
    for (int b = gtid / (int)blockDim.x; b < T.nblk; b += gstride / (int)blockDim.x) {
        const int dim = T.bdim[b];
        if (dim <= kSmallDim || !owns_block(T, b)) continue;   // CTA-uniform
        const int off = T.boff[b], kind = T.bkind[b];
        const bool nb = kind >= kSphere;
        NormRule R;
        R.kind = kind; R.par = T.bpar[b];
        if (nb) {
            const int nsq = (kind == kSoc) ? dim - 1 : dim;
            double a[1] = {0.0};
            for (int j = threadIdx.x; j < nsq; j += blockDim.x) { const double t = xv(off + j); a[0] = fma(t, t, a[0]); }
            cta_sum<1>(a, scratch);
            R.r = sqrt(a[0]); R.last = xv(off + dim - 1); R.finish();
        }
        double a3[3] = {0.0, 0.0, 0.0};   // |x-P(x)|^2, |P(x)|^2, |P(x)[:-1]|^2
        for (int j = threadIdx.x; j < dim; j += blockDim.x) {
            const double x = xv(off + j);
            const double p = nb ? R.apply(x, j == dim - 1) : clamp_elem(kind, x, T.lo[off + j], T.hi[off + j]);
            const double d = x - p;
            a3[0] = fma(d, d, a3[0]);
            a3[1] = fma(p, p, a3[1]);
            if (j < dim - 1) a3[2] = fma(p, p, a3[2]);
        }
        cta_sum<3>(a3, scratch);
        const bool feas = is_close(sqrt(a3[0]), 0.0);
        const double pn = sqrt(a3[1]), pt = sqrt(a3[2]);
        double plast = 0.0;
        if (kind == kSoc) { const double xl = xv(off + dim - 1); plast = R.apply(xl, true); }
        for (int j = threadIdx.x; j < dim; j += blockDim.x) {
            double v = 0.0;
            if (feas && kind != kIdentity) {
                const double x = xv(off + j);
                const double p = nb ? R.apply(x, j == dim - 1) : clamp_elem(kind, x, T.lo[off + j], T.hi[off + j]);
                if (kind == kLower) v = is_close(p, T.lo[off + j]) ? -1.0 : 0.0;
                else if (kind == kUpper) v = is_close(p, T.hi[off + j]) ? 1.0 : 0.0;
                else if (kind == kBox) v = is_close(p, T.hi[off + j]) ? 1.0 : (is_close(p, T.lo[off + j]) ? -1.0 : 0.0);
                else if (kind == kSphere) v = is_close(pn, R.par) ? p / pn : 0.0;
                else if (kind == kSoc) {
                    if (pt > 0.0 && is_close(pt, R.par * plast))
                        v = ((j == dim - 1) ? -R.par : p / pt) / sqrt(1.0 + R.par * R.par);
                }
            }
            out[off + j] = v;
        }
    }
}

}  // namespace ccqp
