"""CPU oracle for the CCQP projected-gradient hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a NumPy restatement of the algorithms in the reference
(`/root/reference/src/ccqppy/solvers.py`, `solution_spaces.py`).  It is the checker that the
CUDA path is compared against; it is never the product.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may
import it.  The product (`ccqppy_b200`) never imports anything from `oracle/`.

Parity pin: `oracle/gen_golden.py` (run in the build container, where the Python reference is
importable) checks every function below against the live reference bit-for-bit and writes the
fixtures under `tests/golden/`.  `tests/test_oracle_golden.py` re-checks the oracle against those
fixtures wherever the tests run.  Beyond the reference's own five 3x3 checks
(`tests/test_module.py:28-65`) the reference pins nothing, so the goldens generated from the live
reference are the pin.

Conventions
-----------
* The feasible set is described by the same flat block table that crosses the C-ABI
  (`include/ccqp_b200.h`): `blocks[k] = (kind, offset, dim, param_off)` + one `params` array.
  One block == one leaf projection operator of the reference (a `DisjointProjOp` is the list).
* All vectors are float64 1-D; A is anything with `.dot`.
* `mv` is the reference's *reported* mat-vec count (some real products are not counted, see
  SURVEY.md section 9 Q4); `gemv` counts the products actually evaluated.
"""
from collections import deque

import numpy as np

# block kinds (shared with include/ccqp_b200.h)
IDENTITY, LOWER, UPPER, BOX, SPHERE, CONE_REF, SOC = range(7)
# solver ids (shared with include/ccqp_b200.h)
PGD, APGD, APGD_AR, BBPGD, BBPGDF, SPG, MPRGP = range(7)

EPS = np.finfo(float).eps
GD = 1e-6          # step of the projected-gradient residual, solvers.py:137


class ConeNormalNotImplemented(NotImplementedError):
    """solution_spaces.py:465 -- the reference's ConeProjOp.normal_vector raises."""


# --------------------------------------------------------------------------------------------
# projections                                                     solution_spaces.py:77-560
# --------------------------------------------------------------------------------------------
def _norm2(v):
    # np.linalg.norm of a real 1-D array is sqrt(v.dot(v))
    return np.sqrt(v.dot(v))


def _project_block(kind, par, x):
    """P(x) for one leaf operator.  `par` is the block's parameter slice."""
    d = x.shape[0]
    if kind == IDENTITY:                      # solution_spaces.py:125
        return x
    if kind == LOWER:                         # :200-201   lb*m + x*(1-m)
        lb = par[:d]
        m = x < lb
        return lb * m + x * (1 - m)
    if kind == UPPER:                         # :276-277
        ub = par[:d]
        m = x > ub
        return ub * m + x * (1 - m)
    if kind == BOX:                           # :363-366
        lb, ub = par[:d], par[d:2 * d]
        mu, ml = x > ub, x < lb
        return lb * ml + ub * mu + x * (1 - mu) * (1 - ml)
    if kind == SPHERE:                        # :431-435   (R*x)/r when r > R
        r = _norm2(x)
        return par[0] * x / r if r > par[0] else x
    if kind == CONE_REF:                      # :484-492   bug-compatible ("this op is bugged")
        mu = par[0]
        r = _norm2(x)                         # norm over the WHOLE block, last entry included
        if mu * x[-1] >= r:
            return x
        if -x[-1] / mu >= r:
            return np.zeros_like(x)
        return (x[-1] + mu * r) / (mu ** 2 + 1) * np.concatenate((x[:-1] / r, [-mu]))
    if kind == SOC:
        # Extension, not in the reference ("parity unpinned"): the correct projection onto
        # {(u,z): |u| <= mu z}.  SURVEY.md section 8a row P6.
        mu = par[0]
        u, z = x[:-1], x[-1]
        t = _norm2(u)
        if t <= mu * z:
            return x
        if mu * t <= -z:
            return np.zeros_like(x)
        s = (mu * t + z) / (mu * mu + 1.0)
        out = np.empty_like(x)
        out[:-1] = (mu * s) * u / t
        out[-1] = s
        return out
    raise ValueError("unknown block kind %r" % (kind,))


def _normal_block(kind, par, x):
    """Outward normal of one leaf operator at x (zero when x is infeasible / interior)."""
    d = x.shape[0]
    if kind == IDENTITY:                      # :98
        return np.zeros(d)
    if kind == CONE_REF:                      # :465
        raise ConeNormalNotImplemented("Cone normal not implemented, yet.")
    px = _project_block(kind, par, x)
    if not np.isclose(_norm2(x - px), 0):     # :153 :229 :313 :396
        return np.zeros(d)
    nv = np.zeros(d)
    if kind == LOWER:                         # :157-159
        nv[np.isclose(px, par[:d])] = -1
    elif kind == UPPER:                       # :233-235
        nv[np.isclose(px, par[:d])] = 1
    elif kind == BOX:                         # :317-321   upper bound is tested first
        at_ub = np.isclose(px, par[d:2 * d])
        at_lb = np.isclose(px, par[:d]) & ~at_ub
        nv[at_ub] = 1
        nv[at_lb] = -1
    elif kind == SPHERE:                      # :400-402
        r = _norm2(px)
        if np.isclose(r, par[0]):
            nv = px / r
    elif kind == SOC:
        # extension: normal of the cone surface mu*|u| = z, or zero in the interior / at apex
        mu = par[0]
        u, z = px[:-1], px[-1]
        t = _norm2(u)
        if t > 0 and np.isclose(t, mu * z):
            nv[:-1] = u / t
            nv[-1] = -mu
            nv = nv / np.sqrt(1.0 + mu * mu)
    return nv


def project(blocks, params, x):
    """P(x) for the whole table.  A single block covering all of x behaves like the bare
    operator; several blocks behave like DisjointProjOp (solution_spaces.py:553-560)."""
    x = np.asarray(x, dtype=np.float64)
    if len(blocks) == 1:
        kind, off, dim, poff = (int(v) for v in blocks[0])
        out = _project_block(kind, params[poff:], x[off:off + dim])
        return out
    out = np.zeros(x.shape[0])
    for kind, off, dim, poff in blocks:
        kind, off, dim, poff = int(kind), int(off), int(dim), int(poff)
        out[off:off + dim] = _project_block(kind, params[poff:], x[off:off + dim])
    return out


def normal_vector(blocks, params, x):
    """normal_vector(x) for the whole table (solution_spaces.py:518-525)."""
    x = np.asarray(x, dtype=np.float64)
    out = np.zeros(x.shape[0])
    for kind, off, dim, poff in blocks:
        kind, off, dim, poff = int(kind), int(off), int(dim), int(poff)
        out[off:off + dim] = _normal_block(kind, params[poff:], x[off:off + dim])
    return out


def _projected_gradient_block(kind, par, x, g):
    """(free gradient, chopped gradient) of one leaf operator, solution_spaces.py:100-109 (Identity: the method has
    no body and returns None), :162-184 (Lower), :238-260 (Upper), :324-347 (Box, its condition kept as written),
    :405-415 (Sphere: raises).  ConeProjOp has no such method at all (:467 is `proximal_gradient`)."""
    d = x.shape[0]
    if kind == IDENTITY:
        return None
    if kind == SPHERE:
        raise NotImplementedError("Cone proximal gradient not implemented, yet.")      # :415, message as written
    if kind in (CONE_REF, SOC):
        raise AttributeError("'ConeProjOp' object has no attribute 'projected_gradient'")
    nv = _normal_block(kind, par, x)
    if kind == LOWER:
        act = np.isclose(x, par[:d])
    elif kind == UPPER:
        act = np.isclose(x, par[:d])
    else:
        lb, ub = par[:d], par[d:2 * d]
        # `np.isclose(x[i], self.lower_bound[i] or x[i]<self.upper_bound[i])` (:340): Python's `or` yields the lower
        # bound unless it is zero (NaN is truthy), else the boolean x[i] < ub[i], which isclose reads as 1.0 / 0.0
        other = np.where((lb != 0) | np.isnan(lb), lb, (x < ub).astype(float))
        act = np.isclose(x, ub) | (x > ub) | np.isclose(x, other)
    t = nv * g
    chop = g - np.where(np.isnan(t) | (t < 0), t, 0.0) * nv      # g[i] - np.min((normal[i]*g[i], 0))*normal[i]
    return np.where(act, 0.0, g), np.where(act, chop, 0.0)


def projected_gradient(blocks, params, x, g):
    """projected_gradient(x, g) of the whole table: a single block behaves like the bare operator, several like
    DisjointProjOp.projected_gradient (solution_spaces.py:527-538), which unpacks every leaf's result -- so an
    Identity leaf (whose method returns None) makes it raise TypeError, as in the reference."""
    x, g = np.asarray(x, dtype=np.float64), np.asarray(g, dtype=np.float64)
    if len(blocks) == 1:
        kind, off, dim, poff = (int(v) for v in blocks[0])
        return _projected_gradient_block(kind, params[poff:], x[off:off + dim], g[off:off + dim])
    free, chop = np.zeros(x.shape[0]), np.zeros(x.shape[0])
    for kind, off, dim, poff in blocks:
        kind, off, dim, poff = int(kind), int(off), int(dim), int(poff)
        r = _projected_gradient_block(kind, params[poff:], x[off:off + dim], g[off:off + dim])
        if r is None:
            raise TypeError("cannot unpack non-iterable NoneType object")
        free[off:off + dim], chop[off:off + dim] = r
    return free, chop


# --------------------------------------------------------------------------------------------
# solvers                                                               solvers.py:71-1224
# --------------------------------------------------------------------------------------------
class _Problem:
    """Bundles (A, b, P) and counts products; keeps the solver bodies short."""

    def __init__(self, A, b, blocks, params):
        self.A = A
        self.b = np.asarray(b)
        self.n = self.b.shape[0]
        self.blocks = blocks
        self.params = params
        self.gemv = 0                          # products actually evaluated
        self.c = 1.0 / (3 * self.n * GD)       # solvers.py:138

    def mul(self, v):
        self.gemv += 1
        return self.A.dot(v)

    def P(self, v):
        return project(self.blocks, self.params, v)

    def normal(self, v):
        return normal_vector(self.blocks, self.params, v)

    def res(self, x, g):                       # solvers.py:137-139 (scale first, then norm)
        return np.linalg.norm(self.c * (x - self.P(x - GD * g)))

    def feasible_mask(self, v):                # solvers.py:1081 etc.
        return np.isclose(v, self.P(v))


def _bb_step(s, y):                            # solvers.py:655-656
    return s.dot(s) / (s.dot(y) + 10 * EPS)


def _pgd_family(pb, x0, tol, max_mv, mode, step):
    """PGD (:114-170), BBPGD (:606-669), BBPGDf (:741-819) share one skeleton."""
    x = np.copy(x0)
    xm = np.copy(x0)
    xmin, gmin, resmin = np.copy(x0), np.copy(x0), np.inf
    gm = pb.mul(xm) + pb.b
    mv = 1
    res = pb.res(xm, gm)
    if res >= tol:
        if mode != PGD:
            step = gm.dot(gm) / gm.dot(pb.mul(gm))      # product NOT counted (:635, :775)
        while True:
            x = pb.P(xm - step * gm)
            g = pb.mul(x) + pb.b
            mv += 1
            if mv >= max_mv:
                break
            res = pb.res(x, g)
            if res < tol:
                break
            if mode == BBPGDF:                            # :793-800
                if res < resmin:
                    resmin, xmin, gmin = res, np.copy(x), np.copy(g)
                if step < 10 * EPS:
                    x = pb.P(xmin - GD * gmin)           # x replaced, g is not
            if mode != PGD:
                step = _bb_step(x - xm, g - gm)
            xm, gm = x, g
    return x, res, mv


def _apgd_family(pb, x0, tol, max_mv, anti_relax):
    """APGD (:242-343) and its anti-relaxation variant (:415-533)."""
    n = pb.n
    x = np.copy(x0)
    y = np.copy(x0)
    xhat = np.ones(n)
    theta = 1.0
    d0 = x - np.ones(n)
    L = np.linalg.norm(pb.mul(d0)) / np.linalg.norm(d0)
    mv = 1
    t = 1.0 / L
    res = np.nan                                           # Q16: the reference raises NameError
    xp = np.copy(x0)
    resmin = np.inf
    while True:
        Ay = pb.mul(y)
        mv += 1
        if mv >= max_mv:
            break
        g = Ay + pb.b
        xp = pb.P(y - t * g)
        r1 = y.dot(Ay) * 0.5
        r2 = y.dot(pb.b)
        while True:
            Axp = pb.mul(xp)
            mv += 1
            if mv >= max_mv:
                break                                      # leaves the INNER loop only (:292)
            l1 = xp.dot(Axp) * 0.5
            l2 = xp.dot(pb.b)
            dxy = xp - y
            r3 = g.dot(dxy)
            r4 = 0.5 * L * dxy.dot(dxy)
            if (l1 + l2) <= (r1 + r2 + r3 + r4):
                break
            L *= 2
            t = 1.0 / L
            xp = pb.P(y - t * g)
        theta_n = 0.5 * (-theta * theta + theta * np.sqrt(4 + theta * theta))
        beta = theta * (1 - theta) / (theta * theta + theta_n)
        yn = (1 + beta) * xp - beta * x
        res = pb.res(xp, Axp + pb.b)
        if anti_relax and res < resmin:                    # :501-503
            resmin = res
            xhat = np.copy(xp)
        if res < tol:
            break
        if anti_relax and g.dot(xp - x) > 0:               # :510-512
            yn = np.copy(xp)
            theta_n = 1
        L *= 0.9
        t = 1.0 / L
        # buffer swap (:332-334): afterwards "xkp1" names the OLD xk, which is what is returned
        # when the outer loop then stops on the mat-vec limit
        y, theta = yn, theta_n
        x, xp = xp, x
    return (xhat if anti_relax else xp), res, mv


def _spg(pb, x0, tol, max_mv, m, tau, sig1, sig2, draw):
    """SPG-QP (:906-975).  `draw()` returns the next U[0,1) sample; the reference draws from
    the global NumPy RNG (:959) and np.random.uniform(lo,hi) == lo + (hi-lo)*random_sample()."""
    x = np.copy(x0)
    g = pb.mul(x) + pb.b
    f = np.dot(g, x)                                        # :923 (not the objective; kept)
    alpha = g.dot(g) / g.dot(pb.mul(g))
    mv = 2
    window = deque([f], maxlen=m)
    dd = np.nan
    draws = 0
    while True:
        d = pb.P(x - alpha * g) - x
        Ad = pb.mul(d)
        mv += 1
        if mv >= max_mv:
            break
        dd = np.dot(d, d)
        dAd = np.dot(d, Ad)
        dg = np.dot(d, g)
        if np.sqrt(dd) <= tol:
            break
        fmax = max(window)
        xi = (fmax - f) / dAd
        beta = -dg / dAd
        bhat = tau * beta + np.sqrt((tau ** 2) * (beta ** 2) + 2 * xi)
        hi = min(bhat, sig2)
        if hi != hi:
            raise OverflowError("Range exceeds valid bounds")   # what np.random.uniform does
        bk = sig1 + (hi - sig1) * draw()
        draws += 1
        x += bk * d
        g += bk * Ad
        f += bk * bk * dg + 0.5 * (bk ** 2) * dAd               # :963, as written
        window.append(f)
        alpha = dd / dAd
    return x, np.sqrt(dd), mv, draws


def _mprgp(pb, x0, tol, max_mv):
    """MPRGP with BB/expansion steps (:1048-1200)."""
    xk = pb.P(x0)
    xn = np.copy(xk)                                        # "xkp1"
    gk = pb.mul(xk) + pb.b
    gn = np.copy(gk)                                        # "gkp1"
    mv = 1
    res = pb.res(xk, gk)
    if res >= tol:
        abb = gk.dot(gk) / gk.dot(pb.mul(gk))               # counted (:1077-1078)
        mv += 1
        p = pb.feasible_mask(xk) * gk
        while True:
            Ax = pb.mul(xk)
            mv += 1
            if mv >= max_mv:
                break
            gk = Ax + pb.b
            delta = pb.feasible_mask(xk)
            psi = delta * gk
            nv = pb.normal(xk)
            bet = (1 - delta) * (gk - np.min([0, nv.dot(gk)]) * nv)
            if bet.dot(bet) < psi.dot(psi):
                Ap = pb.mul(p)
                mv += 1
                if mv >= max_mv:
                    break
                pAp = p.dot(Ap)
                acg = psi.dot(p) / pAp
                y = xk - acg * p
                af = acg + 10 * EPS                          # feasibility bisection :1112-1118
                while True:
                    yf = xk - af * p
                    if np.all(pb.feasible_mask(yf)):
                        break
                    af *= 0.5
                if acg <= af:                                # CG step :1121-1135
                    xn = np.copy(y)
                    gn = gk - acg * Ap
                    dx = xn - xk
                    abb = dx.dot(dx) / (dx.dot(pb.mul(dx)) + 10 * EPS)      # not counted
                    psi_y = pb.feasible_mask(y) * gn
                    p = psi_y - (psi_y * Ap / pAp) * p       # elementwise "beta" (:1134)
                else:                                        # expansion step :1136-1163
                    xh = xk - af * p
                    gh = gk - af * Ap
                    a = _bb_step(xh - xk, gh - gk)
                    xn = pb.P(xh - a * gh)
                    gn = pb.mul(xn) + pb.b
                    mv += 1
                    if mv >= max_mv:
                        break
                    p = pb.feasible_mask(xn) * gn
                    dx = xn - xk
                    abb = dx.dot(dx) / (dx.dot(pb.mul(dx)) + 10 * EPS)      # not counted
            else:                                            # proportioning :1164-1182
                xn = pb.P(xk - abb * gk)
                dx = xn - xk
                abb = dx.dot(dx) / (dx.dot(pb.mul(dx)) + 10 * EPS)          # not counted
                gk = pb.mul(xk) + pb.b
                mv += 1
                if mv >= max_mv:
                    break
                p = pb.feasible_mask(xn) * gn                # stale gn (:1181), kept
            res = pb.res(xn, gn)
            if res < tol:
                break
            xk, xn = xn, xk
            gk, gn = gn, gk
    return xn, res, mv


def solve(solver, A, b, x0=None, blocks=None, params=None, tol=1e-8, max_mv=np.inf,
          step_size=0.01, m=5, tau=0.5, sigma1=0.01, sigma2=0.5, uniforms=None):
    """Run one solver of the reference on (A, b, P).  Returns a dict with the reference's five
    result fields (minus time) plus `gemv` (products evaluated) and `draws` (uniforms consumed).

    `uniforms`: for SPG, a 1-D array of U[0,1) samples consumed in order; None = draw from the
    global NumPy RNG exactly like the reference does (solvers.py:959)."""
    b = np.asarray(b)
    n = b.shape[0]
    if blocks is None:
        blocks = np.array([[IDENTITY, 0, n, 0]], dtype=np.int64)
        params = np.zeros(0)
    params = np.asarray(params, dtype=np.float64)
    pb = _Problem(A, b, blocks, params)
    x0 = np.zeros(n) if x0 is None else x0
    draws = 0
    if solver in (PGD, BBPGD, BBPGDF):
        x, res, mv = _pgd_family(pb, x0, tol, max_mv, solver, step_size)
    elif solver in (APGD, APGD_AR):
        x, res, mv = _apgd_family(pb, x0, tol, max_mv, solver == APGD_AR)
    elif solver == SPG:
        if uniforms is None:
            draw = np.random.random_sample
        else:
            it = iter(np.asarray(uniforms, dtype=np.float64))
            draw = lambda: next(it)
        x, res, mv, draws = _spg(pb, x0, tol, max_mv, m, tau, sigma1, sigma2, draw)
    elif solver == MPRGP:
        x, res, mv = _mprgp(pb, x0, tol, max_mv)
    else:
        raise ValueError("unknown solver id %r" % (solver,))
    return dict(solution=np.array(x, dtype=np.float64, copy=True), residual=float(res),
                converged=bool(mv < max_mv), mv=int(mv), gemv=int(pb.gemv), draws=int(draws))
