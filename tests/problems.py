"""Seeded synthetic problems and projection tables shared by the tests, the golden generator and
bench.py.  Pure NumPy, no dependency on the reference or on the oracle.

Generators follow SURVEY.md section 8(d):
  * "shift":   A = G G^T / n + mu I   (well conditioned; the parity-gate generator)
  * "wishart": A = G G^T              (benchmark_random_ccqp.py:59-60; ill conditioned)
with x* placed so that about half of the [-1,1] box constraints are active and b = -A x*.
"""
import numpy as np

# block kinds / solver ids: must match include/ccqp_b200.h
IDENTITY, LOWER, UPPER, BOX, SPHERE, CONE_REF, SOC = range(7)
PGD, APGD, APGD_AR, BBPGD, BBPGDF, SPG, MPRGP = range(7)
SOLVER_NAMES = {PGD: "PGD", APGD: "APGD", APGD_AR: "APGD_AR", BBPGD: "BBPGD",
                BBPGDF: "BBPGDF", SPG: "SPG", MPRGP: "MPRGP"}


def shift_problem(n, seed, mu=1.0):
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((n, n))
    A = G @ G.T / n + mu * np.eye(n)
    xs = 1.0 - 4.0 * rng.random(n)
    b = -A @ xs
    return A, b


def wishart_problem(n, seed):
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((n, n))
    A = G @ G.T
    xs = 1.0 - 2.0 * rng.random(n)
    b = -A @ xs
    return A, b


def tridiag_problem():
    """The 3x3 problem of README.md:35-37 / problem_suite.py:54 (integer inputs on purpose)."""
    A = np.array([[2, -1, 0], [-1, 2, -1], [0, -1, 2]])
    b = -A.dot(np.array([1, 0, 1]))
    return A, b


class Table:
    """Flat projection table: blocks[k] = (kind, offset, dim, param_off) and a params array."""

    def __init__(self):
        self.rows = []
        self.par = []
        self.n = 0

    def add(self, kind, dim, *par):
        poff = len(self.par)
        for p in par:
            self.par.extend(np.broadcast_to(np.asarray(p, dtype=np.float64), (dim,)).tolist()
                            if kind in (LOWER, UPPER, BOX) else [float(p)])
        self.rows.append((kind, self.n, dim, poff))
        self.n += dim
        return self

    @property
    def blocks(self):
        return np.array(self.rows, dtype=np.int64).reshape(-1, 4)

    @property
    def params(self):
        return np.array(self.par, dtype=np.float64)


def identity_table(n):
    return Table().add(IDENTITY, n)


def box_table(n, lo=-1.0, hi=1.0):
    return Table().add(BOX, n, lo, hi)


def lower_table(n, lo=-1.0):
    return Table().add(LOWER, n, lo)


def upper_table(n, hi=1.0):
    return Table().add(UPPER, n, hi)


def sphere_table(n, radius=1.0):
    return Table().add(SPHERE, n, radius)


def sphere3_table(n, radius=1.0):
    """Contact-style friction discs: n//3 Sphere(3) blocks + Identity for the remainder."""
    t = Table()
    for _ in range(n // 3):
        t.add(SPHERE, 3, radius)
    if n % 3:
        t.add(IDENTITY, n % 3)
    return t


def mixed_table(n=300, seed=7):
    """Disjoint(Box, Lower, Upper, k x Sphere(3), Identity): every working operator kind
    (SURVEY.md appendix B).  Per-element bounds are randomised so the tables are not uniform."""
    assert n >= 60
    rng = np.random.default_rng(seed)
    nb = n // 3
    nl = n // 6
    nu = n // 6
    ns = (n - nb - nl - nu - n // 30) // 3
    t = Table()
    lo = -1.0 - 0.5 * rng.random(nb)
    t.add(BOX, nb, lo, lo + 1.5 + rng.random(nb))
    t.add(LOWER, nl, -0.5 - rng.random(nl))
    t.add(UPPER, nu, 0.5 + rng.random(nu))
    for _ in range(ns):
        t.add(SPHERE, 3, 0.5 + rng.random())
    t.add(IDENTITY, n - t.n)
    assert t.n == n
    return t


def soc3_table(n, mu=0.5):
    t = Table()
    for _ in range(n // 3):
        t.add(SOC, 3, mu)
    if n % 3:
        t.add(IDENTITY, n % 3)
    return t


def cone_ref_table(n, mu=1.0):
    return Table().add(CONE_REF, n, mu)


def spg_uniforms(seed, count):
    """The stream np.random.uniform would consume after np.random.seed(seed)."""
    return np.random.RandomState(seed).random_sample(count)
