"""The oracle against the LIVE reference, bit for bit, on freshly drawn inputs (hypothesis).

Runs only where the reference is mounted (/root/reference, i.e. the build container; the GPU box does not
have it -- there the committed fixtures of tests/golden/ carry the pin).  matplotlib is not installed and
solution_spaces.py:6 imports it, so an empty stub package goes on sys.path first."""
import contextlib
import io
import os
import sys
import tempfile
import warnings

import numpy as np
import pytest

import problems as pr
from oracle import ccqp_oracle as orc

REF_SRC = "/root/reference/src"
if not os.path.isdir(os.path.join(REF_SRC, "ccqppy")):
    pytest.skip("the reference is not mounted here", allow_module_level=True)

hyp = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st     # noqa: E402

_stub = tempfile.mkdtemp(prefix="mplstub")
os.makedirs(os.path.join(_stub, "matplotlib"))
for _f in ("__init__.py", "pyplot.py"):
    open(os.path.join(_stub, "matplotlib", _f), "w").close()
# The repository's own drop-in package is ALSO called `ccqppy`: make sure the name resolves to the reference while it is
# imported here, check that it did, and take the name out of sys.modules / sys.path again so that later tests get the shim.
for _k in [m for m in sys.modules if m == "ccqppy" or m.startswith("ccqppy.")]:
    del sys.modules[_k]
sys.path.insert(0, _stub)
sys.path.insert(0, REF_SRC)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    import ccqppy.solvers as ref_solvers                   # noqa: E402  (the reference)
    import ccqppy.solution_spaces as ref_ss                # noqa: E402
assert os.path.realpath(ref_solvers.__file__).startswith(os.path.realpath(REF_SRC)), ref_solvers.__file__
for _k in [m for m in sys.modules if m == "ccqppy" or m.startswith("ccqppy.")]:
    del sys.modules[_k]
sys.path.remove(REF_SRC)


def ref_op(tab):
    ops, par = [], tab.params
    for kind, off, dim, poff in tab.blocks:
        kind, dim, poff = int(kind), int(dim), int(poff)
        ops.append({pr.IDENTITY: lambda: ref_ss.IdentityProjOp(dim),
                    pr.LOWER: lambda: ref_ss.LowerBoundProjOp(dim, par[poff:poff + dim]),
                    pr.UPPER: lambda: ref_ss.UpperBoundProjOp(dim, par[poff:poff + dim]),
                    pr.BOX: lambda: ref_ss.BoxProjOp(dim, par[poff:poff + dim], par[poff + dim:poff + 2 * dim]),
                    pr.SPHERE: lambda: ref_ss.SphereProjOp(dim, par[poff]),
                    pr.CONE_REF: lambda: ref_ss.ConeProjOp(dim, par[poff])}[kind]())
    return ops[0] if len(ops) == 1 else ref_ss.DisjointProjOp(*ops)


def ref_solve(solver, A, b, tab, tol, max_mv, step, seed, x0=None):
    S = ref_solvers
    s = {pr.PGD: lambda: S.CCQPSolverPGD(tol, max_mv, step), pr.APGD: lambda: S.CCQPSolverAPGD(tol, max_mv),
         pr.APGD_AR: lambda: S.CCQPSolverAPGDAntiRelaxation(tol, max_mv), pr.BBPGD: lambda: S.CCQPSolverBBPGD(tol, max_mv),
         pr.BBPGDF: lambda: S.CCQPSolverBBPGDf(tol, max_mv), pr.SPG: lambda: S.CCQPSolverSPG(tol, max_mv),
         pr.MPRGP: lambda: S.CCQPSolverMPRGP(tol, max_mv)}[solver]()
    np.random.seed(seed)
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        s.solve(A, b, x0=x0, convex_proj_op=ref_op(tab))
    return s


def random_table(rng, n):
    """A random disjoint union of every operator kind the reference implements, dimensions summing to n."""
    t, left = pr.Table(), n
    while left > 0:
        kind = int(rng.integers(0, 6))
        dim = int(min(left, rng.integers(1, 6)))
        if kind == pr.IDENTITY: t.add(kind, dim)
        elif kind == pr.LOWER: t.add(kind, dim, -0.5 - rng.random(dim))
        elif kind == pr.UPPER: t.add(kind, dim, 0.5 + rng.random(dim))
        elif kind == pr.BOX:
            lo = -1.0 - rng.random(dim)
            t.add(kind, dim, lo, lo + 1.0 + 2.0 * rng.random(dim))
        elif kind == pr.SPHERE: t.add(kind, dim, 0.3 + rng.random())
        else: t.add(pr.CONE_REF, dim, 0.3 + rng.random())
        left -= dim
    return t


@settings(max_examples=60, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(1, 40), scale=st.sampled_from([0.1, 1.0, 5.0]))
def test_projections_and_normals_bit_for_bit(seed, n, scale):
    rng = np.random.default_rng(seed)
    tab = random_table(rng, n)
    op = ref_op(tab)
    has_cone = any(int(k) == pr.CONE_REF for k in tab.blocks[:, 0])
    for x in (scale * rng.standard_normal(n), np.zeros(n), np.asarray(op(scale * rng.standard_normal(n)), dtype=float)):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = np.asarray(op(x.copy()), dtype=float)
        got = orc.project(tab.blocks, tab.params, x.copy())
        assert np.array_equal(got, want)
        if has_cone:
            with pytest.raises(NotImplementedError):
                orc.normal_vector(tab.blocks, tab.params, x.copy())
        else:
            assert np.array_equal(orc.normal_vector(tab.blocks, tab.params, x.copy()), np.asarray(op.normal_vector(x.copy()), dtype=float))


@settings(max_examples=60, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(1, 30), kind=st.sampled_from(["lower", "upper", "box", "box0", "disjoint"]),
       on_boundary=st.booleans())
def test_projected_gradient_bit_for_bit(seed, n, kind, on_boundary):
    """Row f-4: solution_spaces.py:162-184, 238-260, 324-347 (incl. the `lower_bound[i] or ...` test of :340), 527-538."""
    rng = np.random.default_rng(seed)
    lo = np.where(rng.random(n) < 0.3, 0.0, -rng.random(n) - 0.2) if kind == "box0" else -rng.random(n) - 0.2
    hi = lo + 0.5 + rng.random(n)
    tab = pr.Table()
    if kind == "lower":
        tab.add(pr.LOWER, n, lo)
    elif kind == "upper":
        tab.add(pr.UPPER, n, hi)
    elif kind in ("box", "box0"):
        tab.add(pr.BOX, n, lo, hi)
    else:
        tab.add(pr.BOX, n, lo, hi).add(pr.LOWER, n, lo).add(pr.UPPER, n, hi)
    op = ref_op(tab)
    x = 1.5 * rng.standard_normal(tab.n)
    if on_boundary:
        x = np.asarray(op(x), dtype=float)
        x[rng.random(tab.n) < 0.2] = [0.0, 1.0][seed % 2]
    g = rng.standard_normal(tab.n)
    f, c = op.projected_gradient(x, g)
    of, oc = orc.projected_gradient(tab.blocks, tab.params, x, g)
    assert np.array_equal(np.asarray(f, float), of) and np.array_equal(np.asarray(c, float), oc)


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(2, 48), solver=st.sampled_from(list(range(7))),
       mu=st.sampled_from([1.0, 0.1]), warm=st.booleans())
def test_solvers_bit_for_bit(seed, n, solver, mu, warm):
    rng = np.random.default_rng(seed)
    A, b = pr.shift_problem(n, seed % 1000, mu)
    tab = random_table(rng, n)
    while any(int(k) == pr.CONE_REF for k in tab.blocks[:, 0]):      # the reference's cone does not converge / raises
        tab = random_table(rng, n)
    x0 = rng.standard_normal(n) if warm else None
    step = 1.0 / np.abs(A).sum(axis=1).max()
    r = ref_solve(solver, A, b, tab, 1e-7, 400, step, seed % 97, x0)
    after_ref = np.random.random_sample()
    np.random.seed(seed % 97)
    o = orc.solve(solver, A, b, x0=x0, blocks=tab.blocks, params=tab.params, tol=1e-7, max_mv=400, step_size=step)
    after_orc = np.random.random_sample()
    assert o["mv"] == r.solution_num_matrix_vector_multiplications and o["converged"] == r.solution_converged
    assert np.array_equal(o["solution"], np.asarray(r.solution, dtype=float))
    assert o["residual"] == r.solution_residual or (np.isnan(o["residual"]) and np.isnan(r.solution_residual))
    assert after_ref == after_orc                                   # SPG leaves the global RNG in the same state


def test_sparse_operator_form_bit_for_bit():
    """A scipy.sparse Hessian goes through A.dot in the reference and in the oracle alike (row f-3)."""
    sp = pytest.importorskip("scipy.sparse")
    n = 60
    rng = np.random.default_rng(5)
    D = sp.random(n, n, density=0.1, random_state=np.random.RandomState(5), format="csr", data_rvs=lambda k: rng.standard_normal(k))
    A = (D.T @ D + 0.5 * sp.identity(n)).tocsr()
    b = -(A @ (1 - 4 * rng.random(n)))
    tab = pr.mixed_table(n)
    for solver in (pr.APGD, pr.BBPGD, pr.SPG, pr.MPRGP):
        r = ref_solve(solver, A, b, tab, 1e-7, 2000, 0.01, 3)
        np.random.seed(3)
        o = orc.solve(solver, A, b, blocks=tab.blocks, params=tab.params, tol=1e-7, max_mv=2000)
        assert o["mv"] == r.solution_num_matrix_vector_multiplications
        assert np.array_equal(o["solution"], np.asarray(r.solution, dtype=float))
