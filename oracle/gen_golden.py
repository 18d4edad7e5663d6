"""Generate tests/golden/* from the LIVE reference and pin the oracle against it.

TEST INFRASTRUCTURE.  Runs only in the build container, where the Python reference is mounted at
/root/reference (it does not exist on the GPU box).  matplotlib is not installed and
solution_spaces.py:6 imports it, so an empty stub package is put on sys.path first.

    python oracle/gen_golden.py            # writes tests/golden/solvers.json, solvers.npz,
                                           #        tests/golden/projections.npz

For every case the reference solver is run (np.random.seed(s) first for SPG), the oracle
restatement is run on the same inputs, and the two are required to agree BIT FOR BIT (solution
array, residual, mat-vec count, converged flag) before the case is written.  The fixtures
therefore are outputs of the reference itself.
"""
import contextlib
import io
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

_stub = tempfile.mkdtemp(prefix="mplstub")
os.makedirs(os.path.join(_stub, "matplotlib"))
for f in ("__init__.py", "pyplot.py"):
    open(os.path.join(_stub, "matplotlib", f), "w").close()
sys.path.insert(0, _stub)
sys.path.insert(0, "/root/reference/src")

import warnings
warnings.simplefilter("ignore")

import ccqppy.solvers as ref_solvers              # noqa: E402  (the reference)
import ccqppy.solution_spaces as ref_ss           # noqa: E402
import ccqppy.problem_suite as ref_suite          # noqa: E402
assert os.path.realpath(ref_solvers.__file__).startswith("/root/reference/"), "the name ccqppy must be the reference here"

import problems as pr                              # noqa: E402  (tests/problems.py)
from oracle import ccqp_oracle as orc              # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def ref_op_from_table(tab):
    ops = []
    par = tab.params
    for kind, off, dim, poff in tab.blocks:
        kind, dim, poff = int(kind), int(dim), int(poff)
        if kind == pr.IDENTITY:
            ops.append(ref_ss.IdentityProjOp(dim))
        elif kind == pr.LOWER:
            ops.append(ref_ss.LowerBoundProjOp(dim, par[poff:poff + dim]))
        elif kind == pr.UPPER:
            ops.append(ref_ss.UpperBoundProjOp(dim, par[poff:poff + dim]))
        elif kind == pr.BOX:
            ops.append(ref_ss.BoxProjOp(dim, par[poff:poff + dim], par[poff + dim:poff + 2 * dim]))
        elif kind == pr.SPHERE:
            ops.append(ref_ss.SphereProjOp(dim, par[poff]))
        elif kind == pr.CONE_REF:
            ops.append(ref_ss.ConeProjOp(dim, par[poff]))
        else:
            raise ValueError("kind %d has no reference operator" % kind)
    return ops[0] if len(ops) == 1 else ref_ss.DisjointProjOp(*ops)


def ref_solver(solver, tol, max_mv, step):
    S = ref_solvers
    return {pr.PGD: lambda: S.CCQPSolverPGD(tol, max_mv, step),
            pr.APGD: lambda: S.CCQPSolverAPGD(tol, max_mv),
            pr.APGD_AR: lambda: S.CCQPSolverAPGDAntiRelaxation(tol, max_mv),
            pr.BBPGD: lambda: S.CCQPSolverBBPGD(tol, max_mv),
            pr.BBPGDF: lambda: S.CCQPSolverBBPGDf(tol, max_mv),
            pr.SPG: lambda: S.CCQPSolverSPG(tol, max_mv),
            pr.MPRGP: lambda: S.CCQPSolverMPRGP(tol, max_mv)}[solver]()


# ---- case catalogue -----------------------------------------------------------------------
def make_problem(spec):
    kind = spec["gen"]
    if kind == "tridiag":
        return pr.tridiag_problem()
    if kind == "shift":
        return pr.shift_problem(spec["n"], spec["seed"], spec["mu"])
    if kind == "wishart":
        return pr.wishart_problem(spec["n"], spec["seed"])
    raise ValueError(kind)


def make_table(spec):
    name, args = spec["table"], spec.get("table_args", {})
    if name == "suite":
        t = pr.Table()
        which = args["which"]
        if which == "identity":
            t.add(pr.IDENTITY, 3)
        elif which == "identity3":
            for _ in range(3):
                t.add(pr.IDENTITY, 1)
        else:
            lo, hi = args["lo"], args["hi"]
            t.add(pr.BOX, 3, np.array(lo, dtype=float), np.array(hi, dtype=float))
        return t
    return getattr(pr, name + "_table")(**args)


def catalogue():
    cases = []
    suite = [("UnconstrainedSPD1", dict(which="identity")),
             ("UnconstrainedSPD2", dict(which="identity3")),
             ("BoxConstrainedSPD", dict(which="box", lo=[0, 0, 0], hi=[2, 2, 2])),
             ("ThinBoxConstrainedSPD", dict(which="box", lo=[-10, -0.1, 0.9], hi=[10, 0.1, 1.1])),
             ("ActiveBoxConstrainedSPD", None)]
    # problem_suite x 7 solvers, the parameters of tests/test_module.py:28-65
    for pname, targs in suite:
        if targs is None:
            prob = ref_suite.ActiveBoxConstrainedSPD()
            op = prob.convex_proj_op
            targs = dict(which="box", lo=np.asarray(op.lower_bound, float).tolist(),
                         hi=np.asarray(op.upper_bound, float).tolist())
            bvec = np.asarray(prob.b, float).tolist()
        else:
            bvec = None
        for s in range(7):
            cases.append(dict(name="suite/%s/%s" % (pname, pr.SOLVER_NAMES[s]), gen="tridiag",
                              b_override=bvec, table="suite", table_args=targs, solver=s,
                              tol=1e-8, max_mv=10000, step=0.1, spg_seed=0))
    # README config, three seeds (README.md:35-43)
    for seed in (0, 1, 2):
        cases.append(dict(name="readme/SPG/seed%d" % seed, gen="tridiag", table="suite",
                          table_args=dict(which="box", lo=[-2, -2, -4], hi=[2, 2, 5]),
                          solver=pr.SPG, tol=1e-10, max_mv=5000, step=0.01, spg_seed=seed))
    # batched-style n=64 box QPs (config 4)
    for seed in range(6):
        for s in (pr.PGD, pr.APGD, pr.BBPGD, pr.BBPGDF, pr.SPG, pr.MPRGP, pr.APGD_AR):
            cases.append(dict(name="n64/seed%d/%s" % (seed, pr.SOLVER_NAMES[s]), gen="shift", n=64,
                              seed=seed, mu=1.0, table="box", table_args=dict(n=64), solver=s,
                              tol=1e-8, max_mv=5000, step=0.1, spg_seed=seed))
    # every operator kind, n=300
    for mu in (1.0, 0.01):
        for tname, targs in (("mixed", dict(n=300)), ("sphere3", dict(n=300)),
                             ("lower", dict(n=300)), ("upper", dict(n=300)),
                             ("sphere", dict(n=300, radius=3.0)), ("identity", dict(n=300)),
                             ("box", dict(n=300))):
            for s in range(7):
                cases.append(dict(name="n300/mu%g/%s/%s" % (mu, tname, pr.SOLVER_NAMES[s]),
                                  gen="shift", n=300, seed=1, mu=mu, table=tname, table_args=targs,
                                  solver=s, tol=1e-5, max_mv=5000, step=0.05, spg_seed=1))
    # dense mid size, the gate generator, box
    for mu in (1.0, 0.01):
        for s in (pr.PGD, pr.APGD, pr.BBPGD, pr.SPG, pr.MPRGP):
            cases.append(dict(name="n1024/mu%g/box/%s" % (mu, pr.SOLVER_NAMES[s]), gen="shift",
                              n=1024, seed=0, mu=mu, table="box", table_args=dict(n=1024), solver=s,
                              tol=1e-5, max_mv=5000, step=0.2, spg_seed=0))
    # non-zero x0 outside the feasible set, and mv-limit behaviour (APGD ends at max+1)
    for s in range(7):
        cases.append(dict(name="n300/x0/%s" % pr.SOLVER_NAMES[s], gen="shift", n=300, seed=3,
                          mu=1.0, table="mixed", table_args=dict(n=300), solver=s, tol=1e-6,
                          max_mv=5000, step=0.05, spg_seed=3, x0_seed=11))
        cases.append(dict(name="n300/maxmv/%s" % pr.SOLVER_NAMES[s], gen="shift", n=300, seed=4,
                          mu=0.01, table="box", table_args=dict(n=300), solver=s, tol=1e-12,
                          max_mv=12, step=0.05, spg_seed=4))
    # benchmark-faithful Wishart at the benchmark tolerance
    for s in (pr.APGD, pr.BBPGD, pr.SPG):
        cases.append(dict(name="wishart256/%s" % pr.SOLVER_NAMES[s], gen="wishart", n=256, seed=0,
                          table="box", table_args=dict(n=256), solver=s, tol=1e-5, max_mv=5000,
                          step=0.01, spg_seed=0))
    return cases


def case_inputs(c):
    A, b = make_problem(c)
    if c.get("b_override") is not None:
        b = np.asarray(c["b_override"], dtype=float)
    tab = make_table(c)
    x0 = None
    if c.get("x0_seed") is not None:
        x0 = 3.0 * np.random.default_rng(c["x0_seed"]).standard_normal(b.shape[0])
    return A, b, tab, x0


def run_reference(c, A, b, tab, x0):
    op = ref_op_from_table(tab)
    sol = ref_solver(c["solver"], c["tol"], c["max_mv"], c["step"])
    np.random.seed(c["spg_seed"])
    with contextlib.redirect_stdout(io.StringIO()):
        r = sol.solve(A, b, x0=x0, convex_proj_op=op)
    return dict(solution=np.array(r.solution, dtype=float), residual=float(r.solution_residual),
                converged=bool(r.solution_converged),
                mv=int(r.solution_num_matrix_vector_multiplications))


def run_oracle(c, A, b, tab, x0, explicit_uniforms):
    np.random.seed(c["spg_seed"])
    uni = pr.spg_uniforms(c["spg_seed"], 20000) if explicit_uniforms else None
    return orc.solve(c["solver"], A, b, x0=x0, blocks=tab.blocks, params=tab.params,
                     tol=c["tol"], max_mv=c["max_mv"], step_size=c["step"], uniforms=uni)


class _Perturbed:
    """A.dot with a relative perturbation of about one ulp per entry: a stand-in for "the same
    product summed in another order" (what any other BLAS, or the GPU kernel, does)."""

    def __init__(self, A, rng):
        self.A = np.asarray(A, dtype=float)
        self.absA = np.abs(self.A)
        self.rng = rng

    def dot(self, v):
        y = self.A.dot(v)
        return y + 2.2e-16 * (self.rng.random(y.shape) - 0.5) * self.absA.dot(np.abs(v))


def stability_band(c, A, b, tab, x0, trials=6):
    """Mat-vec counts of the oracle under rounding-level perturbations of the product.  A case
    whose count never moves is "order stable" and is gated at the north_star tolerance; the
    others (ill-conditioned Hessians, sign tests on rounding-level quantities) are chaotic in
    the reference itself and are only checked against this band."""
    rng = np.random.default_rng(12345)
    counts = []
    for _ in range(trials):
        o = orc.solve(c["solver"], _Perturbed(A, rng), b, x0=x0, blocks=tab.blocks, params=tab.params,
                      tol=c["tol"], max_mv=c["max_mv"], step_size=c["step"],
                      uniforms=pr.spg_uniforms(c["spg_seed"], 20000))
        counts.append(o["mv"])
    return min(counts), max(counts)


def gen_solvers():
    meta, arrays = [], {}
    for c in catalogue():
        A, b, tab, x0 = case_inputs(c)
        ref = run_reference(c, A, b, tab, x0)
        for explicit in (False, True):
            o = run_oracle(c, A, b, tab, x0, explicit)
            same = (np.array_equal(o["solution"], ref["solution"]) and o["mv"] == ref["mv"]
                    and o["converged"] == ref["converged"]
                    and (o["residual"] == ref["residual"]
                         or (np.isnan(o["residual"]) and np.isnan(ref["residual"]))))
            if not same:
                raise SystemExit("ORACLE != REFERENCE for %s (explicit uniforms=%s): ref mv %d res %r, "
                                 "oracle mv %d res %r, max|dx| %g" %
                                 (c["name"], explicit, ref["mv"], ref["residual"], o["mv"],
                                  o["residual"], np.max(np.abs(o["solution"] - ref["solution"]))))
        entry = {k: v for k, v in c.items()}
        lo, hi = stability_band(c, A, b, tab, x0)
        entry.update(mv=ref["mv"], converged=ref["converged"], residual=ref["residual"].hex(),
                     gemv=o["gemv"], draws=o["draws"], mv_band=[min(lo, ref["mv"]), max(hi, ref["mv"])],
                     order_stable=bool(lo == hi == ref["mv"]))
        meta.append(entry)
        arrays[c["name"]] = ref["solution"]
        print("%-40s mv=%5d conv=%d res=%.3e gemv=%d band=%s" %
              (c["name"], ref["mv"], ref["converged"], ref["residual"], o["gemv"], entry["mv_band"]))
    with open(os.path.join(GOLD, "solvers.json"), "w") as f:
        json.dump(meta, f, indent=0)
    np.savez_compressed(os.path.join(GOLD, "solvers.npz"), **arrays)


def gen_projections():
    """P(x) and normal_vector(x) of the reference on seeded points that straddle the boundaries,
    for every operator kind (the reference itself only tests Identity, tests/test_module.py:11)."""
    rng = np.random.default_rng(2024)
    out = {}
    tabs = {"identity": pr.identity_table(17), "box": pr.box_table(64), "lower": pr.lower_table(33),
            "upper": pr.upper_table(33), "sphere3": pr.sphere3_table(64), "sphere": pr.sphere_table(50, 2.0),
            "mixed": pr.mixed_table(300), "cone3": pr.Table().add(pr.CONE_REF, 3, 1.0),
            "cone7": pr.Table().add(pr.CONE_REF, 7, 0.6),
            "cones": pr.Table().add(pr.CONE_REF, 3, 0.5).add(pr.CONE_REF, 3, 2.0).add(pr.CONE_REF, 4, 1.0)}
    for name, tab in tabs.items():
        op = ref_op_from_table(tab)
        n = tab.n
        X, PX, NV, NVP = [], [], [], []
        for trial in range(24):
            scale = [0.3, 1.0, 1.0, 3.0][trial % 4]
            x = scale * rng.standard_normal(n)
            if trial % 6 == 5:                        # exactly-on-boundary / near-boundary points
                x = np.asarray(op(x), dtype=float) * (1.0 + 1e-7 * (trial % 5 - 2))
            px = np.array(op(x), dtype=float)
            o = orc.project(tab.blocks, tab.params, x)
            if not np.array_equal(px, np.asarray(o, dtype=float)):
                raise SystemExit("oracle projection != reference for %s" % name)
            X.append(x)
            PX.append(px)
            if not name.startswith("cone"):
                for pt, acc in ((x, NV), (px, NVP)):
                    nv = np.array(op.normal_vector(pt), dtype=float)
                    on = orc.normal_vector(tab.blocks, tab.params, pt)
                    if not np.array_equal(nv, on):
                        raise SystemExit("oracle normal != reference for %s" % name)
                    acc.append(nv)
        out[name + "/x"] = np.array(X)
        out[name + "/px"] = np.array(PX)
        if NV:
            out[name + "/nv"] = np.array(NV)
            out[name + "/nvp"] = np.array(NVP)
        print("projection %-10s ok (%d points, n=%d)" % (name, len(X), n))
    np.savez_compressed(os.path.join(GOLD, "projections.npz"), **out)


def gen_projected_gradients():
    """projected_gradient(x, g) of the reference (solution_spaces.py:162-184, 238-260, 324-347, 527-538) on seeded
    points incl. points on the bounds, for the operator kinds that implement it, plus the exceptions of the others."""
    rng = np.random.default_rng(77)
    out, errors = {}, {}
    lb0 = np.where(rng.random(40) < 0.4, 0.0, -1.0 - rng.random(40))          # zero lower bounds hit the `or` branch of :340
    tabs = {"box": pr.box_table(64), "box_zero_lb": pr.Table().add(pr.BOX, 40, lb0, lb0 + 1.0 + rng.random(40)),
            "lower": pr.lower_table(33), "upper": pr.upper_table(33),
            "disjoint": pr.Table().add(pr.BOX, 20, -1.0, 1.0).add(pr.LOWER, 10, -0.5).add(pr.UPPER, 10, 0.25).add(pr.BOX, 7, 0.0, 2.0)}
    for name, tab in tabs.items():
        op = ref_op_from_table(tab)
        n = tab.n
        X, G, F, C = [], [], [], []
        for trial in range(24):
            x = [0.3, 1.0, 1.0, 3.0][trial % 4] * rng.standard_normal(n)
            if trial % 3 != 0:                         # feasible points with entries exactly on / next to the bounds
                x = np.asarray(op(x), dtype=float)
                if trial % 3 == 2:
                    x = x * (1.0 + 1e-7 * (trial % 5 - 2))
            if trial % 8 == 1:
                x[::5] = [0.0, 1.0][trial % 2]         # the values the quirk of :340 compares with
            g = rng.standard_normal(n)
            f, c = op.projected_gradient(x, g)
            of, oc = orc.projected_gradient(tab.blocks, tab.params, x, g)
            if not (np.array_equal(np.asarray(f, float), of) and np.array_equal(np.asarray(c, float), oc)):
                raise SystemExit("oracle projected_gradient != reference for %s" % name)
            X.append(x); G.append(g); F.append(np.asarray(f, float)); C.append(np.asarray(c, float))
        out[name + "/x"], out[name + "/g"], out[name + "/free"], out[name + "/chopped"] = map(np.array, (X, G, F, C))
        print("projected_gradient %-12s ok (%d points, n=%d)" % (name, len(X), n))
    for name, tab in {"identity": pr.identity_table(5), "sphere": pr.sphere_table(5), "cone": pr.cone_ref_table(5),
                      "disjoint_with_identity": pr.Table().add(pr.BOX, 3, -1.0, 1.0).add(pr.IDENTITY, 2)}.items():
        op = ref_op_from_table(tab)
        x, g = rng.standard_normal(5), rng.standard_normal(5)
        def outcome(fn):
            try:
                r = fn()
                return "None" if r is None else "value"
            except Exception as e:      # noqa: BLE001
                return type(e).__name__
        ref = outcome(lambda: op.projected_gradient(x, g))
        mine = outcome(lambda: orc.projected_gradient(tab.blocks, tab.params, x, g))
        if ref != mine:
            raise SystemExit("oracle projected_gradient behaviour != reference for %s: %s vs %s" % (name, mine, ref))
        errors[name] = ref
    np.savez_compressed(os.path.join(GOLD, "projected_gradients.npz"), **out)
    with open(os.path.join(GOLD, "projected_gradients.json"), "w") as f:
        json.dump(errors, f, indent=0)
    print("projected_gradient behaviours:", errors)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    if "--projected-gradients-only" in sys.argv:
        gen_projected_gradients()
        sys.exit(0)
    gen_projections()
    gen_projected_gradients()
    gen_solvers()
    print("numpy", np.__version__)
