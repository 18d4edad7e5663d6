"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the public API and
therefore through the C-ABI, against (a) the golden fixtures generated from the live reference
and (b) the oracle on the same seeded inputs.

Tolerances (north_star): solution within 1e-9 relative, same converged flag, mat-vec count within
2 % (reduction order differs).  Elementwise projections are bit-exact."""
import json
import os

import numpy as np
import pytest

import problems as pr
from helpers import op_from_table, run_gpu, make_solver
from oracle import ccqp_oracle as orc
from test_oracle_golden import META, PROJ, TABLES, case_inputs, check_against_golden

pytestmark = pytest.mark.gpu


# ---- pieces of the path ------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(TABLES))
def test_projection_matches_reference_golden(name):
    tab = TABLES[name]()
    op = op_from_table(tab)
    elementwise = name in ("identity", "box", "lower", "upper")
    for x, px in zip(PROJ[name + "/x"], PROJ[name + "/px"]):
        got = np.asarray(op(x.copy()))
        if elementwise:
            assert np.array_equal(got, px)            # bit-exact (sign of zero aside: == treats them equal)
        else:
            np.testing.assert_allclose(got, px, rtol=1e-15, atol=1e-300)


@pytest.mark.parametrize("name", sorted(n for n in TABLES if not n.startswith("cone")))
def test_normal_vector_matches_reference_golden(name):
    tab = TABLES[name]()
    op = op_from_table(tab)
    for x, px, nv, nvp in zip(PROJ[name + "/x"], PROJ[name + "/px"], PROJ[name + "/nv"], PROJ[name + "/nvp"]):
        np.testing.assert_allclose(np.asarray(op.normal_vector(x.copy())), nv, rtol=1e-15, atol=0)
        np.testing.assert_allclose(np.asarray(op.normal_vector(px.copy())), nvp, rtol=1e-15, atol=0)


def test_cone_normal_raises_like_reference():
    from ccqppy_b200 import solution_spaces as ss
    with pytest.raises(NotImplementedError):
        ss.ConeProjOp(3).normal_vector(np.ones(3))


def test_soc_projection_matches_oracle_extension():
    rng = np.random.default_rng(5)
    for tab in (pr.soc3_table(30, 0.5), pr.Table().add(pr.SOC, 40, 0.7)):
        op = op_from_table(tab)
        for _ in range(10):
            x = 2 * rng.standard_normal(tab.n)
            np.testing.assert_allclose(np.asarray(op(x)), orc.project(tab.blocks, tab.params, x), rtol=2e-15, atol=1e-300)
            np.testing.assert_allclose(np.asarray(op.normal_vector(np.asarray(op(x)))),
                                       orc.normal_vector(tab.blocks, tab.params, orc.project(tab.blocks, tab.params, x)),
                                       rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("n", [1, 3, 64, 127, 300, 1000, 1023, 2048, 4100, 8200, 16500])
def test_gemv_matches_numpy(n):
    import ctypes
    from ccqppy_b200 import _capi
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n))
    v = rng.standard_normal(n)
    h = _capi.Handle()
    pa, ma, _ = _capi.f64_ptr(A)
    _capi.check(h.h, h.lib.ccqp_set_matrix(h.h, pa, n, n, 0, n, ma))
    y = np.empty(n)
    pv, mv, _ = _capi.f64_ptr(v)
    py, _, _ = _capi.f64_ptr(y)
    _capi.check(h.h, h.lib.ccqp_gemv(h.h, pv, py, mv))
    ref = A @ v
    scale = np.abs(A) @ np.abs(v)
    assert np.max(np.abs(y - ref) / scale) < 4e-16 * max(4, np.log2(n + 1))
    h.close()


# ---- whole solves against the reference's goldens ------------------------------------------------
@pytest.mark.parametrize("c", META, ids=[c["name"] for c in META])
def test_solver_matches_reference_golden(c):
    A, b, tab, x0 = case_inputs(c)
    out = run_gpu(c["solver"], A, b, tab, x0=x0, tol=c["tol"], max_mv=c["max_mv"], step=c["step"],
                  spg_seed=c["spg_seed"])
    check_against_golden(c, out)
    well_conditioned = c["gen"] != "wishart" and c.get("mu", 1.0) >= 0.1
    if well_conditioned and c["order_stable"] and c["converged"] and out["mv"] == c["mv"]:
        gold_res = float.fromhex(c["residual"])
        assert abs(out["residual"] - gold_res) <= 1e-6 * abs(gold_res) + 1e-12


def test_reference_test_matrix():
    """The reference's own test (tests/test_module.py:19-67): every solver on the five analytic
    problems, converged and within 1e-5 of the exact solution."""
    from ccqppy_b200 import problem_suite, solvers
    probs = [problem_suite.UnconstrainedSPD1(), problem_suite.UnconstrainedSPD2(), problem_suite.BoxConstrainedSPD(),
             problem_suite.ThinBoxConstrainedSPD(), problem_suite.ActiveBoxConstrainedSPD()]
    makers = [lambda: solvers.CCQPSolverPGD(1e-8, 10000, 0.1), lambda: solvers.CCQPSolverAPGD(1e-8, 10000),
              lambda: solvers.CCQPSolverAPGDAntiRelaxation(1e-8, 10000), lambda: solvers.CCQPSolverBBPGD(1e-8, 10000),
              lambda: solvers.CCQPSolverBBPGDf(1e-8, 10000), lambda: solvers.CCQPSolverSPG(1e-8, 10000),
              lambda: solvers.CCQPSolverMPRGP(1e-8, 10000), lambda: solvers.CCQPSolverMPRGPBB(1e-8, 10000)]
    for p in probs:
        for mk in makers:
            r = mk().solve(p.A, p.b, convex_proj_op=p.convex_proj_op)
            assert r.solution_converged
            assert np.linalg.norm(r.solution - p.exact_solution) < 1e-5


def test_readme_example_and_rng_side_effect(capsys):
    """README.md:30-51.  Also: the global NumPy RNG must end where the reference leaves it."""
    from ccqppy_b200 import solvers, solution_spaces as ss
    A = np.array([[2, -1, 0], [-1, 2, -1], [0, -1, 2]])
    exact_x = np.array([1, 0, 1])
    b = -A.dot(exact_x)
    op = ss.BoxProjOp(3, np.array([-2, -2, -4]), np.array([2, 2, 5]))
    np.random.seed(0)
    result = solvers.CCQPSolverSPG(1e-10, 5000).solve(A, b, convex_proj_op=op)
    after_gpu = np.random.random_sample()
    assert capsys.readouterr().out == "solving SPG\n"
    tab = pr.Table().add(pr.BOX, 3, np.array([-2., -2., -4.]), np.array([2., 2., 5.]))
    np.random.seed(0)
    o = orc.solve(pr.SPG, A, b, blocks=tab.blocks, params=tab.params, tol=1e-10, max_mv=5000)
    after_ref = np.random.random_sample()
    assert result.solution_converged and result.solution_num_matrix_vector_multiplications == o["mv"] == 92
    assert after_gpu == after_ref
    assert np.linalg.norm(result.solution - exact_x) < 1e-9
    assert isinstance(result.solution_time, float) and result.solution_time > 0


def test_mprgp_with_reference_cone_raises():
    from ccqppy_b200 import solvers, solution_spaces as ss
    A, b = pr.shift_problem(30, 0)
    op = ss.DisjointProjOp(*[ss.ConeProjOp(3)] * 10)
    s = solvers.CCQPSolverMPRGP(1e-6, 100)
    s.quiet = True
    with pytest.raises(NotImplementedError):
        s.solve(A, b, convex_proj_op=op)


def test_soc_cones_extension_against_oracle():
    """Config 5 flavour: contact-style friction cones (correct SOC projection; oracle = our CPU
    restatement, parity unpinned by the reference)."""
    n = 300
    A, b = pr.shift_problem(n, 2)
    tab = pr.soc3_table(n, 0.5)
    for solver in (pr.APGD, pr.BBPGD, pr.SPG, pr.MPRGP):
        o = orc.solve(solver, A, b, blocks=tab.blocks, params=tab.params, tol=1e-6, max_mv=3000,
                      uniforms=pr.spg_uniforms(3, 3000))
        g = run_gpu(solver, A, b, tab, tol=1e-6, max_mv=3000, spg_seed=3)
        assert g["converged"] == o["converged"]
        assert abs(g["mv"] - o["mv"]) <= max(1, round(0.02 * o["mv"]))
        if g["mv"] == o["mv"]:
            assert np.linalg.norm(g["solution"] - o["solution"]) <= 1e-9 * np.linalg.norm(o["solution"])


def test_torch_device_inputs_are_used_in_place():
    import torch
    n = 512
    A, b = pr.shift_problem(n, 9)
    tab = pr.box_table(n)
    ref = run_gpu(pr.BBPGD, A, b, tab, tol=1e-7, max_mv=500)
    Ad, bd = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
    s = make_solver(pr.BBPGD, 1e-7, 500)
    s.solve(Ad, bd, convex_proj_op=op_from_table(tab))
    assert s.solution.is_cuda
    assert np.array_equal(s.solution.cpu().numpy(), ref["solution"])
    assert s.solution_num_matrix_vector_multiplications == ref["mv"]
    # pinned host tensors go through the same host path as NumPy arrays
    Ap, bp = torch.from_numpy(A).pin_memory(), torch.from_numpy(b).pin_memory()
    s.solve(Ap, bp, convex_proj_op=op_from_table(tab))
    assert np.array_equal(np.asarray(s.solution), ref["solution"])


def test_determinism_run_to_run():
    n = 1500
    A, b = pr.shift_problem(n, 4, 0.01)
    tab = pr.mixed_table(n)
    outs = [run_gpu(pr.SPG, A, b, tab, tol=1e-6, max_mv=2000, spg_seed=2) for _ in range(3)]
    for o in outs[1:]:
        assert o["mv"] == outs[0]["mv"] and np.array_equal(o["solution"], outs[0]["solution"])


@pytest.mark.parametrize("solver", [pr.APGD, pr.BBPGD, pr.SPG, pr.MPRGP])
def test_sync_protocols_and_l2_slice_do_not_change_a_bit(solver, monkeypatch):
    """The closer-less rank-local grid sync adds the CTAs' partials in the closer's order, and the L2-resident slice of A
    only changes a cache policy: a solve is bit-identical with either sync protocol and with or without the slice."""
    n = 3600                                   # A = 104 MB: larger than the 96 MB above which A is streamed evict-first
    A, b = pr.shift_problem(n, 9, 0.05)
    tab = pr.mixed_table(n)
    base = run_gpu(solver, A, b, tab, tol=1e-7, max_mv=1500, spg_seed=3)
    for env in ({"CCQP_SYNC_CLOSERLESS": "0"}, {"CCQP_L2_RESIDENT_MB": "0"}, {"CCQP_SYNC_CLOSERLESS": "0", "CCQP_L2_RESIDENT_MB": "96"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        o = run_gpu(solver, A, b, tab, tol=1e-7, max_mv=1500, spg_seed=3)
        for k in env:
            monkeypatch.delenv(k)
        assert o["mv"] == base["mv"] and o["residual"] == base["residual"] and np.array_equal(o["solution"], base["solution"]), env


@pytest.mark.parametrize("solver", [pr.PGD, pr.APGD, pr.BBPGD, pr.SPG, pr.MPRGP])
def test_dense_4096_config2_against_oracle(solver):
    """Config 2: random dense SPD n=4096, box constraints, all five solvers, vs the oracle on the
    same inputs (the oracle finishes in seconds at this size)."""
    n = 4096
    A, b = pr.shift_problem(n, 0)
    tab = pr.box_table(n)
    step = 1.0 / np.abs(A).sum(axis=1).max()
    max_mv = 300
    o = orc.solve(solver, A, b, blocks=tab.blocks, params=tab.params, tol=1e-5, max_mv=max_mv, step_size=step,
                  uniforms=pr.spg_uniforms(0, max_mv))
    g = run_gpu(solver, A, b, tab, tol=1e-5, max_mv=max_mv, step=step, spg_seed=0)
    assert g["converged"] == o["converged"]
    assert abs(g["mv"] - o["mv"]) <= max(1, round(0.02 * o["mv"])), (g["mv"], o["mv"])
    if g["mv"] == o["mv"]:
        assert np.linalg.norm(g["solution"] - o["solution"]) <= 1e-9 * np.linalg.norm(o["solution"])


def test_shared_divisor_division():
    """The batched SPG kernel computes its three divisions by d.Ad with one shared reciprocal refinement
    (common.cuh div3_same_divisor); the quotients must equal IEEE division bit for bit, including zero,
    tiny, huge and non-finite operands."""
    import ctypes
    import torch
    from ccqppy_b200 import _capi
    g = torch.Generator(device="cuda").manual_seed(11)
    n = 1 << 22
    def rnd(scale_pow):
        m = torch.randn(n, generator=g, device="cuda", dtype=torch.float64)
        e = torch.randint(-scale_pow, scale_pow + 1, (n,), generator=g, device="cuda")
        return torch.ldexp(m, e)
    h = _capi.Handle()
    special = torch.tensor([0.0, -0.0, 1.0, -1.0, 5e-324, -5e-324, 2.2250738585072014e-308, 1.7976931348623157e308,
                            float("inf"), -float("inf"), float("nan"), 1e-300, 1e300, 3.0, 1.0 / 3.0], device="cuda",
                           dtype=torch.float64)
    for pw in (8, 300, 1060):
        a = [rnd(pw) for _ in range(3)]
        b = rnd(pw)
        k = special.numel()
        for j in range(3):
            a[j][:k * k] = special.repeat_interleave(k) if j == 0 else special.repeat(k)
        b[:k * k] = special.repeat(k)
        b[k * k:2 * k * k] = special.repeat_interleave(k)
        q = [torch.empty_like(b) for _ in range(3)]
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        torch.cuda.synchronize()      # a raw handle runs on its own stream, not on torch's
        _capi.check(h.h, h.lib.ccqp_debug_divide(h.h, P(a[0]), P(a[1]), P(a[2]), P(b), P(q[0]), P(q[1]), P(q[2]), n))
        for j in range(3):
            ref = a[j] / b
            same = (q[j].view(torch.int64) == ref.view(torch.int64)) | (torch.isnan(q[j]) & torch.isnan(ref))
            assert bool(same.all()), (pw, j, int((~same).sum()))
    h.close()


def test_benchmark_driver_matches_oracle_on_the_reference_study():
    """The counterpart of benchmarks/benchmark_random_ccqp.py (row f-2): a small slice of the reference's
    disjoint-constraint study; every (solver, constraint type, size, trial) cell must reproduce the oracle's
    mat-vec count and converged flag on the same Wishart problem."""
    from ccqppy_b200 import benchmark, solution_spaces as ss
    study = benchmark.benchmark_disjoint_constraints(problem_sizes=[3, 6, 12], num_random_trials=3)
    summ = study.summary()
    assert summ["sizes"] == [3, 6, 12] and len(summ["solvers"]) == 7 and len(summ["proj_types"]) == 5
    kinds = [pr.IDENTITY, pr.LOWER, pr.UPPER, pr.SPHERE, pr.BOX]
    ids = [pr.PGD, pr.APGD, pr.APGD_AR, pr.BBPGD, pr.BBPGDF, pr.MPRGP]      # SPG consumes the global RNG: checked elsewhere
    pos = [0, 1, 2, 3, 4, 6]
    cells = same_mv = same_flag = 0
    for pi, n in enumerate([3, 6, 12]):
        for trial in range(3):
            A, b = study.generate_random_convex_quadratic_func(n, trial)
            for ti, kind in enumerate(kinds):
                tab = pr.Table()
                for _ in range(n // 3):
                    if kind == pr.IDENTITY: tab.add(kind, 3)
                    elif kind == pr.LOWER: tab.add(kind, 3, -1.0)
                    elif kind in (pr.UPPER, pr.SPHERE): tab.add(kind, 3, 1.0)
                    else: tab.add(kind, 3, -1.0, 1.0)
                for sp, sid in zip(pos, ids):
                    o = orc.solve(sid, A, b, blocks=tab.blocks, params=tab.params, tol=1e-5, max_mv=5000)
                    cells += 1
                    same_flag += bool(study.problem_converged[sp, ti, pi, trial]) == o["converged"]
                    same_mv += study.problem_num_matrix_vector_mults[sp, ti, pi, trial] == o["mv"]
    # Wishart(df = n) at n <= 12 can be nearly singular (the reference's own study): BB / restart decisions
    # then flip on rounding-level differences, so a minority of cells takes a different number of steps
    assert cells == 270 and same_flag >= 0.97 * cells and same_mv >= 0.85 * cells, (cells, same_flag, same_mv)


@pytest.mark.parametrize("n", [5, 63, 1023, 2051])
def test_device_matrix_with_unaligned_rows(n):
    """A CUDA tensor whose rows are not 32-byte aligned (odd n) is borrowed in place and goes through the
    64-bit-load path of the mat-vec; results must equal the host-upload path (which pads the rows)."""
    import torch
    A, b = pr.shift_problem(n, 21)
    tab = pr.box_table(n)
    for solver in (pr.BBPGD, pr.SPG, pr.MPRGP):
        host = run_gpu(solver, A, b, tab, tol=1e-7, max_mv=800, spg_seed=4)
        s = make_solver(solver, 1e-7, 800)
        np.random.seed(4)
        s.solve(torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda(), convex_proj_op=op_from_table(tab))
        assert s.solution_num_matrix_vector_multiplications == host["mv"]
        np.testing.assert_allclose(s.solution.cpu().numpy(), host["solution"], rtol=1e-12, atol=1e-14)
    # leading dimension larger than n (a window of a bigger device matrix), straight through the C-ABI
    import ctypes
    from ccqppy_b200 import _capi
    big = torch.zeros((n, n + 3), device="cuda", dtype=torch.float64)
    big[:, :n] = torch.from_numpy(A).cuda()
    v = np.random.default_rng(1).standard_normal(n)
    torch.cuda.synchronize()          # a raw handle runs on its own stream, not on torch's
    h = _capi.Handle()
    _capi.check(h.h, h.lib.ccqp_set_matrix(h.h, ctypes.c_void_p(big.data_ptr()), n, n + 3, 0, n, _capi.MEM_DEVICE))
    y = np.empty(n)
    _capi.check(h.h, h.lib.ccqp_gemv(h.h, ctypes.c_void_p(v.ctypes.data), ctypes.c_void_p(y.ctypes.data), _capi.MEM_HOST))
    h.close()
    np.testing.assert_allclose(y, A @ v, rtol=1e-12, atol=1e-12)


def test_pipelined_solves_equal_synchronous_solves():
    """ccqp_solve_async / ccqp_solve_wait through SolvePipeline: same kernel, same answers, any order of completion."""
    import torch
    from ccqppy_b200.pipeline import SolvePipeline
    from ccqppy_b200 import solvers
    probs = []
    for i, n in enumerate([700, 1500, 64, 1500, 2048]):
        A, b = pr.shift_problem(n, 30 + i)
        probs.append((torch.from_numpy(A).pin_memory(), torch.from_numpy(b).pin_memory(), pr.mixed_table(n), pr.spg_uniforms(i, 2000)))
    for solver in (pr.SPG, pr.BBPGD, pr.MPRGP):
        pipe = SolvePipeline(make_solver(solver, 1e-7, 2000), depth=2)
        for A, b, tab, uni in probs:
            pipe.submit(A, b, convex_proj_op=op_from_table(tab), uniforms=uni if solver == pr.SPG else None)
        got = pipe.results()
        pipe.close()
        assert len(got) == len(probs)
        for (A, b, tab, uni), r in zip(probs, got):
            s = make_solver(solver, 1e-7, 2000)
            s.solve(A, b, convex_proj_op=op_from_table(tab), uniforms=uni)
            assert r.solution_num_matrix_vector_multiplications == s.solution_num_matrix_vector_multiplications
            assert r.solution_converged == s.solution_converged and r.solution_residual == s.solution_residual
            assert np.array_equal(np.asarray(r.solution), np.asarray(s.solution))
    with pytest.raises(ValueError):
        SolvePipeline(solvers.CCQPSolverSPG(1e-6, 100)).submit(probs[0][0], probs[0][1])


def test_benchmark_driver_with_the_gpu_generator():
    """Row f-2: the study at a size where the GPU matters, Hessians generated on the device and used in place
    (no host GEMM, no PCIe copy of A).  Checked through the reference's own residual recomputed with torch."""
    import torch
    from ccqppy_b200 import benchmark, solution_spaces as ss, solvers
    n = 1536
    ops = [[ss.BoxProjOp(n)], [ss.DisjointProjOp(*[ss.SphereProjOp(3)] * (n // 3))]]
    study = benchmark.BenchmarkRandomCCQP(2, [solvers.CCQPSolverBBPGD(1e-5, 5000), solvers.CCQPSolverMPRGP(1e-5, 5000)], ops,
                                          generator="gpu").run()
    A, b = study.generate_random_convex_quadratic_func(n, 0)
    assert A.is_cuda and b.is_cuda and A.dtype == torch.float64
    A2, _ = study.generate_random_convex_quadratic_func(n, 0)
    assert torch.equal(A, A2)                                                    # seeded: a study is reproducible
    assert float((A - A.t()).abs().max()) == 0.0 or float((A - A.t()).abs().max()) < 1e-9 * float(A.abs().max())
    assert study.problem_converged.shape == (2, 2, 1, 2)
    s = solvers.CCQPSolverBBPGD(1e-5, 5000)
    s.quiet = True
    s.solve(A, b, convex_proj_op=ops[0][0])
    assert s.solution.is_cuda                                                    # stayed on the device end to end
    assert int(s.solution_num_matrix_vector_multiplications) == int(study.problem_num_matrix_vector_mults[0, 0, 0, 0])
    if s.solution_converged:
        x = s.solution
        g = A @ x + b
        assert float((x - (x - 1e-6 * g).clamp(-1, 1)).norm()) / (3 * n * 1e-6) < 1e-5 * 1.01
    summ = study.summary()
    assert summ["sizes"] == [n] and len(summ["solvers"]) == 2
    # "auto" keeps the reference's host generator for the reference's own (tiny) sizes
    small = benchmark.BenchmarkRandomCCQP(1, [], [[ss.BoxProjOp(6)]])
    As, _ = small.generate_random_convex_quadratic_func(6, 0)
    assert isinstance(As, np.ndarray)


def test_projected_gradient_matches_reference_goldens_and_behaviours():
    """Row f-4: `op.projected_gradient(x, g)` on the GPU (ccqp_projected_gradient) against the goldens generated from the
    live reference (solution_spaces.py:162-184, 238-260, 324-347, 527-538): bit for bit; and the reference's behaviour for
    the operators that do not implement it (Identity -> None, Sphere -> NotImplementedError, Cone -> AttributeError,
    Disjoint with an Identity member -> TypeError)."""
    from test_oracle_golden import PG, PG_TABLES, PG_BEHAVIOUR
    from ccqppy_b200 import solution_spaces as ss
    for name, tab in PG_TABLES.items():
        op = op_from_table(tab)
        for x, g, f, c in zip(PG[name + "/x"], PG[name + "/g"], PG[name + "/free"], PG[name + "/chopped"]):
            gf, gc = op.projected_gradient(x, g)
            assert np.array_equal(gf, f) and np.array_equal(gc, c), name
    x, g = np.linspace(-1, 1, 5), np.linspace(2, -2, 5)
    ops = {"identity": ss.IdentityProjOp(5), "sphere": ss.SphereProjOp(5), "cone": ss.ConeProjOp(5),
           "disjoint_with_identity": ss.DisjointProjOp(ss.BoxProjOp(3), ss.IdentityProjOp(2))}
    for name, op in ops.items():
        try:
            r = op.projected_gradient(x, g)
            got = "None" if r is None else "value"
        except Exception as e:      # noqa: BLE001
            got = type(e).__name__
        assert got == PG_BEHAVIOUR[name], name
    with pytest.raises(NotImplementedError):
        ss.ConeProjOp(5).proximal_gradient(x, g)
    with pytest.raises(NotImplementedError):
        ss.DisjointProjOp(ss.BoxProjOp(2), ss.SphereProjOp(3)).projected_gradient(x, g)
    # large vector, NaN gradient entries propagate like np.min((nan, 0))
    n = 5000
    rng = np.random.default_rng(3)
    op = ss.BoxProjOp(n)
    xb = np.clip(2 * rng.standard_normal(n), -1, 1)
    gb = rng.standard_normal(n)
    gb[::97] = np.nan
    tab = pr.box_table(n)
    of, oc = orc.projected_gradient(tab.blocks, tab.params, xb, gb)
    gf, gc = op.projected_gradient(xb, gb)
    assert np.array_equal(gf, of, equal_nan=True) and np.array_equal(gc, oc, equal_nan=True)
