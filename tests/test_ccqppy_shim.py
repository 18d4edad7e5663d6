"""The drop-in import name: `from ccqppy import solvers, solution_spaces, problem_suite` (what the reference's
tests/test_module.py:5-8 and README.md:30-33 do) resolves to the CUDA-backed modules of ccqppy_b200."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fresh_shim():
    """`import ccqppy` from this repository (another test module imports the REFERENCE under the same name)."""
    for k in [m for m in sys.modules if m == "ccqppy" or m.startswith("ccqppy.")]:
        del sys.modules[k]
    if sys.path[0] != ROOT:
        sys.path.insert(0, ROOT)
    import ccqppy
    assert os.path.dirname(os.path.dirname(os.path.realpath(ccqppy.__file__))) == os.path.realpath(ROOT)
    return ccqppy


def test_ccqppy_names_resolve_to_the_b200_modules():
    ccqppy = _fresh_shim()
    import ccqppy_b200
    from ccqppy import problem_suite, solvers
    from ccqppy import solution_spaces as ss
    import ccqppy.solvers as by_path
    assert solvers is ccqppy_b200.solvers is by_path and ss is ccqppy_b200.solution_spaces
    assert problem_suite is ccqppy_b200.problem_suite
    # the reference's __init__ star-exports both modules (ccqppy/__init__.py:3-4)
    for name in ("CCQPSolverPGD", "CCQPSolverAPGD", "CCQPSolverAPGDAntiRelaxation", "CCQPSolverBBPGD", "CCQPSolverBBPGDf",
                 "CCQPSolverSPG", "CCQPSolverMPRGP", "IdentityProjOp", "LowerBoundProjOp", "UpperBoundProjOp", "BoxProjOp",
                 "SphereProjOp", "ConeProjOp", "DisjointProjOp"):
        assert hasattr(ccqppy, name), name


@pytest.mark.gpu
def test_the_reference_test_module_through_the_ccqppy_name():
    """The body of /root/reference/tests/test_module.py (:11-16 TestSolutionSpaces.test_identity, :19-67
    TestSolversAgainstSimpleProblems.test_APGD) run through `import ccqppy`, unittest-style assertions included."""
    import unittest
    _fresh_shim()
    from ccqppy import solution_spaces as ss
    from ccqppy import solvers
    from ccqppy import problem_suite
    case = unittest.TestCase()
    op = ss.IdentityProjOp(10)
    x_random = np.random.rand(10)
    case.assertTrue(np.all(op(x_random) == x_random))
    problems = [problem_suite.UnconstrainedSPD1(), problem_suite.UnconstrainedSPD2(), problem_suite.BoxConstrainedSPD(),
                problem_suite.ThinBoxConstrainedSPD(), problem_suite.ActiveBoxConstrainedSPD()]
    for prob in problems:
        for make in (lambda: solvers.CCQPSolverPGD(1e-8, 10000, 0.1), lambda: solvers.CCQPSolverAPGD(1e-8, 10000),
                     lambda: solvers.CCQPSolverAPGDAntiRelaxation(1e-8, 10000), lambda: solvers.CCQPSolverBBPGD(1e-8, 10000),
                     lambda: solvers.CCQPSolverBBPGDf(1e-8, 10000), lambda: solvers.CCQPSolverSPG(1e-8, 10000),
                     lambda: solvers.CCQPSolverMPRGP(1e-8, 10000)):
            result = make().solve(prob.A, prob.b, convex_proj_op=prob.convex_proj_op)
            case.assertTrue(result.solution_converged)
            case.assertTrue(np.linalg.norm(result.solution - prob.exact_solution) < 1e-5)
