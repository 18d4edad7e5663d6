"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports exactly what
include/ccqp_b200.h declares, the Python surface mirrors the reference's names/ctors/strings,
and the product path fails loudly (no CPU fallback) when there is no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import ccqppy_b200
from ccqppy_b200 import _capi, problem_suite, solution_spaces as ss, solvers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "ccqp_b200.h")).read()
    declared = set(re.findall(r"\b(ccqp_[a-z0-9_]+)\s*\(", header))
    declared -= {"ccqp_status"}
    assert declared == set(_capi.EXPORTS), declared ^ set(_capi.EXPORTS)
    lib = _capi.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ccqp_abi_version() == 3
    assert lib.ccqp_status_string(0) == b"ok"
    assert b"Cone normal" in lib.ccqp_status_string(6)


def test_struct_sizes_match_header():
    assert ctypes.sizeof(_capi.Block) == 32
    assert ctypes.sizeof(_capi.Params) == 56
    assert ctypes.sizeof(_capi.Result) == 72


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_without_gpu():
    lib = _capi.load()
    h = ctypes.c_void_p()
    assert lib.ccqp_create(ctypes.byref(h), -1) == 2          # CCQP_ERR_NO_DEVICE
    A = np.eye(3)
    with pytest.raises(_capi.CCQPError):
        solvers.CCQPSolverBBPGD(1e-8, 100).solve(A, np.ones(3))
    with pytest.raises(_capi.CCQPError):
        ss.BoxProjOp(3)(np.zeros(3))


def test_unsupported_inputs_raise_instead_of_falling_back():
    class OnlyDot:
        def dot(self, v):
            return v
    with pytest.raises(TypeError):
        solvers.CCQPSolverPGD(1e-8, 10).solve(OnlyDot(), np.ones(3))
    with pytest.raises(TypeError):
        solvers.CCQPSolverPGD(1e-8, 10).solve(np.eye(3), np.ones(3), convex_proj_op=lambda x: x)


def test_python_surface_mirrors_reference():
    # names and strings: SURVEY.md section 8(b), solvers.py:174,347,537,673,823,979,1204
    expect = {"CCQPSolverPGD": "PGD", "CCQPSolverAPGD": "APGD", "CCQPSolverAPGDAntiRelaxation": "Anti-relaxation APGD",
              "CCQPSolverBBPGD": "BBGPD", "CCQPSolverBBPGDf": "BBPDGf", "CCQPSolverSPG": "SPG-QP",
              "CCQPSolverMPRGP": "MPRGP"}
    for cls, name in expect.items():
        s = getattr(ccqppy_b200, cls)(1e-6, 50)
        assert s.name == name
        assert s.desired_residual_tol == 1e-6 and s.max_matrix_vector_multiplications == 50
        for prop in ("solution", "solution_residual", "solution_converged", "solution_time",
                     "solution_num_matrix_vector_multiplications"):
            assert getattr(s, prop) is None
    assert ccqppy_b200.CCQPSolverMPRGPBB is ccqppy_b200.CCQPSolverMPRGP
    assert solvers.CCQPSolverPGD(1e-3).step_size == 0.01
    assert solvers.CCQPSolverPGD(1e-3).max_matrix_vector_multiplications == np.inf
    spg = solvers.CCQPSolverSPG(1e-3)
    assert (spg.m, spg.t, spg.sigma1, spg.sigma2) == (5, 0.5, 0.01, 0.5)
    names = {"IdentityProjOp": "Identity", "LowerBoundProjOp": "Lower Bound", "UpperBoundProjOp": "Upper Bound",
             "BoxProjOp": "Box", "SphereProjOp": "Sphere", "ConeProjOp": "Cone"}
    for cls, name in names.items():
        op = getattr(ccqppy_b200, cls)(4)
        assert op.name == name and op.embedded_dimension == 4 and op.dim == 4
    assert np.all(ss.BoxProjOp(3).lower_bound == -1) and np.all(ss.BoxProjOp(3).upper_bound == 1)
    assert ss.SphereProjOp(3).radius == 1 and ss.ConeProjOp(3).aspect_ratio == 1
    d = ss.DisjointProjOp(ss.BoxProjOp(2), ss.SphereProjOp(3), ss.IdentityProjOp(1))
    assert d.name == "DisjointUnion" and d.embedded_dimension == 6 and len(d.proj_ops) == 3


def test_descriptor_flattening():
    sph = ss.SphereProjOp(3, 2.0)
    d = ss.DisjointProjOp(ss.BoxProjOp(2, np.array([0., 1.]), np.array([2., 3.])), sph, sph,
                          ss.DisjointProjOp(ss.LowerBoundProjOp(2), ss.IdentityProjOp(1)))
    blocks, params, rows = d.descriptor()
    assert rows == [(3, 0, 2, 0), (4, 2, 3, 4), (4, 5, 3, 4), (1, 8, 2, 5), (0, 10, 1, 7)]
    assert params.tolist() == [0., 1., 2., 3., 2.0, -1., -1.]
    assert len(blocks) == 5 and blocks[3].offset == 8


def test_problem_suite_fixtures():
    for cls in ("UnconstrainedSPD1", "UnconstrainedSPD2", "BoxConstrainedSPD", "ThinBoxConstrainedSPD",
                "ActiveBoxConstrainedSPD"):
        p = getattr(problem_suite, cls)()
        assert p.number_of_unknowns == 3 and p.A.shape == (3, 3) and p.A.dtype.kind == "i"
        assert p.convex_proj_op.embedded_dimension == 3
    p = problem_suite.ActiveBoxConstrainedSPD()
    assert p.b.tolist() == [-1, 0, -1] and p.exact_solution.tolist() == [9, 9, 9]
    assert problem_suite.BoxConstrainedSPD().b.tolist() == [-2, 2, -2]
