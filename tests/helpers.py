"""Glue shared by the GPU tests: build ccqppy_b200 operators from the flat test tables and run
a golden case through the public API."""
import numpy as np

import problems as pr
from ccqppy_b200 import solution_spaces as ss, solvers


def op_from_table(tab):
    ops = []
    par = tab.params
    for kind, off, dim, poff in tab.blocks:
        kind, dim, poff = int(kind), int(dim), int(poff)
        if kind == pr.IDENTITY:
            ops.append(ss.IdentityProjOp(dim))
        elif kind == pr.LOWER:
            ops.append(ss.LowerBoundProjOp(dim, par[poff:poff + dim]))
        elif kind == pr.UPPER:
            ops.append(ss.UpperBoundProjOp(dim, par[poff:poff + dim]))
        elif kind == pr.BOX:
            ops.append(ss.BoxProjOp(dim, par[poff:poff + dim], par[poff + dim:poff + 2 * dim]))
        elif kind == pr.SPHERE:
            ops.append(ss.SphereProjOp(dim, par[poff]))
        elif kind == pr.CONE_REF:
            ops.append(ss.ConeProjOp(dim, par[poff]))
        elif kind == pr.SOC:
            ops.append(ss.SOCProjOp(dim, par[poff]))
    return ops[0] if len(ops) == 1 else ss.DisjointProjOp(*ops)


def make_solver(solver, tol, max_mv, step=0.01):
    S = solvers
    s = {pr.PGD: lambda: S.CCQPSolverPGD(tol, max_mv, step), pr.APGD: lambda: S.CCQPSolverAPGD(tol, max_mv),
         pr.APGD_AR: lambda: S.CCQPSolverAPGDAntiRelaxation(tol, max_mv), pr.BBPGD: lambda: S.CCQPSolverBBPGD(tol, max_mv),
         pr.BBPGDF: lambda: S.CCQPSolverBBPGDf(tol, max_mv), pr.SPG: lambda: S.CCQPSolverSPG(tol, max_mv),
         pr.MPRGP: lambda: S.CCQPSolverMPRGP(tol, max_mv)}[solver]()
    s.quiet = True
    return s


def run_gpu(solver, A, b, tab, x0=None, tol=1e-8, max_mv=np.inf, step=0.01, spg_seed=0):
    s = make_solver(solver, tol, max_mv, step)
    np.random.seed(spg_seed)
    s.solve(A, b, x0=x0, convex_proj_op=op_from_table(tab))
    return dict(solution=np.asarray(s.solution), residual=s.solution_residual, converged=s.solution_converged,
                mv=s.solution_num_matrix_vector_multiplications, gemv=s.solution_gemv_count, solver=s)
