"""Analytic 3x3 known-answer problems: mirror of the reference's `ccqppy.problem_suite`
(/root/reference/src/ccqppy/problem_suite.py:42-177) with the same class names and properties
(`number_of_unknowns`, `A`, `b`, `convex_proj_op`, `exact_solution`).

All five use the tridiagonal Hessian tridiag(-1, 2, -1) (integer dtype on purpose, as in the
reference, :54) and b = -A x* so that the minimiser of 1/2 x'Ax + b'x is known."""
import numpy as np

from . import solution_spaces as ss

_A = ((2, -1, 0), (-1, 2, -1), (0, -1, 2))


class TestProblemBase:
    __test__ = False          # not a pytest class
    _unconstrained_minimiser = (1, 0, 1)
    _exact = (1, 0, 1)

    def __init__(self):
        pass

    @property
    def number_of_unknowns(self):
        return 3

    @property
    def A(self):
        return np.array(_A)

    @property
    def b(self):
        return -self.A.dot(np.array(self._unconstrained_minimiser))

    @property
    def exact_solution(self):
        return np.array(self._exact)

    @property
    def convex_proj_op(self):
        raise NotImplementedError


class UnconstrainedSPD1(TestProblemBase):
    """problem_suite.py:42-66"""

    @property
    def convex_proj_op(self):
        return ss.IdentityProjOp(3)


class UnconstrainedSPD2(TestProblemBase):
    """problem_suite.py:69-93: the same set written as a disjoint union of three 1-D identities."""

    @property
    def convex_proj_op(self):
        return ss.DisjointProjOp(ss.IdentityProjOp(1), ss.IdentityProjOp(1), ss.IdentityProjOp(1))


class BoxConstrainedSPD(TestProblemBase):
    """problem_suite.py:96-121: box [0,2]^3, unconstrained minimiser inside."""

    @property
    def convex_proj_op(self):
        return ss.BoxProjOp(3, lower_bound=np.array([0, 0, 0]), upper_bound=np.array([2, 2, 2]))


class ThinBoxConstrainedSPD(TestProblemBase):
    """problem_suite.py:124-149: a thin box around the minimiser."""

    @property
    def convex_proj_op(self):
        return ss.BoxProjOp(3, lower_bound=np.array([-10, -0.1, 0.9]), upper_bound=np.array([10, 0.1, 1.1]))


class ActiveBoxConstrainedSPD(TestProblemBase):
    """problem_suite.py:152-177: box [9,10]^3, every constraint active at the solution."""
    _unconstrained_minimiser = (1, 1, 1)
    _exact = (9, 9, 9)

    @property
    def convex_proj_op(self):
        return ss.BoxProjOp(3, lower_bound=np.array([9, 9, 9]), upper_bound=np.array([10, 10, 10]))
