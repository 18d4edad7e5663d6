"""Ensemble benchmark of the solvers on random CCQPs: the B200 counterpart of the reference's
`benchmarks/benchmark_random_ccqp.py`.

Same driver surface -- `BenchmarkRandomCCQP(num_random_trials, solvers_to_benchmark,
convex_proj_ops_to_benchmark).run()`, the Wishart problem generator of
benchmark_random_ccqp.py:36-63, the two ready-made studies of :155-216 -- but the results come
back as arrays / one JSON document instead of matplotlib figures (:104-152 are presentation and
out of scope), and the problem sizes are arguments, so the same study can be run at sizes where a
GPU matters.

    python -m ccqppy_b200.benchmark disjoint --sizes 3 6 9 12 --trials 100 > study.json
    python -m ccqppy_b200.benchmark single --sizes 256 1024 4096 --trials 3 --generator gpu

`generator="gpu"`: the Wishart Hessian (an n^3 product, the step BEFORE the path) is drawn and multiplied on
the device in fp64 and handed to `solve()` as a CUDA tensor, which the solver uses in place: no host GEMM, no
PCIe transfer of A, so the study runs at sizes where a GPU matters.  `"host"` is the reference's NumPy/SciPy
generator (benchmark_random_ccqp.py:59-62); `"auto"` picks the device from n = 1024 on.
"""
import argparse
import json
import sys

import numpy as np

from . import solution_spaces as ss
from . import solvers


class BenchmarkRandomCCQP:
    """benchmark_random_ccqp.py:15-102.  `convex_proj_ops_to_benchmark[t][s]` is the operator of
    constraint type t at problem size s (all types share the sizes of type 0, :25-29)."""

    def __init__(self, num_random_trials, solvers_to_benchmark, convex_proj_ops_to_benchmark, quiet=True, generator="auto",
                 device=None):
        if generator not in ("auto", "host", "gpu"):
            raise ValueError("generator must be 'auto', 'host' or 'gpu'")
        self.generator = generator
        self.device = device
        self.num_trials = int(num_random_trials)
        self.solvers_to_benchmark = list(solvers_to_benchmark)
        self.convex_proj_ops_to_benchmark = [list(row) for row in convex_proj_ops_to_benchmark]
        self.problem_sizes = np.array([op.embedded_dimension for op in self.convex_proj_ops_to_benchmark[0]], dtype=int)
        self.quiet = quiet
        self._problem_residual = None
        self._problem_converged = None
        self._problem_time = None
        self._problem_gpu_time = None
        self._problem_num_matrix_vector_mults = None

    def _on_gpu(self, problem_size):
        return self.generator == "gpu" or (self.generator == "auto" and problem_size >= 1024)

    def generate_random_convex_quadratic_func(self, problem_size, seed=1234):
        """A ~ Wishart(df = n, I_n) seeded with `seed`, b = -A x*, x* = 1 - 2 U (:59-62).  The
        reference draws x* from the unseeded global RNG; here it is seeded too, so a
        study is reproducible.  On the device A = G G^T with G an n x n standard normal matrix from
        torch's seeded generator: the same distribution (that IS the Wishart(n, I) construction), a
        different random stream than SciPy's."""
        if self._on_gpu(problem_size):
            import torch
            dev = torch.device(self.device if self.device is not None else "cuda")
            g = torch.Generator(device=dev).manual_seed(int(seed))
            G = torch.randn((problem_size, problem_size), generator=g, device=dev, dtype=torch.float64)
            A = G @ G.t()
            del G
            x_star = 1 - 2 * torch.rand(problem_size, generator=g, device=dev, dtype=torch.float64)
            return A, -(A @ x_star)
        rng = np.random.RandomState(seed)
        try:
            from scipy.stats import wishart
            A = np.atleast_2d(wishart.rvs(problem_size, np.eye(problem_size), size=1, random_state=rng))
        except ImportError:
            G = rng.standard_normal((problem_size, problem_size))
            A = G @ G.T
        x_star = 1 - 2 * rng.rand(problem_size)
        return A, -A.dot(x_star)

    def run(self):
        shape = (len(self.solvers_to_benchmark), len(self.convex_proj_ops_to_benchmark), len(self.problem_sizes),
                 self.num_trials)
        self._problem_residual = np.zeros(shape)
        self._problem_converged = np.zeros(shape, dtype=int)
        self._problem_time = np.zeros(shape)
        self._problem_gpu_time = np.zeros(shape)
        self._problem_num_matrix_vector_mults = np.zeros(shape, dtype=int)
        problems = {}
        for si, solver in enumerate(self.solvers_to_benchmark):
            solver.quiet = self.quiet
            for ti, ops in enumerate(self.convex_proj_ops_to_benchmark):
                for pi, op in enumerate(ops):
                    for trial in range(self.num_trials):
                        key = (int(self.problem_sizes[pi]), trial)
                        if key not in problems:
                            problems[key] = self.generate_random_convex_quadratic_func(*key)
                        A, b = problems[key]
                        r = solver.solve(A, b, convex_proj_op=op)
                        idx = (si, ti, pi, trial)
                        self._problem_residual[idx] = r.solution_residual
                        self._problem_converged[idx] = r.solution_converged
                        self._problem_time[idx] = r.solution_time
                        self._problem_gpu_time[idx] = r.solution_gpu_time
                        self._problem_num_matrix_vector_mults[idx] = r.solution_num_matrix_vector_multiplications
        return self

    # -- results (instead of the figures of :104-152) ------------------------------------------------
    @property
    def problem_residual(self):
        return self._problem_residual

    @property
    def problem_converged(self):
        return self._problem_converged

    @property
    def problem_time(self):
        return self._problem_time

    @property
    def problem_num_matrix_vector_mults(self):
        return self._problem_num_matrix_vector_mults

    def summary(self):
        """Means over the trials, indexed [solver][constraint type][size] -- what the reference plots."""
        out = dict(sizes=self.problem_sizes.tolist(), trials=self.num_trials, solvers=[], proj_types=[])
        out["proj_types"] = [ops[0].name if not isinstance(ops[0], ss.DisjointProjOp)
                             else "DisjointUnion(%s)" % ops[0].proj_ops[0].name for ops in self.convex_proj_ops_to_benchmark]
        for si, solver in enumerate(self.solvers_to_benchmark):
            out["solvers"].append(dict(
                name=solver.name,
                mean_wall_s=self._problem_time[si].mean(axis=-1).tolist(),
                mean_gpu_s=self._problem_gpu_time[si].mean(axis=-1).tolist(),
                mean_mat_vecs=self._problem_num_matrix_vector_mults[si].mean(axis=-1).tolist(),
                mean_residual=self._problem_residual[si].mean(axis=-1).tolist(),
                converged_fraction=self._problem_converged[si].mean(axis=-1).tolist()))
        return out

    def process_results(self, file=None):
        json.dump(self.summary(), file or sys.stdout, indent=1)
        print("", file=file or sys.stdout)


def _solver_set(tol, max_mv, with_mprgp):
    s = [solvers.CCQPSolverPGD(tol, max_mv), solvers.CCQPSolverAPGD(tol, max_mv),
         solvers.CCQPSolverAPGDAntiRelaxation(tol, max_mv), solvers.CCQPSolverBBPGD(tol, max_mv),
         solvers.CCQPSolverBBPGDf(tol, max_mv), solvers.CCQPSolverSPG(tol, max_mv)]
    if with_mprgp:
        s.append(solvers.CCQPSolverMPRGP(tol, max_mv))
    return s


def benchmark_single_constraint(problem_sizes=None, num_random_trials=10, desired_tol=1e-5, max_mv_mults=5000, generator="auto"):
    """benchmark_random_ccqp.py:155-183: one operator over the whole vector, six solvers."""
    sizes = np.linspace(2, 12, 10, dtype=int) if problem_sizes is None else np.asarray(problem_sizes, dtype=int)
    ops = [[kind(int(d)) for d in sizes] for kind in
           (ss.IdentityProjOp, ss.LowerBoundProjOp, ss.UpperBoundProjOp, ss.SphereProjOp, ss.BoxProjOp)]
    return BenchmarkRandomCCQP(num_random_trials, _solver_set(desired_tol, max_mv_mults, False), ops, generator=generator).run()


def benchmark_disjoint_constraints(problem_sizes=None, num_random_trials=100, desired_tol=1e-5, max_mv_mults=5000,
                                   generator="auto"):
    """benchmark_random_ccqp.py:186-216: disjoint unions of 3-wide blocks, seven solvers."""
    sizes = np.arange(3, 13, 3) if problem_sizes is None else np.asarray(problem_sizes, dtype=int)
    ops = [[ss.DisjointProjOp(*[kind(3)] * (int(d) // 3)) for d in sizes] for kind in
           (ss.IdentityProjOp, ss.LowerBoundProjOp, ss.UpperBoundProjOp, ss.SphereProjOp, ss.BoxProjOp)]
    return BenchmarkRandomCCQP(num_random_trials, _solver_set(desired_tol, max_mv_mults, True), ops, generator=generator).run()


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("study", choices=["single", "disjoint"])
    ap.add_argument("--sizes", type=int, nargs="*")
    ap.add_argument("--trials", type=int)
    ap.add_argument("--tol", type=float, default=1e-5)
    ap.add_argument("--max-mv", type=int, default=5000)
    ap.add_argument("--generator", choices=["auto", "host", "gpu"], default="auto")
    a = ap.parse_args(argv)
    fn = benchmark_single_constraint if a.study == "single" else benchmark_disjoint_constraints
    kw = dict(desired_tol=a.tol, max_mv_mults=a.max_mv, generator=a.generator)
    if a.sizes:
        kw["problem_sizes"] = a.sizes
    if a.trials:
        kw["num_random_trials"] = a.trials
    fn(**kw).process_results()


if __name__ == "__main__":
    main()
