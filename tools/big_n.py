"""One solve at sizes near the memory of a single B200 (A = 34 GB at n = 65536, 51 GB at n = 80000): checks the
64-bit addressing of the kernels and that the roofline fraction holds when A is 400x the L2."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import problems as pr
from helpers import make_solver
from ccqppy_b200 import solution_spaces as ss
from test_gpu_fullsize import gpu_problem, residual

for n in [int(a) for a in sys.argv[1:]] or [65536]:
    A, b = gpu_problem(n)
    torch.cuda.empty_cache()
    op = ss.BoxProjOp(n)
    for solver in (pr.BBPGD, pr.SPG):
        s = make_solver(solver, 1e-6, 2000)
        s.solve(A, b, convex_proj_op=op, uniforms=torch.from_numpy(pr.spg_uniforms(0, 2000)).cuda())
        x = s.solution
        print("n=%d %s: mv %d conv %s kernel %.1f ms  %.0f GB/s  feasible %s  residual(torch) %.2e" %
              (n, pr.SOLVER_NAMES[solver], s.solution_num_matrix_vector_multiplications, s.solution_converged,
               1e3 * s.solution_gpu_time, s.solution_hbm_bytes / s.solution_gpu_time / 1e9, bool((x.abs() <= 1).all()),
               residual(A, b, x, lambda v: v.clamp(-1, 1))), flush=True)
    del A, b
    torch.cuda.empty_cache()
