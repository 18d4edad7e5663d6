import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, problems as pr
from helpers import op_from_table, make_solver
from test_gpu_fullsize import gpu_problem
n = 16384
A, b = gpu_problem(n, seed=3)
for name, tab in (("box", pr.box_table(n)), ("sphere3", pr.sphere3_table(n))):
    op = op_from_table(tab)
    for it in range(2):
        s = make_solver(pr.MPRGP, 1e-7, 3000); s.solve(A, b, convex_proj_op=op)
    print(name, s.solution_num_matrix_vector_multiplications, s.solution_gemv_count, 1e3 * s.solution_gpu_time, flush=True)
