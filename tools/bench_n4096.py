"""Config 2 (n = 4096, Box, every solver) on the GPU only: kernel time and algorithmic GB/s per solver, one line.
    [CCQP_L2_RESIDENT_MB=..] python tools/bench_n4096.py [n]"""
import json
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import problems as pr
from helpers import op_from_table, make_solver

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
A, b = pr.shift_problem(n, 0)
Ad, bd = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
step = 1.0 / np.abs(A).sum(axis=1).max()
op = op_from_table(pr.box_table(n))
uni = torch.from_numpy(pr.spg_uniforms(0, 2000)).cuda()
out = {}
for solver in range(7):
    best = None
    for _ in range(4):
        s = make_solver(solver, 1e-5, 2000, step)
        s.solve(Ad, bd, convex_proj_op=op, uniforms=uni)
        if best is None or s.solution_gpu_time < best[0]:
            best = (s.solution_gpu_time, s.solution_hbm_bytes, s.solution_gemv_count, s.solution_num_matrix_vector_multiplications)
    out[pr.SOLVER_NAMES[solver]] = dict(GBps=round(best[1] / best[0] / 1e9), us_per_gemv=round(1e6 * best[0] / best[2], 2), mv=int(best[3]))
print(json.dumps(out))
