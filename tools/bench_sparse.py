"""Operator-form (CSR) Hessian: solver throughput on a banded-plus-random sparse SPD matrix.
Algorithmic bytes per mat-vec = 12 nnz + 8 (n+1) + 16 n (values + column ids once, row pointers, v in, y out)."""
import json
import os
import sys
import numpy as np
import scipy.sparse as sp
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import problems as pr
from helpers import make_solver
from ccqppy_b200 import solution_spaces as ss

out = {}
for n, band, extra in ((1 << 20, 24, 8), (1 << 18, 96, 32), (1 << 16, 8, 0)):
    rng = np.random.default_rng(0)
    offs = np.arange(1, band + 1)
    diags = [rng.standard_normal(n - o) * 0.5 / band for o in offs]
    B = sp.diags(diags, offs, shape=(n, n), format="csr")
    if extra:
        rows = np.repeat(np.arange(n), extra // 2)
        cols = rng.integers(0, n, rows.size)
        B = B + sp.csr_matrix((rng.standard_normal(rows.size) * 0.5 / band, (rows, cols)), shape=(n, n))
    A = (B + B.T + 2.0 * sp.identity(n)).tocsr()          # symmetric, strictly diagonally dominant
    xs = 1.0 - 4.0 * rng.random(n)
    b = -(A @ xs)
    At = torch.sparse_csr_tensor(torch.from_numpy(A.indptr.astype(np.int64)), torch.from_numpy(A.indices.astype(np.int64)),
                                 torch.from_numpy(A.data), size=(n, n)).cuda()
    bt = torch.from_numpy(b).cuda()
    op = ss.BoxProjOp(n)
    row = {}
    for solver in (pr.BBPGD, pr.SPG, pr.MPRGP):
        for _ in range(2):
            s = make_solver(solver, 1e-6, 2000)
            s.solve(At, bt, convex_proj_op=op, uniforms=torch.rand(2000, dtype=torch.float64).cuda())
        row[pr.SOLVER_NAMES[solver]] = dict(mv=int(s.solution_num_matrix_vector_multiplications), gemv=int(s.solution_gemv_count),
                                            converged=bool(s.solution_converged), kernel_ms=1e3 * s.solution_gpu_time,
                                            GBps=s.solution_hbm_bytes / s.solution_gpu_time / 1e9,
                                            us_per_matvec=1e6 * s.solution_gpu_time / s.solution_gemv_count)
    out["n=%d nnz/row=%.1f" % (n, A.nnz / n)] = row
print(json.dumps(out, indent=1))
