"""Debug: cycles of the pieces of a BBPGD iteration in the symmetric batched kernel (variant build with CCQP_BSYM_TIMING)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from ccqppy_b200 import solvers
n, batch = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 148
g = torch.Generator(device="cuda").manual_seed(1)
G = torch.randn((batch, n, n), generator=g, device="cuda", dtype=torch.float64)
A = G @ G.transpose(1, 2) / n + torch.eye(n, device="cuda", dtype=torch.float64)
A = 0.5 * (A + A.transpose(1, 2))
xs = 1 - 4 * torch.rand((batch, n), generator=g, device="cuda", dtype=torch.float64)
b = -(A @ xs.unsqueeze(-1)).squeeze(-1)
lb, ub = -torch.ones_like(b), torch.ones_like(b)
s = solvers.CCQPSolverBBPGD(1e-8, 5000); s.quiet = True
for _ in range(2):
    s.solve_batched(A, b, lb, ub, symmetric=True)
r = s._batched_records
iters = r["mv"] // 1000000
print("problems", batch, "iterations mean", iters.mean(), "whole solve, cycles", r["residual"].mean(), "per iteration", r["residual"].mean() / (iters.mean() + 2))
print("per iteration: projection+step %.0f  mat-vec %.0f  products %.0f  reduction %.0f" % (
    (r["mv"] % 1000000).mean(), r["gemv"].mean(), r["draws"].mean(), r["it"].mean()))
pk = r["residual"].astype(np.int64)
print("inside the mat-vec, cumulative cycles: x loads + diagonal %.0f | s1 %.0f | s2 %.0f | slots read + added %.0f" % tuple(
    4 * ((pk >> (12 * k)) & 0xfff).mean() for k in range(4)))
