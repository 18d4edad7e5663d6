"""Short, fixed program for ncu (run it plainly first, then under ncu; see profiles/README.md).

  part 1: three launches of the mat-vec hook kernel at n=32768 (the hot phase in isolation)
  part 2: one whole SPG solve at n=16384 (A = 2.1 GB >> L2), the persistent solver kernel
  part 3: batched BBPGD and SPG, 16384 problems of n=64 (`batched_sym`: through ccqp_solve_batched_sym)
"""
import ctypes as C
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ccqppy_b200 import _capi, solvers, solution_spaces as ss

which = sys.argv[1:] or ["gemv", "spg", "batched"]
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
if "gemv" in which:
    n = 32768
    A = torch.randn((n, n), generator=g, device=dev, dtype=torch.float64)
    v = torch.zeros(n + 64, device=dev, dtype=torch.float64); v[:n] = torch.randn(n, generator=g, device=dev, dtype=torch.float64)
    y = torch.zeros(n + 64, device=dev, dtype=torch.float64)
    torch.cuda.synchronize()          # a raw handle runs on its own stream, not on torch's
    h = _capi.Handle()
    _capi.check(h.h, h.lib.ccqp_set_matrix(h.h, C.c_void_p(A.data_ptr()), n, n, 0, n, 1))
    sec = C.c_double()
    _capi.check(h.h, h.lib.ccqp_gemv_timed(h.h, C.c_void_p(v.data_ptr()), C.c_void_p(y.data_ptr()), 2, C.byref(sec)))
    print("gemv n=%d: %.1f us, %.0f GB/s" % (n, sec.value * 1e6, (8.0 * n * n + 16 * n) / sec.value / 1e9))
    h.close(); del A
if "spg" in which:
    n = 16384
    G = torch.randn((n, n), generator=g, device=dev, dtype=torch.float64)
    A = G @ G.t() / n; A.diagonal().add_(1.0); del G
    b = -(A @ (1 - 4 * torch.rand(n, generator=g, device=dev, dtype=torch.float64)))
    s = solvers.CCQPSolverSPG(1e-5, 2000); s.quiet = True
    uni = torch.from_numpy(np.random.RandomState(0).random_sample(2000)).to(dev)
    s.solve(A, b, convex_proj_op=ss.BoxProjOp(n), uniforms=uni)
    print("spg n=%d: mv %d, %.2f ms, %.0f GB/s" % (n, s.solution_gemv_count, 1e3 * s.solution_gpu_time, s.solution_hbm_bytes / s.solution_gpu_time / 1e9))
    del A
if "batched" in which or "batched_sym" in which:
    B, nb = 16384, 64
    G = torch.randn((B, nb, nb), generator=g, device=dev, dtype=torch.float64)
    A = G @ G.transpose(1, 2) / nb + torch.eye(nb, device=dev, dtype=torch.float64)
    A = 0.5 * (A + A.transpose(1, 2))
    b = -(A @ (1 - 4 * torch.rand((B, nb, 1), generator=g, device=dev, dtype=torch.float64))).squeeze(-1)
    lb, ub = -torch.ones_like(b), torch.ones_like(b)
    uni = torch.rand((B, 256), generator=g, device=dev, dtype=torch.float64)
    for cls in (solvers.CCQPSolverBBPGD, solvers.CCQPSolverSPG):
        s = cls(1e-8, 5000); s.quiet = True
        s.solve_batched(A, b, lb, ub, uniforms=uni, symmetric="batched_sym" in which)
        print("batched %s: %.3f ms, %.1f M QP/s" % (s.name, 1e3 * s.solution_gpu_time, B / s.solution_gpu_time / 1e6))
