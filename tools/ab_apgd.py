"""APGD / SPG at n = 32768 (or argv[1]) on device-resident data: kernel ms and GB/s per solver; used for A/B runs with env switches."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from ccqppy_b200 import solvers, solution_spaces as ss
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
g = torch.Generator(device="cuda").manual_seed(0)
G = torch.randn((n, n), generator=g, device="cuda", dtype=torch.float64)
A = G @ G.t() / n; A.diagonal().add_(1.0); del G
b = -(A @ (1 - 4 * torch.rand(n, generator=g, device="cuda", dtype=torch.float64)))
uni = torch.from_numpy(np.random.RandomState(0).random_sample(2000)).cuda()
out = {}
for name, cls in (("SPG", solvers.CCQPSolverSPG), ("APGD", solvers.CCQPSolverAPGD), ("BBPGD", solvers.CCQPSolverBBPGD)):
    best = None
    for _ in range(3):
        s = cls(1e-5, 2000); s.quiet = True
        s.solve(A, b, convex_proj_op=ss.BoxProjOp(n, -1.0, 1.0), uniforms=uni)
        if best is None or s.solution_gpu_time < best[0]:
            best = (s.solution_gpu_time, s.solution_hbm_bytes, s.solution_gemv_count)
    out[name] = dict(ms=round(1e3 * best[0], 3), GBps=round(best[1] / best[0] / 1e9), gemv=best[2])
print(json.dumps(out))
