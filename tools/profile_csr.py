"""CSR mat-vec phase in isolation (the hook kernel, ccqp_gemv_timed) and inside one SPG solve, n = 2^20, ~57 entries per row.
Run plainly, then under ncu (-k regex:dense_kernel).  CCQP_DEBUG_TIMING=1 prints the phase times of an SPG iteration."""
import ctypes as C
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from ccqppy_b200 import _capi, solvers, solution_spaces as ss

n = 1 << 20
A, b = bench.sparse_matrix(n, 24, 8)
dev = torch.device("cuda")
ptr = torch.from_numpy(A.indptr.astype(np.int64)).to(dev)
idx = torch.from_numpy(A.indices.astype(np.int32)).to(dev)
val = torch.from_numpy(A.data).to(dev)
v = torch.zeros(n + 64, dtype=torch.float64, device=dev); v[:n] = torch.randn(n, dtype=torch.float64, device=dev)
y = torch.zeros(n + 64, dtype=torch.float64, device=dev)
torch.cuda.synchronize()
h = _capi.Handle()
P = lambda t: C.c_void_p(t.data_ptr())
_capi.check(h.h, h.lib.ccqp_set_matrix_csr(h.h, P(ptr), P(idx), P(val), n, int(val.numel()), 0, n, _capi.MEM_DEVICE))
sec = C.c_double()
_capi.check(h.h, h.lib.ccqp_gemv_timed(h.h, P(v), P(y), int(os.environ.get("REPS", 3)), C.byref(sec)))
bytes_ = 12.0 * A.nnz + 8.0 * (n + 1) + 16.0 * n
print("csr gemv hook n=%d nnz=%d: %.1f us, %.0f GB/s" % (n, A.nnz, sec.value * 1e6, bytes_ / sec.value / 1e9))
ref = torch.from_numpy(A @ v[:n].cpu().numpy()).to(dev)
print("max rel err vs scipy: %.2e" % float(((y[:n] - ref).abs() / (ref.abs() + 1e-300)).max()))
h.close()
if "--solve" in sys.argv:
    At = torch.sparse_csr_tensor(ptr, idx.to(torch.int64), val, size=(n, n))
    s = solvers.CCQPSolverSPG(1e-6, 2000); s.quiet = True
    s.solve(At, torch.from_numpy(b).to(dev), convex_proj_op=ss.BoxProjOp(n), uniforms=torch.rand(2000, dtype=torch.float64, device=dev))
    print("spg: mv %d, %.2f ms, %.1f us per mat-vec" % (s.solution_gemv_count, 1e3 * s.solution_gpu_time, 1e6 * s.solution_gpu_time / s.solution_gemv_count))
