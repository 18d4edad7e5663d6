"""Phase times of an MPRGP iteration on the n = 2^20 CSR problem (CCQP_DEBUG_TIMING=1)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from ccqppy_b200 import solvers, solution_spaces as ss
n = 1 << 20
A, b = bench.sparse_matrix(n, 24, 8)
dev = torch.device("cuda")
At = torch.sparse_csr_tensor(torch.from_numpy(A.indptr.astype(np.int64)).to(dev), torch.from_numpy(A.indices.astype(np.int64)).to(dev),
                             torch.from_numpy(A.data).to(dev), size=(n, n))
for cls in (solvers.CCQPSolverMPRGP, solvers.CCQPSolverBBPGD):
    s = cls(1e-6, 2000); s.quiet = True
    s.solve(At, torch.from_numpy(b).to(dev), convex_proj_op=ss.BoxProjOp(n))
    print(cls.__name__, "mv", s.solution_num_matrix_vector_multiplications, "gemv", s.solution_gemv_count, "%.2f ms" % (1e3 * s.solution_gpu_time))
