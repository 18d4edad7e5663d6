"""Operator-form (CSR) Hessian: solver throughput on banded-plus-random sparse SPD matrices (bench.bench_sparse,
all three matrices and BBPGD / SPG / MPRGP).  Algorithmic bytes per mat-vec = 12 nnz + 8 (n+1) + 16 n."""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
print(json.dumps(bench.bench_sparse(torch.device("cuda", 0), cases=((1 << 20, 24, 8), (1 << 18, 96, 32), (1 << 16, 8, 0)),
                                    solvers_=("BBPGD", "SPG", "MPRGP")), indent=1))
