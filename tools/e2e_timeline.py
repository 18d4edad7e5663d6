"""Where the time of a pipelined stream of host-resident solves goes: wall-clock of every C-ABI call of SolvePipeline.submit /
results at the bench size (debug tool; prints one line per call)."""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import bench
from ccqppy_b200 import _capi, solvers, solution_spaces as ss
from ccqppy_b200.pipeline import SolvePipeline

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
A, b = bench.make_dense_problem(n, torch.device("cuda", 0))
A_host = torch.empty((n, n), dtype=torch.float64, pin_memory=True); A_host.copy_(A); b_host = b.cpu().pin_memory()
del A; torch.cuda.empty_cache()
uni = torch.from_numpy(bench.spg_uniform_stream(bench.MAX_MV)).pin_memory()
op = ss.BoxProjOp(n, -np.ones(n), np.ones(n))
lib = _capi.load()
T0 = time.perf_counter()
log = []


class Timed:
    def __init__(self, lib, names):
        self._lib = lib
        self._names = names

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if name not in self._names:
            return fn

        def wrapped(*a):
            t = time.perf_counter()
            r = fn(*a)
            log.append((name, 1e3 * (t - T0), 1e3 * (time.perf_counter() - t)))
            return r
        return wrapped


# raw upload rates first: the copy alone, nothing else on the GPU
import ctypes
h = _capi.Handle(0)
pa, mem, _ = _capi.f64_ptr(A_host)
for label, call in (("full (CCQP_SYM_UPLOAD=0)", "full"), ("upper block triangle, declared", "decl"), ("upper block triangle + host test", "auto")):
    for rep in range(2):
        os.environ["CCQP_SYM_UPLOAD"] = "0" if call == "full" else "1"
        torch.cuda.synchronize()
        t = time.perf_counter()
        if call == "decl":
            _capi.check(h.h, lib.ccqp_set_matrix_symmetric(h.h, pa, n, n, mem))
        else:
            _capi.check(h.h, lib.ccqp_set_matrix(h.h, pa, n, n, 0, n, mem))
        t_call = time.perf_counter() - t
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
    nb = h.upload_info()[0]
    print("upload alone: %-36s call returns after %6.1f ms, done after %6.1f ms, %.2f GB -> %.1f GB/s" % (label, 1e3 * t_call, 1e3 * dt, nb / 1e9, nb / dt / 1e9))
os.environ["CCQP_SYM_UPLOAD"] = "1"
h.close()
mode = sys.argv[2] if len(sys.argv) > 2 else "auto"
sym = mode == "declared"
names = ("ccqp_set_matrix", "ccqp_set_matrix_symmetric", "ccqp_set_projection", "ccqp_solve_async", "ccqp_solve_wait")
pipe = SolvePipeline(solvers.CCQPSolverSPG(bench.TOL, bench.MAX_MV), depth=2, device=0)
for s in pipe.slots:
    s.handle.lib = Timed(lib, names)
for _ in range(2):
    pipe.submit(A_host, b_host, convex_proj_op=op, uniforms=uni, symmetric=sym)
pipe.results()
torch.cuda.synchronize()
log.clear()
T0 = time.perf_counter()
for _ in range(5):
    pipe.submit(A_host, b_host, convex_proj_op=op, uniforms=uni, symmetric=sym)
res = pipe.results()
total = time.perf_counter() - T0
for name, at, dur in log:
    print("%8.1f ms  %-22s %7.1f ms" % (at, name, dur))
print("5 solves in %.1f ms; kernel times" % (1e3 * total), [round(1e3 * r.solution_gpu_time, 1) for r in res])
pipe.close()
