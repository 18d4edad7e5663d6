"""The BASELINE.json configs that are not the bench.py headline (1: README 3x3, 2: n=4096 all solvers,
5: n=16384 MPRGP with friction-style blocks), timed on the GPU next to the NumPy/OpenBLAS port of the
reference where that finishes in seconds.  Prints one JSON object (kept under profiles/).

    python tools/bench_configs.py > gpurun_out/configs.json
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import problems as pr                                  # noqa: E402
from helpers import op_from_table, make_solver         # noqa: E402
from oracle import ccqp_oracle as orc                  # noqa: E402
from test_gpu_fullsize import gpu_problem              # noqa: E402

out = {}


def gpu_time(solver, A, b, op, tol, max_mv, step=0.01, reps=3, uniforms=None):
    best, s = None, None
    for _ in range(reps):
        s = make_solver(solver, tol, max_mv, step)
        t0 = time.perf_counter()
        s.solve(A, b, convex_proj_op=op, uniforms=uniforms)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        best = wall if best is None else min(best, wall)
    n = b.shape[0]
    return dict(mv=int(s.solution_num_matrix_vector_multiplications), gemv=int(s.solution_gemv_count),
                converged=bool(s.solution_converged), residual=float(s.solution_residual),
                kernel_ms=1e3 * s.solution_gpu_time, wall_ms=1e3 * best,
                GBps=s.solution_hbm_bytes / s.solution_gpu_time / 1e9 if n >= 1024 else None), s


# ---- config 1: README example (host NumPy int arrays through the public API)
A, b = pr.tridiag_problem()
tab = pr.Table().add(pr.BOX, 3, np.array([-2., -2., -4.]), np.array([2., 2., 5.]))
np.random.seed(0)
g, s = gpu_time(pr.SPG, A, b, op_from_table(tab), 1e-10, 5000, reps=5, uniforms=pr.spg_uniforms(0, 5000))
t0 = time.perf_counter()
o = orc.solve(pr.SPG, A, b, blocks=tab.blocks, params=tab.params, tol=1e-10, max_mv=5000, uniforms=pr.spg_uniforms(0, 5000))
out["config1_readme_3x3_spg"] = dict(gpu=g, cpu_port_ms=1e3 * (time.perf_counter() - t0), cpu_mv=o["mv"],
                                     solution=np.asarray(s.solution).tolist(), note="launch-latency bound; README printed 5.9 ms / 86 mv")

# ---- config 2: n = 4096, box, every solver, device-resident A vs the port on the host cores
n = 4096
A, b = pr.shift_problem(n, 0)
tabb = pr.box_table(n)
Ad, bd = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
step = 1.0 / np.abs(A).sum(axis=1).max()
c2 = {}
for solver in range(7):
    uni = pr.spg_uniforms(0, 2000)
    g, s = gpu_time(solver, Ad, bd, op_from_table(tabb), 1e-5, 2000, step, uniforms=torch.from_numpy(uni).cuda())
    t0 = time.perf_counter()
    o = orc.solve(solver, A, b, blocks=tabb.blocks, params=tabb.params, tol=1e-5, max_mv=2000, step_size=step, uniforms=uni)
    cpu = time.perf_counter() - t0
    err = float(np.linalg.norm(s.solution.cpu().numpy() - o["solution"]) / np.linalg.norm(o["solution"]))
    c2[pr.SOLVER_NAMES[solver]] = dict(gpu=g, cpu_port_ms=1e3 * cpu, cpu_mv=o["mv"], rel_err_vs_port=err,
                                       speedup_kernel=cpu / (g["kernel_ms"] * 1e-3))
out["config2_n4096_box"] = c2
del Ad, bd

# ---- config 5: n = 16384, MPRGP (and BBPGD/SPG for comparison), contact-style blocks
n = 16384
A, b = gpu_problem(n, seed=3)
c5 = {}
for name, tab in (("sphere3_discs", pr.sphere3_table(n)), ("mixed_box_lower_upper_sphere3", pr.mixed_table(n)),
                  ("soc3_cones_mu0.5(extension)", pr.soc3_table(n, 0.5)), ("box", pr.box_table(n))):
    op = op_from_table(tab)
    row = {}
    for solver in (pr.MPRGP, pr.BBPGD, pr.SPG):
        g, s = gpu_time(solver, A, b, op, 1e-5, 3000, reps=2, uniforms=torch.from_numpy(pr.spg_uniforms(0, 3000)).cuda())
        row[pr.SOLVER_NAMES[solver]] = g
    c5[name] = row
out["config5_n16384"] = c5
# the port of the reference on a size it finishes: MPRGP spends its time in the per-element Python loops of normal_vector
n_small = 1536
As, bs = pr.shift_problem(n_small, 3)
tabs = pr.sphere3_table(n_small)
t0 = time.perf_counter()
o = orc.solve(pr.MPRGP, As, bs, blocks=tabs.blocks, params=tabs.params, tol=1e-5, max_mv=3000)
cpu = time.perf_counter() - t0
g, s = gpu_time(pr.MPRGP, As, bs, op_from_table(tabs), 1e-5, 3000)
out["config5_port_vs_gpu_n1536_sphere3_mprgp"] = dict(gpu=g, cpu_port_ms=1e3 * cpu, cpu_mv=o["mv"],
    rel_err_vs_port=float(np.linalg.norm(np.asarray(s.solution) - o["solution"]) / np.linalg.norm(o["solution"])))
print(json.dumps(out, indent=1))
