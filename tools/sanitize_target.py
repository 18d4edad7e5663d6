"""Tiny pass over every kernel family for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import problems as pr
from helpers import op_from_table, make_solver
import scipy.sparse as sp

n = 300
A, b = pr.shift_problem(n, 0)
for tab in (pr.mixed_table(n), pr.box_table(n), pr.sphere_table(n)):
    op = op_from_table(tab)
    for solver in range(7):
        s = make_solver(solver, 1e-6, 300, 0.1)
        s.solve(A, b, convex_proj_op=op, uniforms=pr.spg_uniforms(0, 300))
    v = np.random.default_rng(0).standard_normal(n)
    op(v)
    if tab is not None:
        op.normal_vector(np.asarray(op(v)))
print("dense ok", flush=True)
As = (sp.random(500, 500, density=0.02, random_state=np.random.RandomState(1), format="csr") + 2 * sp.identity(500)).tocsr()
As = (As + As.T).tocsr()
bs = -(As @ np.ones(500))
for solver in (pr.BBPGD, pr.SPG, pr.MPRGP):
    make_solver(solver, 1e-6, 300).solve(As, bs, convex_proj_op=op_from_table(pr.box_table(500)), uniforms=pr.spg_uniforms(0, 300))
print("csr ok", flush=True)
for nb in (64, 17):
    B = 40
    Ab = np.empty((B, nb, nb)); bb = np.empty((B, nb))
    for i in range(B):
        Ab[i], bb[i] = pr.shift_problem(nb, i)
    for solver in range(7):
        make_solver(solver, 1e-7, 2000, 0.1).solve_batched(Ab, bb, -np.ones((B, nb)), np.ones((B, nb)))
print("batched ok", flush=True)
for nb in (64, 62, 17):      # the one-warp kernels for symmetric Hessians (staged and direct fills)
    B = 40
    Ab = np.empty((B, nb, nb)); bb = np.empty((B, nb))
    for i in range(B):
        Ab[i], bb[i] = pr.shift_problem(nb, i)
    Ab = 0.5 * (Ab + Ab.transpose(0, 2, 1))
    for solver in (pr.PGD, pr.BBPGD, pr.BBPGDF, pr.SPG):
        make_solver(solver, 1e-7, 2000, 0.1).solve_batched(Ab, bb, -np.ones((B, nb)), np.ones((B, nb)), symmetric=True)
print("batched symmetric ok", flush=True)
for n2, declared in ((2100, False), (1500, True)):      # symmetric upload: upper block triangle + mirror kernel
    A2, b2 = pr.shift_problem(n2, 3)
    A2 = 0.5 * (A2 + A2.T)
    make_solver(pr.BBPGD, 1e-6, 50).solve(A2, b2, convex_proj_op=op_from_table(pr.box_table(n2)), symmetric=declared)
print("upload ok", flush=True)
