"""Run under torchrun (one rank per GPU): row-sharded solves against the single-GPU path and the
oracle.  Every rank checks its own copy of the results; rank 0 prints a JSON summary.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import problems as pr                                   # noqa: E402
from helpers import op_from_table, make_solver          # noqa: E402
from ccqppy_b200.dist import ShardedSolver, solve_batched_sharded, batch_range   # noqa: E402
from oracle import ccqp_oracle as orc                   # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    report, ok = [], True
    cases = [("tiny", pr.box_table(40), 40, 1.0),      # one CTA per rank: local syncs are plain barriers, cross syncs are not
             ("mixed", pr.mixed_table(1200), 1200, 1.0), ("box", pr.box_table(4096), 4096, 1.0),
             ("sphere3", pr.sphere3_table(1000), 1000, 1.0), ("box_odd", pr.box_table(1023), 1023, 0.3),
             ("box_uneven", pr.box_table(3050), 3050, 1.0)]   # 3050 rows over 8 ranks: shards of 381 / 382 rows, one grid
    for tname, tab, n, mu in cases:
        A, b = pr.shift_problem(n, 5, mu)
        step = 1.0 / np.abs(A).sum(axis=1).max()
        x0 = 2.0 * np.random.default_rng(8).standard_normal(n) if tname == "mixed" else None
        op = op_from_table(tab)
        Ad = torch.from_numpy(A).to(dev)
        for solver in range(7):
            tol, max_mv = 1e-6, 1500
            uni = pr.spg_uniforms(2, max_mv)
            one = make_solver(solver, tol, max_mv, step)
            one.solve(Ad, torch.from_numpy(b).to(dev), x0=None if x0 is None else torch.from_numpy(x0).to(dev),
                      convex_proj_op=op, uniforms=uni)
            runner = ShardedSolver(make_solver(solver, tol, max_mv, step), Ad, op, rank, world, dev)
            r = runner.solve(b, x0=x0, uniforms=uni)
            r2 = runner.solve(b, x0=x0, uniforms=uni)            # the exchange buffers are reusable
            xs, x1 = r.solution.cpu().numpy(), one.solution.cpu().numpy()
            err = float(np.linalg.norm(xs - x1) / max(np.linalg.norm(x1), 1e-300))
            same = (r.solution_num_matrix_vector_multiplications == one.solution_num_matrix_vector_multiplications
                    and r.solution_converged == one.solution_converged)
            rerun = bool(torch.equal(r.solution, r2.solution)) and \
                r.solution_num_matrix_vector_multiplications == r2.solution_num_matrix_vector_multiplications
            # every rank must hold the identical full solution and identical scalars
            gathered = [torch.empty_like(r.solution) for _ in range(world)]
            dist.all_gather(gathered, r.solution)
            ident = all(bool(torch.equal(g, gathered[0])) for g in gathered)
            mvs = [None] * world
            dist.all_gather_object(mvs, (r.solution_num_matrix_vector_multiplications, r.solution_residual))
            ident = ident and all(m == mvs[0] for m in mvs)
            good = same and err < 1e-9 and rerun and ident
            if rank == 0 and solver in (pr.SPG, pr.BBPGD, pr.MPRGP) and tname in ("mixed", "sphere3"):
                o = orc.solve(solver, A, b, x0=x0, blocks=tab.blocks, params=tab.params, tol=tol, max_mv=max_mv,
                              step_size=step, uniforms=uni)
                oerr = float(np.linalg.norm(xs - o["solution"]) / np.linalg.norm(o["solution"]))
                good = good and o["mv"] == r.solution_num_matrix_vector_multiplications and oerr < 1e-9
            ok = ok and good
            report.append(dict(table=tname, solver=pr.SOLVER_NAMES[solver], mv=r.solution_num_matrix_vector_multiplications,
                               mv_single=one.solution_num_matrix_vector_multiplications, err=err, rerun_identical=rerun,
                               ranks_identical=ident, ok=good, gpu_ms=1e3 * r.solution_gpu_time,
                               single_ms=1e3 * one.solution_gpu_time))
            runner.close()
    # ---- re-upload of a row shard from (pinned) host memory: same answers as the resident shard
    n = 2048
    A, b = pr.shift_problem(n, 6)
    tab = pr.box_table(n)
    op = op_from_table(tab)
    runner = ShardedSolver(make_solver(pr.BBPGD, 1e-7, 500), torch.from_numpy(A).to(dev), op, rank, world, dev)
    r_dev = runner.solve(b)
    r0, r1 = runner.ranges[rank]
    runner.set_matrix(torch.from_numpy(A[r0:r1].copy()).pin_memory())
    r_host = runner.solve(b)
    runner.set_matrix(np.ascontiguousarray(2.0 * A[r0:r1]))           # a different Hessian must give a different answer
    r_other = runner.solve(b)
    good = bool(torch.equal(r_dev.solution, r_host.solution)) and not bool(torch.equal(r_dev.solution, r_other.solution))
    o = orc.solve(pr.BBPGD, 2.0 * A, b, blocks=tab.blocks, params=tab.params, tol=1e-7, max_mv=500)
    good = good and float(np.linalg.norm(r_other.solution.cpu().numpy() - o["solution"]) / np.linalg.norm(o["solution"])) < 1e-9
    ok = ok and good
    report.append(dict(table="box", solver="set_matrix", mv=r_host.solution_num_matrix_vector_multiplications,
                       mv_single=r_dev.solution_num_matrix_vector_multiplications, err=0.0, ok=good, gpu_ms=0.0, single_ms=0.0))
    runner.close()

    # ---- operator-form (CSR) Hessian, rows of the CSR arrays sharded over the ranks
    from test_gpu_sparse import contact_like            # well conditioned: the counts do not depend on the summation order
    n = 3000
    As, bs = contact_like(n, 6, seed=12)
    tab = pr.mixed_table(n)
    op = op_from_table(tab)
    for solver in (pr.BBPGD, pr.SPG, pr.MPRGP):
        uni = pr.spg_uniforms(4, 2000)
        one = make_solver(solver, 1e-7, 2000)
        one.solve(As, bs, convex_proj_op=op, uniforms=uni)
        runner = ShardedSolver(make_solver(solver, 1e-7, 2000), As, op, rank, world, dev)
        r = runner.solve(bs, uniforms=uni)
        err = float(np.linalg.norm(r.solution.cpu().numpy() - np.asarray(one.solution)) / np.linalg.norm(np.asarray(one.solution)))
        good = r.solution_num_matrix_vector_multiplications == one.solution_num_matrix_vector_multiplications and err < 1e-9
        ok = ok and good
        report.append(dict(table="csr", solver=pr.SOLVER_NAMES[solver], mv=r.solution_num_matrix_vector_multiplications,
                           mv_single=one.solution_num_matrix_vector_multiplications, err=err, ok=good,
                           gpu_ms=1e3 * r.solution_gpu_time, single_ms=1e3 * one.solution_gpu_time))
        runner.close()

    # ---- batched mode split over the ranks (no communication on the data path) against one GPU
    batch, nb = 37, 64
    Ab = np.empty((batch, nb, nb)); bb = np.empty((batch, nb))
    for i in range(batch):
        Ab[i], bb[i] = pr.shift_problem(nb, 50 + i)
    lb, ub = -np.ones((batch, nb)), np.ones((batch, nb))
    for solver in (pr.BBPGD, pr.SPG):
        one = make_solver(solver, 1e-8, 5000)
        one.solve_batched(Ab, bb, lb, ub, seeds=np.arange(batch))
        x, res, conv, mv, (i0, i1) = solve_batched_sharded(make_solver(solver, 1e-8, 5000), Ab, bb, lb, ub, seeds=np.arange(batch))
        good = (i0, i1) == batch_range(batch, rank, world) and np.array_equal(x.cpu().numpy(), np.asarray(one.solution)) \
            and np.array_equal(mv.cpu().numpy(), one.solution_num_matrix_vector_multiplications) and bool(conv.all())
        ok = ok and good
        report.append(dict(table="batched", solver=pr.SOLVER_NAMES[solver], mv=int(mv.sum()), mv_single=int(one.solution_num_matrix_vector_multiplications.sum()),
                           err=0.0, ok=good, gpu_ms=0.0, single_ms=0.0))

    flags = [None] * world
    dist.all_gather_object(flags, ok)
    if rank == 0:
        out = dict(world=world, all_ok=all(flags), cases=report)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", "multi_gpu_check_w%d.json" % world), "w"), indent=0)
        for c in report:
            print("%-8s %-8s mv %4d / %4d  err %.1e  %s  sharded %.2f ms single %.2f ms" %
                  (c["table"], c["solver"], c["mv"], c["mv_single"], c["err"], "ok" if c["ok"] else "FAIL",
                   c["gpu_ms"], c["single_ms"]))
        print("MULTI_GPU_CHECK", "PASS" if all(flags) else "FAIL", "world", world)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if all(flags) else 1)


if __name__ == "__main__":
    main()
