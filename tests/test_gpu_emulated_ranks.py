"""The multi-GPU exchange protocol on ONE GPU (runs on the driver's single-GPU test box).

`world` emulated ranks share one cooperative launch of the SAME kernel body a real sharded solve runs
(dense.cuh dense_kernel_emu): every rank streams only its row shard, writes its rows of each mat-vec result into
every peer's buffer from the mat-vec epilogue, and the closing sync exchanges the scalar partial sums as
{data, epoch} packets.  Checked against the plain single-GPU kernel, against the oracle, and rank against rank
(identical bits on every rank: the invariant the in-kernel control flow relies on)."""
import numpy as np
import pytest

import problems as pr
from emulated import EmulatedBox
from helpers import op_from_table, make_solver
from oracle import ccqp_oracle as orc

pytestmark = pytest.mark.gpu

CASES = [("tiny", pr.box_table, 40, 1.0),            # one CTA per rank
         ("mixed", pr.mixed_table, 1200, 1.0),
         ("sphere3", pr.sphere3_table, 1000, 1.0),
         ("box_uneven", pr.box_table, 3050, 1.0),    # 3050 rows over 4 / 8 ranks: shards of unequal size
         ("box_odd", pr.box_table, 1023, 0.3)]


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_emulated_ranks_match_single_gpu_and_oracle(case, world):
    name, table, n, mu = case
    A, b = pr.shift_problem(n, 5, mu)
    tab = table(n)
    op = op_from_table(tab)
    step = 1.0 / np.abs(A).sum(axis=1).max()
    x0 = 2.0 * np.random.default_rng(8).standard_normal(n) if name == "mixed" else None
    box = EmulatedBox(A, op, world)
    assert sum(r1 - r0 for r0, r1 in box.ranges) == n
    tol, max_mv = 1e-6, 1500
    for solver in range(7):
        uni = pr.spg_uniforms(2, max_mv)
        one = make_solver(solver, tol, max_mv, step)
        one.solve(A, b, x0=x0, convex_proj_op=op, uniforms=uni)
        ranks = box.solve(make_solver(solver, tol, max_mv, step), b, x0=x0, uniforms=uni)
        again = box.solve(make_solver(solver, tol, max_mv, step), b, x0=x0, uniforms=uni)     # buffers are reusable
        x1 = np.asarray(one.solution)
        for r in ranks[1:]:                                        # every rank: identical bits
            assert np.array_equal(r["solution"], ranks[0]["solution"]) and r["mv"] == ranks[0]["mv"]
            assert r["residual"] == ranks[0]["residual"] or (np.isnan(r["residual"]) and np.isnan(ranks[0]["residual"]))
        assert np.array_equal(again[0]["solution"], ranks[0]["solution"]) and again[0]["mv"] == ranks[0]["mv"]
        assert ranks[0]["converged"] == bool(one.solution_converged), (name, solver)
        assert ranks[0]["mv"] == one.solution_num_matrix_vector_multiplications, (name, solver)
        assert np.linalg.norm(ranks[0]["solution"] - x1) <= 1e-9 * max(np.linalg.norm(x1), 1e-300), (name, solver)
        if solver in (pr.SPG, pr.BBPGD, pr.MPRGP) and name in ("mixed", "sphere3", "box_uneven"):
            o = orc.solve(solver, A, b, x0=x0, blocks=tab.blocks, params=tab.params, tol=tol, max_mv=max_mv, step_size=step, uniforms=uni)
            assert o["mv"] == ranks[0]["mv"] and o["converged"] == ranks[0]["converged"]
            assert np.linalg.norm(ranks[0]["solution"] - o["solution"]) <= 1e-9 * np.linalg.norm(o["solution"])
    box.close()


@pytest.mark.parametrize("world", [2, 8])
def test_emulated_ranks_csr(world):
    from test_gpu_sparse import contact_like
    n = 3000
    A, b = contact_like(n, 6, seed=12)
    tab = pr.mixed_table(n)
    op = op_from_table(tab)
    box = EmulatedBox(A, op, world)
    for solver in (pr.BBPGD, pr.SPG, pr.MPRGP):
        uni = pr.spg_uniforms(4, 2000)
        one = make_solver(solver, 1e-7, 2000)
        one.solve(A, b, convex_proj_op=op, uniforms=uni)
        ranks = box.solve(make_solver(solver, 1e-7, 2000), b, uniforms=uni)
        assert all(np.array_equal(r["solution"], ranks[0]["solution"]) for r in ranks)
        assert ranks[0]["mv"] == one.solution_num_matrix_vector_multiplications
        x1 = np.asarray(one.solution)
        assert np.linalg.norm(ranks[0]["solution"] - x1) <= 1e-9 * np.linalg.norm(x1)
    box.close()


def test_emulated_box_rejects_plain_solve():
    """Handles of an emulated box cannot be used for ordinary solves (their work buffers are wired to each other)."""
    import ctypes
    from ccqppy_b200 import _capi
    n = 64
    A, b = pr.shift_problem(n, 0)
    box = EmulatedBox(A, op_from_table(pr.box_table(n)), 2)
    prm, res, x = make_solver(pr.BBPGD, 1e-6, 100)._params(), _capi.Result(), np.empty(n)
    P = lambda a: ctypes.c_void_p(a.ctypes.data)
    assert box.lib.ccqp_solve(box.handles[0].h, _capi.BBPGD, ctypes.byref(prm), P(b), None, None, 0, P(x), _capi.MEM_HOST,
                              ctypes.byref(res)) == 4
    box.close()
