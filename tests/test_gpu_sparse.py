"""Operator-form Hessians: the reference takes any `A` with `.dot`, in particular
scipy.sparse matrices; here they run through the CSR mat-vec phase of the same persistent solver kernel.
The oracle is the NumPy restatement fed the SAME scipy matrix (it only calls A.dot, like the reference)."""
import ctypes

import numpy as np
import pytest
import scipy.sparse as sp

import problems as pr
from helpers import op_from_table, make_solver
from oracle import ccqp_oracle as orc

pytestmark = pytest.mark.gpu


def contact_like(n, m_per_row, seed, mu=0.5):
    """A = D^T D / scale + mu I with a sparse random D: symmetric positive definite, ~2 m^2 entries per row."""
    rng = np.random.default_rng(seed)
    D = sp.random(n, n, density=min(1.0, m_per_row / n), random_state=np.random.RandomState(seed), format="csr",
                  data_rvs=lambda k: rng.standard_normal(k))
    A = (D.T @ D).tocsr()
    A = A / max(abs(A).sum(axis=1).max(), 1e-300) * 4.0 + mu * sp.identity(n, format="csr")
    A = A.tocsr()
    xs = 1.0 - 4.0 * rng.random(n)
    return A, -(A @ xs)


# (20000, 1e-4): 2 entries per row and many empty rows -> ~2000 rows per 4096-entry tile, far more than the shared row-pointer
# window holds (the overflow path); (6000, 0.8): 4800 entries per row -> every row spans two tiles (the carry path)
@pytest.mark.parametrize("n,dens", [(1, 1.0), (7, 0.5), (300, 0.02), (1000, 0.2), (4097, 0.004), (20000, 0.001), (20000, 1e-4), (6000, 0.8)])
def test_csr_gemv_matches_scipy(n, dens):
    from ccqppy_b200 import _capi
    rng = np.random.default_rng(n)
    A = sp.random(n, n, density=dens, random_state=np.random.RandomState(n), format="csr", data_rvs=lambda k: rng.standard_normal(k))
    v = rng.standard_normal(n)
    h = _capi.Handle()
    ptr, idx, val = A.indptr.astype(np.int64), A.indices.astype(np.int32), A.data.astype(np.float64)
    P = lambda a: ctypes.c_void_p(a.ctypes.data)
    _capi.check(h.h, h.lib.ccqp_set_matrix_csr(h.h, P(ptr), P(idx), P(val), n, val.size, 0, n, _capi.MEM_HOST))
    y = np.empty(n)
    _capi.check(h.h, h.lib.ccqp_gemv(h.h, P(v), P(y), _capi.MEM_HOST))
    ref = A @ v
    scale = abs(A) @ abs(v) + 1e-300
    # rounding of a reordered sum of m terms: ~ sqrt(m) eps relative to sum |a||v|
    assert np.max(np.abs(y - ref) / scale) < 1e-15 * max(4, np.log2(n + 1), 0.5 * np.sqrt(A.nnz / n))
    # malformed input is refused, not executed
    bad = idx.copy()
    if bad.size:
        bad[0] = n
        assert h.lib.ccqp_set_matrix_csr(h.h, P(ptr), P(bad), P(val), n, val.size, 0, n, _capi.MEM_HOST) == 1
    h.close()


@pytest.mark.parametrize("solver", [pr.PGD, pr.APGD, pr.APGD_AR, pr.BBPGD, pr.BBPGDF, pr.SPG, pr.MPRGP])
@pytest.mark.parametrize("table", ["box", "sphere3", "mixed"])
def test_sparse_solves_match_oracle(solver, table):
    n = 1200
    A, b = contact_like(n, 6, seed=3)
    tab = {"box": pr.box_table, "sphere3": pr.sphere3_table, "mixed": pr.mixed_table}[table](n)
    uni = pr.spg_uniforms(1, 3000)
    o = orc.solve(solver, A, b, blocks=tab.blocks, params=tab.params, tol=1e-6, max_mv=3000, step_size=0.2, uniforms=uni)
    s = make_solver(solver, 1e-6, 3000, 0.2)
    s.solve(A, b, convex_proj_op=op_from_table(tab), uniforms=uni)
    assert s.solution_converged == o["converged"]
    assert abs(s.solution_num_matrix_vector_multiplications - o["mv"]) <= max(1, round(0.02 * o["mv"]))
    if s.solution_num_matrix_vector_multiplications == o["mv"]:
        assert np.linalg.norm(np.asarray(s.solution) - o["solution"]) <= 1e-9 * np.linalg.norm(o["solution"])
    # a dense copy of the same matrix gives the same answer through the dense path
    d = make_solver(solver, 1e-6, 3000, 0.2)
    d.solve(A.toarray(), b, convex_proj_op=op_from_table(tab), uniforms=uni)
    assert d.solution_num_matrix_vector_multiplications == s.solution_num_matrix_vector_multiplications
    np.testing.assert_allclose(np.asarray(d.solution), np.asarray(s.solution), rtol=1e-10, atol=1e-12)


def test_torch_sparse_csr_on_device_and_accounting():
    import torch
    n = 30000
    A, b = contact_like(n, 5, seed=8)
    tab = pr.box_table(n)
    host = make_solver(pr.BBPGD, 1e-7, 2000)
    host.solve(A, b, convex_proj_op=op_from_table(tab))
    At = torch.sparse_csr_tensor(torch.from_numpy(A.indptr.astype(np.int64)), torch.from_numpy(A.indices.astype(np.int64)),
                                 torch.from_numpy(A.data), size=(n, n)).cuda()
    s = make_solver(pr.BBPGD, 1e-7, 2000)
    s.solve(At, torch.from_numpy(b).cuda(), convex_proj_op=op_from_table(tab))
    assert s.solution.is_cuda and s.solution_converged
    assert np.array_equal(s.solution.cpu().numpy(), np.asarray(host.solution))
    assert s.solution_hbm_bytes == s.solution_gemv_count * (12.0 * A.nnz + 8.0 * (n + 1) + 16.0 * n)
    x = np.asarray(host.solution)
    g = A @ x + b
    assert np.linalg.norm(x - np.clip(x - 1e-6 * g, -1, 1)) / (3 * n * 1e-6) < 1e-7 * 1.01
