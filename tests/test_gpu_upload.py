"""Symmetric upload (csrc/upload.cu): a symmetric host matrix crosses PCIe as its upper block triangle and is mirrored on the
device; the device copy -- hence every result -- must be bit-identical to a full upload, for symmetric and for general A."""
import ctypes
import os

import numpy as np
import pytest

import problems as pr
from helpers import make_solver, op_from_table
from ccqppy_b200 import _capi

pytestmark = pytest.mark.gpu

# the host-side symmetry test (and with it the automatic half upload) runs only where the process has at least 12 hardware
# threads (csrc/upload.cu); on a smaller host ccqp_set_matrix uploads every matrix whole
AUTO = len(os.sched_getaffinity(0)) >= 12


def upload_info():
    h = _capi.default_handle(-1)
    nbytes, mirrored = ctypes.c_int64(), ctypes.c_int32()
    _capi.check(h.h, h.lib.ccqp_get_upload_info(h.h, ctypes.byref(nbytes), ctypes.byref(mirrored)))
    return nbytes.value, bool(mirrored.value)


def solve_host(solver, A, b, tab, scheme_on):
    old = os.environ.get("CCQP_SYM_UPLOAD")
    os.environ["CCQP_SYM_UPLOAD"] = "1" if scheme_on else "0"
    try:
        s = make_solver(solver, 1e-6, 400)
        np.random.seed(0)
        s.solve(A, b, convex_proj_op=op_from_table(tab))
        return np.array(s.solution), int(s.solution_num_matrix_vector_multiplications), upload_info()
    finally:
        if old is None:
            del os.environ["CCQP_SYM_UPLOAD"]
        else:
            os.environ["CCQP_SYM_UPLOAD"] = old


@pytest.mark.parametrize("n", [2048, 3000, 4100])
def test_symmetric_matrix_uploads_upper_block_triangle(n):
    B = _capi.load().ccqp_upload_block_rows()
    A, b = pr.shift_problem(n, 11)
    A = 0.5 * (A + A.T)
    tab = pr.box_table(n)
    for solver in (pr.BBPGD, pr.SPG):
        x1, mv1, (bytes1, mir1) = solve_host(solver, A, b, tab, True)
        x0, mv0, (bytes0, mir0) = solve_host(solver, A, b, tab, False)
        assert mir1 == AUTO and not mir0
        assert bytes0 == 8 * n * n
        if AUTO:
            assert bytes1 == 8 * sum(min(B, n - i0) * (n - i0) for i0 in range(0, n, B)) < 0.8 * bytes0
        else:
            assert bytes1 == bytes0
        assert mv1 == mv0 and np.array_equal(x1, x0)          # the device copies are bit-identical


def test_general_matrix_is_uploaded_whole():
    """Not symmetric (one entry below the block diagonal differs, or the matrix is plainly unsymmetric): all of it is
    uploaded, and the solver works on exactly the matrix that was passed."""
    n = 2560
    B = _capi.load().ccqp_upload_block_rows()
    A, b = pr.shift_problem(n, 5)
    A = 0.5 * (A + A.T)
    tab = pr.box_table(n)
    C = A.copy()
    C[B + 17, 3] += 0.25
    x1, mv1, (bytes1, mir1) = solve_host(pr.BBPGD, C, b, tab, True)
    x0, mv0, (bytes0, mir0) = solve_host(pr.BBPGD, C, b, tab, False)
    assert not mir1 and bytes1 == 8 * n * n
    assert mv1 == mv0 and np.array_equal(x1, x0)
    xs, _, _ = solve_host(pr.BBPGD, A, b, tab, True)
    assert not np.array_equal(xs, x1)                          # ... and that entry matters


def test_gemv_after_each_upload_path():
    """y = A v through the unit-test hook: the mirrored copy of a symmetric matrix and the full copy of an unsymmetric one."""
    import torch
    n = 2304
    rng = np.random.default_rng(3)
    G = rng.standard_normal((n, n))
    v = rng.standard_normal(n)
    for A in (G + G.T, G):
        h = _capi.Handle(-1)
        pa, mem, _ = _capi.f64_ptr(A)
        _capi.check(h.h, h.lib.ccqp_set_matrix(h.h, pa, n, n, 0, n, mem))
        y = np.empty(n)
        _capi.check(h.h, h.lib.ccqp_gemv(h.h, _capi.f64_ptr(v)[0], _capi.f64_ptr(y)[0], _capi.MEM_HOST))
        ref = A @ v
        assert np.abs(y - ref).max() <= 1e-11 * np.abs(ref).max()
        h.close()
    assert torch.cuda.is_available()


@pytest.mark.parametrize("n", [2600, 1500])
def test_declared_symmetric_reads_upper_block_triangle_only(n):
    """solve(..., symmetric=True) -> ccqp_set_matrix_symmetric: no test, the blocks below the block diagonal of the host
    matrix are never read (NaN-poisoned they change nothing) and the result is bit-identical to the plain solve of the
    symmetric matrix; SolvePipeline.submit(symmetric=True) goes the same way."""
    from ccqppy_b200.pipeline import SolvePipeline
    B = _capi.load().ccqp_upload_block_rows()
    A, b = pr.shift_problem(n, 21)
    A = 0.5 * (A + A.T)
    op = op_from_table(pr.box_table(n))
    ref = make_solver(pr.BBPGD, 1e-6, 400).solve(A, b, convex_proj_op=op)
    poisoned = A.copy()
    rr, cc = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    poisoned[cc < B * (rr // B)] = np.nan
    s = make_solver(pr.BBPGD, 1e-6, 400).solve(poisoned, b, convex_proj_op=op, symmetric=True)
    nbytes, mirrored = upload_info()
    assert mirrored and nbytes < 0.8 * 8 * n * n
    assert np.array_equal(np.asarray(s.solution), np.asarray(ref.solution))
    assert s.solution_num_matrix_vector_multiplications == ref.solution_num_matrix_vector_multiplications
    pipe = SolvePipeline(make_solver(pr.BBPGD, 1e-6, 400), depth=2)
    for _ in range(3):
        pipe.submit(poisoned, b, convex_proj_op=op, symmetric=True)
    out = pipe.results()
    pipe.close()
    assert all(np.array_equal(np.asarray(r.solution), np.asarray(ref.solution)) for r in out)
    # a device-resident matrix is used in place: the declaration changes nothing
    import torch
    d = make_solver(pr.BBPGD, 1e-6, 400).solve(torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda(), convex_proj_op=op, symmetric=True)
    assert np.array_equal(d.solution.cpu().numpy(), np.asarray(ref.solution))


def test_pending_mirror_does_not_outlive_its_matrix():
    """The mirror kernel of a symmetric host matrix is launched by the first user of the matrix; a matrix that is replaced
    before it was ever used (here by a CSR Hessian of another size) must not be mirrored into the new one's buffers."""
    import scipy.sparse as sp
    n = 2304
    A, b = pr.shift_problem(n, 2)
    A = 0.5 * (A + A.T)
    h = _capi.Handle(-1)
    pa, mem, _ = _capi.f64_ptr(A)
    _capi.check(h.h, h.lib.ccqp_set_matrix_symmetric(h.h, pa, n, n, mem))      # upload enqueued, mirror pending
    assert h.upload_info()[1]
    m = 3000
    S = (sp.random(m, m, density=0.002, random_state=1, format="csr") + sp.identity(m) * 4.0).tocsr()
    S = (S + S.T).tocsr()
    S.sort_indices()
    indptr = S.indptr.astype(np.int64); indices = S.indices.astype(np.int32); values = S.data.astype(np.float64)
    _capi.check(h.h, h.lib.ccqp_set_matrix_csr(h.h, ctypes.c_void_p(indptr.ctypes.data), ctypes.c_void_p(indices.ctypes.data),
                                               ctypes.c_void_p(values.ctypes.data), m, S.nnz, 0, m, _capi.MEM_HOST))
    assert h.upload_info() == (0, False)
    v = np.random.default_rng(0).standard_normal(m)
    y = np.empty(m)
    _capi.check(h.h, h.lib.ccqp_gemv(h.h, _capi.f64_ptr(v)[0], _capi.f64_ptr(y)[0], _capi.MEM_HOST))
    ref = S @ v
    assert np.abs(y - ref).max() <= 1e-12 * np.abs(ref).max()
    h.close()
