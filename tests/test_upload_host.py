"""Host half of the symmetric-upload scheme (csrc/upload.cu): the multi-threaded test that decides whether the blocks below
the block diagonal may be mirrored on the device instead of crossing PCIe.  Pure host code: runs without a GPU."""
import ctypes

import numpy as np
import pytest

from ccqppy_b200 import _capi


def is_sym(A, threads=0):
    A = np.ascontiguousarray(A, dtype=np.float64)
    return _capi.load().ccqp_host_matrix_is_block_symmetric(ctypes.c_void_p(A.ctypes.data), A.shape[0], A.shape[1], threads)


@pytest.fixture(params=["avx2", "scalar"], autouse=True)
def code_path(request, monkeypatch):
    """Every test runs through both inner loops of csrc/symcheck.h: the AVX2 4 x 4 transposes and the portable one."""
    monkeypatch.setenv("CCQP_SYMCHECK_SCALAR", "1" if request.param == "scalar" else "0")


def test_block_rows_constant():
    B = _capi.load().ccqp_upload_block_rows()
    assert B >= 64 and B % 32 == 0 and B % 4 == 0


@pytest.mark.parametrize("n", [2048, 2500, 3071, 4096])
def test_symmetric_matrices_pass(n):
    rng = np.random.default_rng(n)
    G = rng.standard_normal((n, n))
    A = G + G.T
    assert is_sym(A) == 1
    assert is_sym(A, threads=1) == 1
    assert is_sym(A, threads=3) == 1


@pytest.mark.parametrize("n", [2048, 3071])
def test_every_mirrored_entry_is_checked(n):
    """One perturbed entry anywhere below the block diagonal must be found; entries inside the diagonal blocks are uploaded
    as they are (both triangles), so a difference there must NOT reject the matrix."""
    B = _capi.load().ccqp_upload_block_rows()
    rng = np.random.default_rng(7)
    G = rng.standard_normal((n, n))
    A = G + G.T
    for _ in range(25):
        i = int(rng.integers(B, n))
        j = int(rng.integers(0, (i // B) * B))
        C = A.copy()
        C[i, j] = np.nextafter(C[i, j], np.inf)
        assert is_sym(C) == 0, (i, j)
        C = A.copy()
        C[j, i] = np.nextafter(C[j, i], -np.inf)            # the partner above the diagonal
        assert is_sym(C) == 0, (i, j)
    corners = [(B, 0), (B, B - 1), (n - 1, 0), (n - 1, ((n - 1) // B) * B - 1), (2 * B - 1, B - 1)]
    for i, j in corners:
        C = A.copy()
        C[i, j] += 1.0
        assert is_sym(C) == 0, (i, j)
    C = A.copy()
    C[5, 3] += 1.0                                           # inside the first diagonal block
    C[B + 7, B + 2] -= 1.0                                   # inside the second
    assert is_sym(C) == 1


def test_nan_and_signed_zero_reject():
    n = 2048
    B = _capi.load().ccqp_upload_block_rows()
    A = np.ones((n, n))
    assert is_sym(A) == 1
    C = A.copy(); C[B + 1, 2] = np.nan; C[2, B + 1] = np.nan      # a NaN mirrored by a NaN is still not "equal"
    assert is_sym(C) == 0
    C = np.zeros((n, n)); C[B + 1, 2] = -0.0                      # -0.0 below, +0.0 above: the mirror would flip the sign bit
    assert is_sym(C) == 0
    C[2, B + 1] = -0.0
    assert is_sym(C) == 1


def test_leading_dimension_and_bad_arguments():
    n, lda = 2100, 2300
    rng = np.random.default_rng(1)
    G = rng.standard_normal((n, n))
    buf = np.full((n, lda), np.nan)
    buf[:, :n] = G + G.T
    lib = _capi.load()
    assert lib.ccqp_host_matrix_is_block_symmetric(ctypes.c_void_p(buf.ctypes.data), n, lda, 0) == 1
    assert lib.ccqp_host_matrix_is_block_symmetric(None, n, lda, 0) == -1
    assert lib.ccqp_host_matrix_is_block_symmetric(ctypes.c_void_p(buf.ctypes.data), n, n - 1, 0) == -1


def test_random_single_entry_perturbations_property():
    """Hypothesis-style sweep without the dependency: sizes that are and are not multiples of the 4 x 4 / 64 x 64 tiles, one
    flipped mantissa bit at a random mirrored position, at a random thread count."""
    B = _capi.load().ccqp_upload_block_rows()
    rng = np.random.default_rng(2024)
    for n in (B + 1, B + 3, B + 64, 2 * B + 5, 2 * B + 66, 3 * B - 1):
        G = rng.standard_normal((n, n))
        A = G + G.T
        assert is_sym(A, threads=int(rng.integers(1, 9))) == 1
        for _ in range(6):
            i = int(rng.integers(B, n))
            j = int(rng.integers(0, (i // B) * B))
            C = A.copy()
            bits = C[i:i + 1, j:j + 1].view(np.uint64)
            bits ^= np.uint64(1) << np.uint64(int(rng.integers(0, 52)))
            assert is_sym(C, threads=int(rng.integers(1, 9))) == 0, (n, i, j)
