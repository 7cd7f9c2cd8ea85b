"""GPU parity of the CSR SpMV (C ABI) against the serial C oracle / scipy."""
import numpy as np
import pytest
import scipy.sparse as sp

import cuadmm_b200 as cu
from conftest import ip, dp

pytestmark = pytest.mark.gpu


def _rand_csr(rows, cols, density, seed, long_rows=0):
    """random CSR built straight from COO triplets (LIL assignment is far too slow at size)"""
    rng = np.random.default_rng(seed)
    nnz = max(1, int(rows * cols * density))
    r = rng.integers(0, rows, nnz); c = rng.integers(0, cols, nnz); v = rng.standard_normal(nnz)
    for _ in range(long_rows):
        k = min(cols, 3000)
        r = np.concatenate([r, np.full(k, rng.integers(0, rows))])
        c = np.concatenate([c, rng.choice(cols, k, replace=False)])
        v = np.concatenate([v, rng.standard_normal(k)])
    M = sp.csr_matrix((v, (r, c)), shape=(rows, cols))
    M.sum_duplicates(); M.sort_indices()
    return M


def test_reference_golden_example():
    # test/cusparse_test.hpp:41-94 (2*A*x + 3*y), asserted here
    A = cu.SpMV(4, 4, [0, 1, 2, 5, 6], [0, 1, 0, 2, 3, 1], [10.0, 20, 30, 40, 50, 60])
    y = A.apply_host([1.0, 2, 3, 4], alpha=2.0, beta=3.0, y=[5.0, 6, 7, 8])
    assert y.tolist() == [35.0, 98.0, 721.0, 264.0]


@pytest.mark.parametrize("rows,cols,density,long_rows", [(1000, 700, 0.003, 0), (5000, 20000, 0.0005, 3),
                                                          (300, 300, 0.2, 0), (20000, 900, 0.05, 2), (7, 5, 0.5, 0)])
def test_random_matrices(rows, cols, density, long_rows):
    M = _rand_csr(rows, cols, density, 1, long_rows)
    rng = np.random.default_rng(2)
    x = rng.standard_normal(cols); y0 = rng.standard_normal(rows)
    A = cu.SpMV(rows, cols, M.indptr, M.indices, M.data)
    y = A.apply_host(x, alpha=-1.0, beta=0.0)
    ref = -(M @ x)
    scale = (abs(M) @ abs(x)) + 1e-300
    assert np.max(np.abs(y - ref) / scale) < 1e-14
    y = A.apply_host(x, alpha=2.0, beta=3.0, y=y0)
    assert np.max(np.abs(y - (2 * (M @ x) + 3 * y0)) / (2 * scale + 3 * abs(y0))) < 1e-14


def test_empty_rows_and_empty_matrix():
    A = cu.SpMV(5, 3, [0, 0, 0, 2, 2, 2], [0, 2], [1.5, -2.0])
    y = A.apply_host([1.0, 1.0, 1.0])
    assert y.tolist() == [0, 0, -0.5, 0, 0]
    A = cu.SpMV(3, 3, [0, 0, 0, 0], [], [])
    assert A.apply_host([1.0, 2, 3], beta=2.0, y=[1.0, 1, 1]).tolist() == [2.0, 2.0, 2.0]


def test_linearity_at_size():
    M = _rand_csr(200000, 100000, 2.5e-5, 3, 4)
    A = cu.SpMV(*M.shape, M.indptr, M.indices, M.data)
    rng = np.random.default_rng(4)
    x1, x2 = rng.standard_normal(M.shape[1]), rng.standard_normal(M.shape[1])
    y12 = A.apply_host(x1 + 2 * x2)
    y = A.apply_host(x1) + 2 * A.apply_host(x2)
    assert np.allclose(y12, y, rtol=0, atol=1e-12 * np.abs(y).max())
    assert np.allclose(A.apply_host(x1), M @ x1, rtol=0, atol=1e-13 * np.abs(M @ x1).max())


def test_large_short_row_operator_with_medium_rows():
    """the bench operator's regime: >= 262,144 rows with ~1-2 entries each run one lane per row (two rows in flight, epilogue
    operands prefetched); rows of 17..256 entries take the warp-per-row part there (threshold 16), longer ones always"""
    rows, cols = 300000, 150000
    rng = np.random.default_rng(5)
    nnz = 450000
    r = rng.integers(0, rows, nnz); c = rng.integers(0, cols, nnz); v = rng.standard_normal(nnz)
    for k in (17, 40, 200, 256, 257, 1000):          # a few medium / long rows
        row = rng.integers(0, rows)
        r = np.concatenate([r, np.full(k, row)]); c = np.concatenate([c, rng.choice(cols, k, replace=False)])
        v = np.concatenate([v, rng.standard_normal(k)])
    M = sp.csr_matrix((v, (r, c)), shape=(rows, cols)); M.sum_duplicates(); M.sort_indices()
    A = cu.SpMV(rows, cols, M.indptr, M.indices, M.data)
    x = rng.standard_normal(cols); y0 = rng.standard_normal(rows)
    scale = (abs(M) @ abs(x)) + 1e-300
    y = A.apply_host(x, alpha=1.0, beta=0.0)
    assert np.max(np.abs(y - M @ x) / scale) < 1e-14
    y = A.apply_host(x, alpha=-0.5, beta=2.0, y=y0)          # beta != 0: the prefetched operand is y itself
    assert np.max(np.abs(y - (-0.5 * (M @ x) + 2 * y0)) / (0.5 * scale + 2 * abs(y0))) < 1e-14
