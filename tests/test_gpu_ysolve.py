"""GPU parity of the AA^T y-solve (C ABI) against a host sparse solve of (A A^T + eps I) y = rhs."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import cuadmm_b200 as cu
import oracle_np as onp
from util_problems import load_fixture

pytestmark = pytest.mark.gpu


def _solve_check(A, eps=1e-15, tol=1e-10, rhs_in_range=True, seed=0):
    A = A.tocsr(); A.sort_indices()
    m, n = A.shape
    ys = cu.YSolve(m, n, A.indptr, A.indices, A.data, eps=eps)
    rng = np.random.default_rng(seed)
    rhs = A @ rng.standard_normal(n) if rhs_in_range else rng.standard_normal(m)
    y = ys.solve_host(rhs)
    M = (A @ A.T + eps * sp.eye(m)).tocsc()
    res = np.linalg.norm(M @ y - rhs) / np.linalg.norm(rhs)
    assert res < tol, (res, ys.stats())
    return ys, y, rhs, M


def test_reference_dense_golden():
    # test/cholesky_cpu_test.hpp:3-55: A = 11^T + I => A A^T = 6*11^T + I ; rhs = 25 -> y = 1 (tol 1e-5 there)
    A = sp.csr_matrix(np.ones((4, 4)) + np.eye(4))
    ys = cu.YSolve(4, 4, A.indptr, A.indices, A.data, eps=1e-16)
    assert np.allclose(ys.solve_host(np.full(4, 25.0)), 1.0, atol=1e-12)
    assert sorted(ys.perm().tolist()) == [0, 1, 2, 3]


def test_diagonal_aat():
    # max-cut / ros_2000 style: every constraint touches its own entries => A A^T diagonal
    P = load_fixture("ros_2000")
    normA, vals = onp.get_normA_fast(P["col_ptrs"], P["vals"])
    A = sp.csr_matrix((vals, P["row_ids"], P["col_ptrs"]), shape=(P["con_num"], P["vec_len"]))
    ys, y, rhs, M = _solve_check(A, tol=1e-13)
    st = ys.stats()
    assert st["nnz_L"] == P["con_num"] and st["dense_tail"] == 0 and st["levels"] == 1


@pytest.mark.parametrize("m,n,density,seed", [(300, 500, 0.01, 1), (2000, 3000, 0.002, 2), (1200, 900, 0.004, 3)])
def test_random_sparse(m, n, density, seed):
    A = sp.random(m, n, density=density, random_state=seed, format="csr") + sp.eye(m, n, format="csr")
    # m > n makes A A^T rank deficient (redundant constraints are pinned): then rhs must be in range(A)
    _solve_check(A, eps=1e-15, tol=1e-11 if m <= n else 1e-9, rhs_in_range=(m > n), seed=seed)


def test_matches_oracle_solver_on_truss8_and_biggs():
    for name in ["truss8", "biggs", "hinf12"]:
        P = load_fixture(name)
        normA, vals = onp.get_normA_fast(P["col_ptrs"], P["vals"])
        A = sp.csr_matrix((vals, P["row_ids"], P["col_ptrs"]), shape=(P["con_num"], P["vec_len"]))
        ys, y, rhs, M = _solve_check(A, tol=1e-9)
        ref = onp.AATSolver(A, 1e-15).solve(rhs)
        assert np.linalg.norm(A.T @ (y - ref)) <= 1e-8 * np.linalg.norm(A.T @ ref)


def test_dense_tail_path_pusht():
    # PushT_N=10: a shared moment entry couples 2720 constraints => dense trailing block; A A^T is
    # rank deficient by ~600, the redundant directions are pinned (see chol_host.cpp).  Only A^T y is
    # determined, which is all the iteration uses.
    P = load_fixture("pusht_n10")
    normA, vals = onp.get_normA_fast(P["col_ptrs"], P["vals"])
    A = sp.csr_matrix((vals, P["row_ids"], P["col_ptrs"]), shape=(P["con_num"], P["vec_len"]))
    ys = cu.YSolve(P["con_num"], P["vec_len"], A.indptr, A.indices, A.data, eps=1e-15)
    st = ys.stats()
    assert st["dense_tail"] > 1000 and st["levels"] < 400
    rng = np.random.default_rng(0)
    x = rng.standard_normal(P["vec_len"])
    rhs = A @ x
    y = ys.solve_host(rhs)
    # A A^T y = rhs in the least-squares sense: A^T y is the projection of x onto range(A^T)
    assert np.linalg.norm(A @ (A.T @ y) - rhs) <= 1e-9 * np.linalg.norm(rhs)
    y2 = ys.solve_host(rhs)
    assert np.array_equal(y, y2)                        # deterministic


def test_forced_dense_tail_equals_sparse_path(monkeypatch):
    A = sp.random(1500, 2500, density=0.004, random_state=5, format="csr") + sp.eye(1500, 2500, format="csr")
    A = A.tocsr(); A.sort_indices()
    rng = np.random.default_rng(1)
    rhs = rng.standard_normal(1500)
    monkeypatch.setenv("CUADMM_YSOLVE_MAX_TAIL", "0")
    y_sparse = cu.YSolve(1500, 2500, A.indptr, A.indices, A.data).solve_host(rhs)
    monkeypatch.setenv("CUADMM_YSOLVE_MAX_TAIL", "700")
    monkeypatch.setenv("CUADMM_YSOLVE_MIN_DEPTH", "4")
    ys = cu.YSolve(1500, 2500, A.indptr, A.indices, A.data)
    assert ys.stats()["dense_tail"] > 0
    y_tail = ys.solve_host(rhs)
    assert np.linalg.norm(y_tail - y_sparse) <= 1e-10 * np.linalg.norm(y_sparse)


def test_linearity_and_repeatability_at_size():
    # full-size C2b-like operator: 2000 blocks U{6..60}, m = 700k chain-structured constraints
    from util_problems import chain_sdp
    rng = np.random.default_rng(0)
    blk = rng.integers(6, 61, 2000)
    P = chain_sdp(blk, 700000, seed=0)
    normA, vals = onp.get_normA_fast(P["col_ptrs"], P["vals"])
    m, n = P["con_num"], P["vec_len"]
    A = sp.csr_matrix((vals, P["row_ids"], P["col_ptrs"]), shape=(m, n))
    ys = cu.YSolve(m, n, A.indptr, A.indices, A.data)
    # ~2 % of these constraints are redundant (as in the SPOT data): right-hand sides live in range(A)
    r1, r2 = A @ rng.standard_normal(n), A @ rng.standard_normal(n)
    y12 = ys.solve_host(r1 + 3 * r2)
    y = ys.solve_host(r1) + 3 * ys.solve_host(r2)
    # y itself is huge along the near-redundant directions; the iteration only ever uses A^T y
    assert np.linalg.norm(A.T @ (y12 - y)) <= 1e-9 * np.linalg.norm(A.T @ y)
    # backward-stable residual: eps * |y| with |y| ~ 1e5 |rhs| here (near-redundant constraints)
    assert np.linalg.norm(A @ (A.T @ y12) - (r1 + 3 * r2)) <= 1e-15 * np.sqrt(m) * np.linalg.norm(y12)
    assert np.linalg.norm(A @ (A.T @ y12) - (r1 + 3 * r2)) <= 1e-6 * np.linalg.norm(r1 + 3 * r2)
    assert np.array_equal(ys.solve_host(r1), ys.solve_host(r1))
    assert ys.stats()["nnz_L"] < 2e7


def test_empty_and_tiny():
    ys = cu.YSolve(1, 3, [0, 2], [0, 2], [3.0, 4.0], eps=0.0)
    assert np.allclose(ys.solve_host([50.0]), [2.0])
    with pytest.raises(cu.CuadmmError):
        cu.YSolve(2, 3, [0, 1, 5], [0, 1], [1.0, 1.0])
