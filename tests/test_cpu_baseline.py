"""The C++ CPU baseline (oracle/cpu_baseline.cpp: the reference's dsyevd thread-pool projection path) against the
reference's known-answer eigenproblem and against the numpy oracle.  CPU only."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_np as onp  # noqa: E402


def test_known_answer_4x4():
    # test/eig_cpu_test.hpp:7-66 of the reference: eigenvalues (5 -+ sqrt5)/2, (11 -+ sqrt37)/2
    M = np.array([[4, 1, 2, 2], [1, 4, 1, 2], [2, 1, 4, 1], [2, 2, 1, 4]], float)
    out, eig = onp.project_svec_cpp([4], onp.svec(M), 1, want_eig=True)
    exp = np.sort([(5 - 5 ** 0.5) / 2, (5 + 5 ** 0.5) / 2, (11 - 37 ** 0.5) / 2, (11 + 37 ** 0.5) / 2])
    assert np.allclose(eig, exp, rtol=1e-13)
    assert np.allclose(out, onp.svec(M), rtol=1e-12)          # M is positive definite: projection = identity


def test_matches_numpy_oracle_and_thread_ranges():
    rng = np.random.default_rng(3)
    blk = rng.integers(1, 40, 57).astype(np.int32)
    x = np.concatenate([onp.svec((lambda G: (G + G.T) / 2)(rng.standard_normal((n, n)))) for n in blk])
    ref, reig = onp.project_svec(blk, x, want_eig=True)
    for threads in (1, 3, 8):
        out, eig = onp.project_svec_cpp(blk, x, threads, want_eig=True)
        assert np.linalg.norm(out - ref) <= 1e-13 * np.linalg.norm(ref)
        assert np.allclose(eig, reig, rtol=0, atol=1e-12)
    lib = onp.cpu_baseline_lib()
    for count, T in [(14, 3), (100, 30), (5, 8), (7, 1), (0, 4)]:
        p = (C.c_int * (T + 1))()
        lib.cb_thread_ranges(count, T, p)
        got = [(p[t], p[t + 1]) for t in range(T)]
        assert got == onp.thread_ranges(count, T)            # restates src/duo_solver.cu:346-371
