"""End-to-end parity of the solver (C ABI / cuadmm_exe) against (a) the numpy restatement of
SDPSolver::solve (oracle_np.ADMMOracle, LAPACK projection + host sparse solve) and (b) the
iteration-indexed trajectories in the reference's own committed logs."""
import os
import subprocess

import numpy as np
import pytest

import cuadmm_b200 as cu
import oracle_np as onp
from util_problems import GOLD, ROOT, load_fixture, make_solver, parse_log, synthetic_sdp, write_txt

pytestmark = pytest.mark.gpu


def _oracle(P, **kw):
    return onp.ADMMOracle(P["vec_len"], P["con_num"], P["col_ptrs"], P["row_ids"], P["vals"], P["b_idx"], P["b_val"],
                          P["C_idx"], P["C_val"], P["blk"], **kw)


def _compare_histories(s, o, n, rtol):
    for key in ["errRp", "errRd", "pobj", "dobj", "relgap", "sig"]:
        a, b = s.history(key)[:n], np.array(o.hist[key][:n])
        scale = np.maximum(np.abs(b), 1e-12 if key.startswith("err") or key == "relgap" else 1e-9)
        assert np.max(np.abs(a - b) / scale) < rtol, key


@pytest.mark.parametrize("mode", ["sgs", "admm"])
def test_small_synthetic_matches_oracle_iteration_by_iteration(mode):
    P = synthetic_sdp([5, 8, 3, 12, 20, 6, 6, 33, 7], m=120, seed=3)
    switch = 11000 if mode == "sgs" else 1
    iters = 60
    s = make_solver(P)
    s.solve(iters, 1e-12, 500, 50, 100, switch, 1.05)
    o = _oracle(P)
    X, y, S, it = o.solve(iters, 1e-12, 500, 50, 100, switch, 1.05)
    assert s.info_iter_num == it == iters
    _compare_histories(s, o, iters, 1e-7)
    assert np.linalg.norm(s.X - X) <= 1e-8 * np.linalg.norm(X)
    assert np.linalg.norm(s.S - S) <= 1e-8 * max(np.linalg.norm(S), 1e-12)
    # y: the half-step of the iteration in which the loop breaks (sGS) / the best iterate (ADMM)
    assert np.linalg.norm(o.A.T @ (s.y * o.normA / o.Cscale) - o.A.T @ (y * o.normA / o.Cscale)) <= \
        1e-7 * max(np.linalg.norm(o.A.T @ (y * o.normA / o.Cscale)), 1e-12)


def test_converges_to_known_optimum_and_counts_match_oracle():
    # time-to-1e-6 KKT: iteration count within +-2 % of the reference algorithm (BASELINE north_star)
    P = synthetic_sdp([6, 10, 4, 15, 9], m=40, seed=11)
    s = make_solver(P)
    s.solve(8000, 1e-6, 500, 50, 100, 11000, 1.05)
    o = _oracle(P)
    X, y, S, it = o.solve(8000, 1e-6, 500, 50, 100, 11000, 1.05)
    assert it < 8000, "oracle did not converge"
    assert abs(s.info_iter_num - it) <= max(1, int(0.02 * it))
    pobj = s.history("pobj")[-1]
    assert abs(pobj - P["pstar"]) <= 1e-4 * (1 + abs(P["pstar"]))
    assert max(s.history("errRp")[-1], s.history("errRd")[-1], s.history("relgap")[-1]) < 1e-6
    assert np.linalg.norm(s.X - X) <= 1e-6 * np.linalg.norm(X)


def _check_against_log(name, logfile, max_iter, tol, switch, upto, rtol):
    P = load_fixture(name)
    s = make_solver(P)
    s.solve(max_iter, tol, 0, 50, 100, switch, 1.05)
    rows = [r for r in parse_log(os.path.join(GOLD, logfile)) if 0 < r["it"] <= upto]
    assert len(rows) >= 3
    H = {k: s.history(k) for k in ["errRp", "errRd", "pobj", "dobj", "relgap", "sig"]}
    for r in rows:
        k = r["it"] - 1
        for key in ["errRp", "errRd", "relgap"]:
            if r[key] > 1e-9:           # below that the logs show rounding noise of the residual
                assert abs(H[key][k] - r[key]) <= rtol * r[key] + 6e-3 * r[key], (name, r["it"], key, H[key][k], r[key])
        for key in ["pobj", "dobj"]:
            assert abs(H[key][k] - r[key]) <= rtol * abs(r[key]) + 6e-5 * abs(r[key]) + 1e-9, (name, r["it"], key)
        assert abs(H["sig"][k] - r["sig"]) <= 0.06 * r["sig"]
    return s


def test_ros_2000_trajectory_matches_reference_log():
    # examples/benchmarks/ros_2000/sGS-cuADMM.log (main.cu defaults: tol 1e-3, switch_admm 5000)
    _check_against_log("ros_2000", "ros_2000_sgs.log", 1000, 1e-3, 5000, upto=1000, rtol=2e-2)


def test_pusht_n10_trajectory_matches_reference_log():
    # examples/benchmarks/PushT_N=10_MOMENT/sGS-cuADMM.log: 6015 blocks, rank-deficient A A^T, dense tail
    _check_against_log("pusht_n10", "pusht_n10_sgs.log", 600, 1e-3, 5000, upto=600, rtol=2e-2)


def test_rose13_single_block_trajectory():
    # examples/plato/logs/rose13.log: one block n=105
    _check_against_log("rose13", "rose13.log", 300, 1e-3, 5000, upto=300, rtol=5e-2)


def test_warm_restart_and_history_api():
    P = synthetic_sdp([6, 7, 9], m=30, seed=5)
    s = make_solver(P)
    s.solve(50, 1e-9)
    X1, y1, S1 = s.X, s.y, s.S
    sig = s.history("sig")[-1]
    s2 = make_solver(P, X=X1, y=y1, S=S1, sig=sig)
    # warm start from iteration 50 resumes with a much smaller residual than a cold start
    s2.solve(1, 1e-9)
    cold = make_solver(P); cold.solve(1, 1e-9)
    assert s2.history("errRp")[0] < cold.history("errRp")[0]
    assert len(s.history("pobj")) == 50 and s.launches > 0
    t = s.times()
    assert t["total"] > 0 and t["init"] > 0


def test_exe_reads_txt_and_writes_x_opt(tmp_path):
    P = synthetic_sdp([5, 6, 4, 10], m=25, seed=2)
    d = str(tmp_path / "prob")
    write_txt(P, d)
    exe = os.path.join(ROOT, "cuadmm_b200", "lib", "cuadmm_exe")
    out = subprocess.run([exe, d + "/", "--max-iter", "300", "--tol", "1e-5"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "vector length: %d" % P["vec_len"] in out.stdout and "Solver ended" in out.stdout
    rows = parse_log(out.stdout)
    assert rows and rows[0]["it"] == 0
    X = np.loadtxt(os.path.join(d, "X_opt.txt"))
    s = make_solver(P); s.solve(300, 1e-5, 0, 50, 100, 5000, 1.05)
    assert X.shape == (P["vec_len"],) and np.allclose(X, s.X, rtol=1e-9, atol=1e-12)
    # directory without trailing slash is accepted too (the reference needs the slash)
    out2 = subprocess.run([exe, d, "--max-iter", "5", "--quiet", "--no-output"], capture_output=True, text=True, timeout=300)
    assert out2.returncode == 0
    bad = subprocess.run([exe, str(tmp_path / "nope") + "/"], capture_output=True, text=True)
    assert bad.returncode != 0 and "could not open file" in bad.stderr
