"""Effect of the relative pivot tolerance of the A A^T factorisation (CUADMM_PIVOT_TOL) on (a) agreement with the
ADMM oracle (SuperLU solve of A A^T + 1e-15 I, no pivots dropped) at the full C2b size and (b) the stop iterations
of the bundled examples."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import oracle_np as onp
from util_problems import load_fixture, make_solver
from cuadmm_b200.synthetic import c2b_blocks, chain_sdp
P = chain_sdp(c2b_blocks(2000, 6, 60, 0), 700000, seed=0)
blk = np.ascontiguousarray(P["blk"], np.int32)
iters = 10
o = onp.ADMMOracle(P["vec_len"], P["con_num"], P["col_ptrs"], P["row_ids"], P["vals"], P["b_idx"], P["b_val"],
                   P["C_idx"], P["C_val"], blk, project=lambda v: onp.project_svec_cpp(blk, v, min(30, os.cpu_count() or 1)))
X, y, S, it = o.solve(iters, 1e-12, 500, 50, 100, 11000, 1.05)
out = []
for tol in sys.argv[1:] or ["1e-11", "1e-12", "1e-13", "1e-14"]:
    os.environ["CUADMM_PIVOT_TOL"] = tol
    s = make_solver(P)
    s.solve(iters, 1e-12, 500, 50, 100, 11000, 1.05)
    row = {"tol": tol, "deficient": s.ysolve_stats()["deficient"]}
    for key in ["errRp", "errRd", "pobj", "dobj", "relgap"]:
        a, b = s.history(key), np.array(o.hist[key])
        row[key] = float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-9)))
    row["X"] = float(np.linalg.norm(s.X - X) / np.linalg.norm(X))
    r = s.run_iterations(40, sgs=True, profile=True)
    row["ysolve_ms"] = r["ysolve_ms"] / 40
    s.close()
    for name, stop in (("ros_2000", 12732), ("pusht_n10", 6149), ("planarhand_n1", 800)):
        s2 = make_solver(load_fixture(name))
        s2.solve(20000, 1e-3, 0, 50, 100, 11000, 1.05)
        row[name] = [int(s2.info_iter_num), stop, s2.ysolve_stats()["deficient"]]
        s2.close()
    print(json.dumps(row), flush=True)
    out.append(row)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "pivot_tol_probe.json"), "w"), indent=1)
