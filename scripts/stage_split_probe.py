"""Stage split of one sGS-ADMM iteration on the bench workload C2b (ms: projection / two y-solves / SpMV + rest) for the
build selected by CUADMM_LIB_PATH and the environment switches under test.  One JSON line."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuadmm_b200 as cu
from cuadmm_b200.synthetic import c2b_blocks, chain_sdp
P = chain_sdp(c2b_blocks(), 700000, seed=0)
s = cu.Solver(verbose=False)
s.init(15, 30, P["vec_len"], P["con_num"], P["col_ptrs"], P["row_ids"], P["vals"], P["b_idx"], P["b_val"],
       P["C_idx"], P["C_val"], P["blk"], None, None, None, 1.0)
s.solve(1, -1.0, 500, 50, 100, 1 << 30, 1.05)
s.run_iterations(int(os.environ.get("WARM", "100")), sgs=True)
n = int(os.environ.get("ITERS", "100"))
best = None
for rep in range(3):
    r = s.run_iterations(n, sgs=True, profile=True)
    row = {k: r[k] / n for k in ("total_ms", "projection_ms", "ysolve_ms", "other_ms")}
    if best is None or row["other_ms"] < best["other_ms"]: best = row
t = s.run_iterations(n, sgs=True)
best["unprofiled_total_ms"] = t["total_ms"] / n
best["lib"] = os.environ.get("CUADMM_LIB_PATH", "default")
print(json.dumps(best), flush=True)
s.close()
