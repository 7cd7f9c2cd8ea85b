"""y-solve timing probe: device ms per solve and true residual for several sweep settings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, scipy.sparse as sp
import cuadmm_b200 as cu
import oracle_np as onp
from cuadmm_b200.synthetic import c2b_blocks, chain_sdp
from util_problems import load_fixture
import json
ENVS = json.loads(os.environ.get("PROBE_ENVS", "[{}]"))
for which in os.environ.get("PROBE_WHICH", "c2b,pusht_n10").split(","):
    P = chain_sdp(c2b_blocks(), 700000, seed=0) if which == "c2b" else load_fixture(which)
    normA, vals = onp.get_normA_fast(P["col_ptrs"], P["vals"])
    m, n = P["con_num"], P["vec_len"]
    A = sp.csr_matrix((vals, P["row_ids"], P["col_ptrs"]), shape=(m, n))
    rhs_h = A @ np.random.default_rng(0).standard_normal(n)
    rhs = torch.from_numpy(rhs_h).cuda(); y = torch.empty_like(rhs)
    for env in ENVS:
        for k in set().union(*[set(e) for e in ENVS]):
            os.environ.pop(k, None)
        os.environ.update(env)
        ys = cu.YSolve(m, n, P["col_ptrs"], P["row_ids"], vals)
        for _ in range(3): ys.solve_device(rhs.data_ptr(), y.data_ptr())
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): ys.solve_device(rhs.data_ptr(), y.data_ptr())
        e1.record(); torch.cuda.synchronize()
        yh = y.cpu().numpy()
        res = np.linalg.norm(A @ (A.T @ yh) - rhs_h) / np.linalg.norm(rhs_h)
        print(which, env, "solve %.3f ms" % (e0.elapsed_time(e1) / 10), "res %.1e" % res, ys.stats(), flush=True)
        ys.close()
