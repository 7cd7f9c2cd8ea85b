"""y-solve timing probe on the bench workload: device ms per solve for several sweep settings."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import cuadmm_b200 as cu
import oracle_np as onp
from cuadmm_b200.synthetic import c2b_blocks, chain_sdp
which = sys.argv[1] if len(sys.argv) > 1 else "c2b"
if which == "c2b":
    P = chain_sdp(c2b_blocks(), 700000, seed=0)
else:
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util_problems import load_fixture
    P = load_fixture(which)
normA, vals = onp.get_normA_fast(P["col_ptrs"], P["vals"])
m, n = P["con_num"], P["vec_len"]
rhs = torch.randn(m, dtype=torch.float64, device="cuda"); y = torch.empty_like(rhs)
for env in [dict(), dict(CUADMM_SWEEP_BACKOFF_NS="0"), dict(CUADMM_SWEEP_BACKOFF_NS="400"), dict(CUADMM_SWEEP_MAXGRID="296"),
            dict(CUADMM_SWEEP_MAXGRID="148", CUADMM_SWEEP_BACKOFF_NS="200"), dict(CUADMM_YSOLVE_MAX_TAIL="12000"),
            dict(CUADMM_MD_MODE="classic"), dict(CUADMM_MD_SLACK="100"), dict(CUADMM_YSOLVE_MAX_TAIL="0")]:
    for k in ["CUADMM_SWEEP_BACKOFF_NS", "CUADMM_SWEEP_MAXGRID", "CUADMM_YSOLVE_MAX_TAIL", "CUADMM_MD_MODE", "CUADMM_MD_SLACK"]:
        os.environ.pop(k, None)
    os.environ.update(env)
    t = time.time()
    ys = cu.YSolve(m, n, P["col_ptrs"], P["row_ids"], vals)
    tc = time.time() - t
    for _ in range(3): ys.solve_device(rhs.data_ptr(), y.data_ptr())
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ys.solve_device(rhs.data_ptr(), y.data_ptr())
    e1.record(); torch.cuda.synchronize()
    print(env, "create %.1fs" % tc, "solve %.3f ms" % (e0.elapsed_time(e1) / 10), ys.stats(), flush=True)
    ys.close()
