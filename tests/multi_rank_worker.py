"""Worker of the multi-rank parity tests (tests/test_gpu_multi.py): one process per rank, no torch.distributed —
the ranks only share the 128-byte job id (env JOB_ID, hex).  With fewer GPUs than ranks the ranks share GPU 0
(the peer transport works between processes on one device as well), so the sharded data path is exercised on a
1-GPU box too.  Rank 0 also solves the problem unsharded and compares."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import cuadmm_b200 as cu  # noqa: E402
from cuadmm_b200.synthetic import chain_sdp, c2b_blocks, random_sdp  # noqa: E402

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
job = bytes.fromhex(os.environ["JOB_ID"])
ndev = cu.device_count()
dev = rank % max(ndev, 1)
case = os.environ.get("CASE", "sgs")
nblk = int(os.environ.get("NBLK", "120")); con = int(os.environ.get("CON", "20000"))
if case == "mixed":      # a few large blocks (dense sign path) + small ones, random coupling
    blk = np.array([200, 30, 12, 180, 45, 7, 60, 220, 9, 33], np.int32)
    P = random_sdp(blk, 3000, seed=1)
else:
    P = chain_sdp(c2b_blocks(nblk, 6, 60, 0), con, seed=0)


def make(distributed):
    s = cu.Solver(verbose=False)
    s.set_device(dev)
    if distributed:
        s.set_distributed(rank, world, job)
    s.init(15, 30, P["vec_len"], P["con_num"], P["col_ptrs"], P["row_ids"], P["vals"], P["b_idx"], P["b_val"],
           P["C_idx"], P["C_val"], P["blk"])
    return s


def run(s):
    out = {}
    if case in ("sgs", "mixed"):
        s.solve(40, 1e-12, 500, 50, 100, 11000, 1.05)
    elif case == "admm":            # sGS -> ADMM switch at 12, best-iterate tracking, sigma updates every 5
        s.solve(60, 1e-12, 20, 5, 10, 12, 1.05)
    elif case == "tol":             # stops by tolerance in the middle of a log batch
        s.solve(4000, float(os.environ.get("TOL", "2e-3")), 500, 50, 100, 11000, 1.05)
    elif case == "restart":         # warm restart through set_XyS / if_first=False
        s.solve(15, 1e-12, 500, 50, 100, 11000, 1.05)
        X, y, S = s.X, s.y, s.S
        s.set_XyS(X, y, S, 1.3)
        s.solve(10, 1e-12, 500, 50, 100, 11000, 1.05, if_first=False)
    out["iters"] = s.info_iter_num
    out["X"], out["y"], out["S"] = s.X, s.y, s.S
    for k in ["errRp", "errRd", "pobj", "dobj", "sig", "relgap"]:
        out[k] = s.history(k)
    return out


sd = make(True)
d = run(sd)
ok = True
if rank == 0:
    s1 = make(False)
    r = run(s1)
    print("case", case, "world", world, "devices", ndev, "iters sharded", d["iters"], "single", r["iters"])
    ok = d["iters"] == r["iters"]
    if case == "tol":
        ok = ok and d["iters"] % 100 not in (0, 1) and d["iters"] < 4000     # really stopped inside a batch
    for k in ["errRp", "errRd", "pobj", "dobj", "sig", "relgap"]:
        if len(d[k]) != len(r[k]):
            ok = False
            continue
        err = float(np.max(np.abs(d[k] - r[k]) / np.maximum(np.abs(r[k]), 1e-9))) if len(r[k]) else 0.0
        print(" ", k, "max rel diff vs 1 GPU %.2e" % err)
        ok = ok and err < 1e-6
    # A^T y is what the iteration uses; y itself is only determined up to the redundant constraints of the data
    # (their pivots are regularised by eps = 1e-15), so rounding differences in the right-hand side show there
    import scipy.sparse as sp
    A = sp.csr_matrix((P["vals"], P["row_ids"], P["col_ptrs"]), shape=(P["con_num"], P["vec_len"]))
    d["Aty"], r["Aty"] = A.T @ d["y"], A.T @ r["y"]
    for k, tol in [("X", 1e-7), ("S", 1e-7), ("Aty", 1e-7), ("y", 1e-2)]:
        err = float(np.linalg.norm(d[k] - r[k]) / max(np.linalg.norm(r[k]), 1e-300))
        print(" ", k, "rel diff %.2e" % err)
        ok = ok and err < tol
    print("MULTI_RANK_OK" if ok else "MULTI_RANK_MISMATCH", flush=True)
sd.close()
sys.exit(0 if ok else 1)
