"""Per-CTA timeline of the packed subtree kernel (debug): where the time of the sweep goes."""
import os, sys, ctypes
os.environ["CUADMM_YSOLVE_TIMELINE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import cuadmm_b200 as cu
from cuadmm_b200 import capi
import oracle_np as onp
from cuadmm_b200.synthetic import c2b_blocks, chain_sdp
P = chain_sdp(c2b_blocks(), 700000, seed=0)
normA, vals = onp.get_normA_fast(P["col_ptrs"], P["vals"])
m, n = P["con_num"], P["vec_len"]
rhs = torch.randn(m, dtype=torch.float64, device="cuda"); y = torch.empty_like(rhs)
ys = cu.YSolve(m, n, P["col_ptrs"], P["row_ids"], vals)
for _ in range(5): ys.solve_device(rhs.data_ptr(), y.data_ptr())
torch.cuda.synchronize()
lib = capi.lib
for which in (0, 1):
    cnt = ctypes.c_int64(0)
    lib.cuadmm_debug_ysolve_timeline(ys.h, which, None, None, ctypes.c_int64(0), ctypes.byref(cnt))
    k = cnt.value
    tl = np.zeros(4 * k, np.int64); meta = np.zeros(3 * k, np.int64)
    lib.cuadmm_debug_ysolve_timeline(ys.h, which, tl.ctypes.data_as(ctypes.c_void_p), meta.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(k), ctypes.byref(cnt))
    tl = tl.reshape(k, 4); meta = meta.reshape(k, 3)
    t0 = tl[:, 0].min()
    print("sweep", which, "CTAs", k, "kernel span %.1f us" % ((tl[:, 3].max() - t0) / 1e3))
    order = np.argsort(-(tl[:, 3] - tl[:, 0]))
    for t in list(order[:12]) + list(range(0, k, max(1, k // 12))):
        print("  cta %4d chunks %3d rows %5d depth %3d | start %7.1f prologue %5.1f levels %6.1f epilogue %5.1f (us)" % (
            t, meta[t, 0], meta[t, 1], meta[t, 2], (tl[t, 0] - t0) / 1e3, (tl[t, 1] - tl[t, 0]) / 1e3,
            (tl[t, 2] - tl[t, 1]) / 1e3, (tl[t, 3] - tl[t, 2]) / 1e3))
