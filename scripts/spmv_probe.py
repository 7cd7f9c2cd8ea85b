"""SpMV tuning probe on the bench operator (A: 700k x 1.4M, At its transpose): us per plain product for a few
lanes-per-row / grid-cap settings (CUADMM_SPMV_G, CUADMM_SPMV_CAP), with the algorithmic GB/s of SURVEY 8d."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch, scipy.sparse as sp
import cuadmm_b200 as cu
from cuadmm_b200.synthetic import c2b_blocks, chain_sdp
P = chain_sdp(c2b_blocks(), 700000, seed=0)
m, n = P["con_num"], P["vec_len"]
A = sp.csr_matrix((P["vals"], P["row_ids"], P["col_ptrs"]), shape=(m, n)); At = A.T.tocsr(); At.sort_indices()
out = []
for name, M in (("A", A), ("At", At)):
    x = torch.randn(M.shape[1], dtype=torch.float64, device="cuda"); y = torch.zeros(M.shape[0], dtype=torch.float64, device="cuda")
    ref = None
    for G, cap in ((0, 8), (0, 5), (0, 6), (1, 8), (1, 5), (1, 16), (2, 5), (2, 16), (4, 8)):
        os.environ.pop("CUADMM_SPMV_G", None)
        if G: os.environ["CUADMM_SPMV_G"] = str(G)
        os.environ["CUADMM_SPMV_CAP"] = str(cap)
        S = cu.SpMV(M.shape[0], M.shape[1], M.indptr.astype(np.int32), M.indices.astype(np.int32), M.data)
        for _ in range(5): S.apply_device(x.data_ptr(), y.data_ptr())
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): S.apply_device(x.data_ptr(), y.data_ptr())
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 50
        yh = y.cpu().numpy()
        if ref is None: ref = yh.copy()
        alg = 12 * M.nnz + 4 * (M.shape[0] + 1) + 8 * M.shape[0] + 8 * M.shape[1]
        row = {"matrix": name, "G": G or "auto", "cap": cap, "us": us, "alg_GBs": alg / us / 1e3, "max_diff_vs_default": float(np.abs(yh - ref).max())}
        print(json.dumps(row), flush=True); out.append(row)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "spmv_probe.json"), "w"), indent=1)
