"""Large-block (n > 168) projection probe: time of the GEMM-only sign path vs Baseline A (the reference's
cuSOLVER Xsyevd + gemm stage, oracle/_ref) on the same input, accuracy against Baseline A / LAPACK, and the
projection identities at sizes where a CPU eigendecomposition is too slow."""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import cuadmm_b200 as cu
from cuadmm_b200.synthetic import random_svec
try:
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libcuadmm_ref.so"))
    ref.ref_proj_create.restype = C.c_void_p; ref.ref_proj_run.restype = C.c_double
except OSError:
    ref = None
cases = [[200] * 64, [800] * 16, [2000], [4000]]
if len(sys.argv) > 1: cases = [json.loads(a) for a in sys.argv[1:]]
out = []
for blk in cases:
    blk = np.array(blk, np.int32); n = int(blk[0]); nb1 = n * (n + 1) // 2
    x = np.concatenate([random_svec(blk[:1], seed=10 + k) for k in range(len(blk))])
    p = cu.Plan(blk)
    dx = torch.from_numpy(x).cuda(); dy = torch.empty_like(dx); dz = torch.empty_like(dx)
    for _ in range(2): p.project_device(dx.data_ptr(), dy.data_ptr())
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): p.project_device(dx.data_ptr(), dy.data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    # identities: idempotence, Moreau decomposition x = P(x) - P(-x), <P(x), P(-x)> = 0
    p.project_device(dy.data_ptr(), dz.data_ptr()); idem = float((dz - dy).norm() / dy.norm())
    dneg = -dx; p.project_device(dneg.data_ptr(), dz.data_ptr())
    moreau = float((dy - dz - dx).norm() / dx.norm()); orth = float(torch.dot(dy, dz).abs() / (dy.norm() * dz.norm()))
    row = {"blk": "%d x %d" % (len(blk), n), "ours_ms": ms, "F_alg_TFLOPs": (20 / 3) * n ** 3 * len(blk) / ms / 1e9,
           "idempotence": idem, "moreau": moreau, "orth": orth}
    if ref is not None:
        h = ref.ref_proj_create(blk.ctypes.data_as(C.POINTER(C.c_int)), len(blk), 15)
        ro = np.empty_like(x)
        row["baselineA_ms"] = ref.ref_proj_run(C.c_void_p(h), x.ctypes.data_as(C.POINTER(C.c_double)), ro.ctypes.data_as(C.POINTER(C.c_double)), 2)
        ref.ref_proj_destroy(C.c_void_p(h))
        row["rel_err_vs_baselineA"] = float(np.linalg.norm(dy.cpu().numpy() - ro) / np.linalg.norm(ro))
        row["speedup_vs_A"] = row["baselineA_ms"] / ms
    print(json.dumps(row), flush=True)
    out.append(row)
    del p, dx, dy, dz
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "dense_probe.json"), "w"), indent=1)
