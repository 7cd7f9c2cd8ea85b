"""One-off GPU probe: projection timing on the C2 workloads vs Baseline A (reference cuSOLVER stage),
FP64 GEMM peak, sweep statistics.  Writes gpurun_out/probe.json."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import cuadmm_b200 as cu
import oracle_np as onp
from conftest import random_svec

out = {}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)

# FP64 peaks
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(2): torch.matmul(a, b)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
out["dgemm_8192_tflops"] = 2 * n ** 3 / best / 1e9
del a, b

def time_plan(blk, x, reps=20):
    p = cu.Plan(blk, device=0)
    dx = torch.from_numpy(x).cuda(); dy = torch.empty_like(dx)
    for _ in range(3): p.project_device(dx.data_ptr(), dy.data_ptr())
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): p.project_device(dx.data_ptr(), dy.data_ptr())
    e1.record(); torch.cuda.synchronize()
    _, eig, sweeps = p.project_eig_host(x)
    return e0.elapsed_time(e1) / reps, sweeps

ref = None
try:
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libcuadmm_ref.so"))
    ref.ref_proj_create.restype = C.c_void_p; ref.ref_proj_run.restype = C.c_double
except OSError as e:
    out["ref_error"] = str(e)

def time_ref(blk, x, reps=5):
    blk = np.ascontiguousarray(blk, np.int32)
    h = ref.ref_proj_create(blk.ctypes.data_as(C.POINTER(C.c_int)), len(blk), 15)
    o = np.zeros_like(x)
    ms = ref.ref_proj_run(C.c_void_p(h), x.ctypes.data_as(C.POINTER(C.c_double)), o.ctypes.data_as(C.POINTER(C.c_double)), reps)
    ref.ref_proj_destroy(C.c_void_p(h))
    return ms, o

rng = np.random.default_rng(0)
workloads = {
    "C2b_2000xU6_60": rng.integers(6, 61, 2000).astype(np.int32),
    "planarhand_n1": np.loadtxt(os.path.join(ROOT, "tests", "golden", "planarhand_n1_blk.txt"), dtype=np.int32),
    "pendulum_N80": np.array([55] * 80 + [10] * 159, np.int32),
    "ros_2000": np.array([6] * 1999, np.int32),
    "n32x592": np.array([32] * 592, np.int32),
    "n64x592": np.array([64] * 592, np.int32),
    "n128x148": np.array([128] * 148, np.int32),
    "n168x148": np.array([168] * 148, np.int32),
    "n16x4000": np.array([16] * 4000, np.int32),
    "n10x6000": np.array([10] * 6000, np.int32),
}
for name, blk in workloads.items():
    x = random_svec(blk, seed=0)
    ms, sweeps = time_plan(blk, x)
    rec = {"ours_ms": ms, "sweeps_mean": float(sweeps.mean()), "sweeps_max": int(sweeps.max()),
           "F_alg": float(sum((20 / 3) * float(n) ** 3 for n in blk)), "vec_len": int(len(x))}
    if ref is not None and name in ("C2b_2000xU6_60", "planarhand_n1", "pendulum_N80", "ros_2000"):
        rms, o = time_ref(blk, x, reps=3)
        rec["refA_ms"] = rms
        lap = onp.project_svec(blk, x)
        rec["refA_relerr_vs_lapack"] = float(np.linalg.norm(o - lap) / np.linalg.norm(lap))
        ours = cu.Plan(blk).project_host(x)
        rec["ours_relerr_vs_lapack"] = float(np.linalg.norm(ours - lap) / np.linalg.norm(lap))
    out[name] = rec
    print(name, rec, flush=True)

json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
print(json.dumps(out))
