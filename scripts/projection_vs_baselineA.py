"""PSD-projection stage (ms) on the block sets of the bundled / bench workloads: this repo (cold plan, and warm-started
on a drifting sequence as inside the solver) next to Baseline A = the reference's own cuSOLVER stage
(src/solver.cu:531-647, unmodified reference sources in oracle/_ref), same GPU, same input."""
import ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import cuadmm_b200 as cu
from cuadmm_b200.synthetic import c2b_blocks, c4_blocks, random_svec
from util_problems import load_fixture
ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libcuadmm_ref.so"))
ref.ref_proj_create.restype = C.c_void_p; ref.ref_proj_run.restype = C.c_double
sets = {"planarhand_n1 (C1)": load_fixture("planarhand_n1")["blk"], "pushbox_n50 (C2a)": load_fixture("pushbox_n50")["blk"],
        "pendulum_n80 (C2c)": load_fixture("pendulum_n80")["blk"], "c2b": c2b_blocks(), "pusht_n10": load_fixture("pusht_n10")["blk"]}
out = []
for name, blk in sets.items():
    blk = np.ascontiguousarray(blk, np.int32)
    x = random_svec(blk, seed=0); d = random_svec(blk, seed=1)
    h = ref.ref_proj_create(blk.ctypes.data_as(C.POINTER(C.c_int)), len(blk), 15)
    msA = ref.ref_proj_run(C.c_void_p(h), x.ctypes.data_as(C.POINTER(C.c_double)), None, 3)
    ref.ref_proj_destroy(C.c_void_p(h))
    p = cu.Plan(blk, device=0)
    ms = []
    for t in range(8):
        dx = torch.from_numpy(x + t * 1e-4 * d).cuda(); dy = torch.empty_like(dx)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); p.project_device(dx.data_ptr(), dy.data_ptr()); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    row = {"blocks": name, "nblk": int(len(blk)), "n_min": int(blk.min()), "n_max": int(blk.max()), "baseline_A_ms": msA,
           "ours_cold_ms": ms[0], "ours_warm_ms": float(np.mean(ms[2:])), "ratio_cold": msA / ms[0], "ratio_warm": msA / float(np.mean(ms[2:]))}
    print(json.dumps(row), flush=True); out.append(row)
    p.close()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "projection_vs_baselineA.json"), "w"), indent=1)
