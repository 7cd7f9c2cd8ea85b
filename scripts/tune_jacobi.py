"""Tuning sweep of the Jacobi kernel variants (CUADMM_JACOBI_CLASSES override): for each block size
time every feasible variant on a batch that fills the GPU a few times.  Writes gpurun_out/tune.json."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import cuadmm_b200 as cu
from conftest import random_svec

VARIANTS = {0: (128, 4, 4, True), 1: (64, 4, 8, False), 2: (128, 8, 4, False), 3: (128, 4, 16, False), 4: (256, 8, 8, False),
            5: (512, 16, 4, False), 6: (384, 8, 12, False), 7: (768, 16, 6, False), 8: (512, 8, 16, False),
            9: (1024, 16, 8, False), 10: (672, 16, 11, False), 11: (192, 4, 24, False), 12: (256, 4, 32, False),
            13: (352, 8, 21, False), 14: (256, 8, 4, True), 15: (128, 8, 8, False)}
res = {}
for n in [6, 10, 16, 20, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 150, 168]:
    count = max(148, min(6000, int(148 * 4 * (64 / n) ** 2)))
    blk = np.full(count, n, np.int32)
    x = random_svec(blk[:8], seed=n)
    x = np.tile(x, count // 8 + 1)[: count * n * (n + 1) // 2]
    dx = torch.from_numpy(x).cuda(); dy = torch.empty_like(dx)
    for v, (T, L, RPL, W) in VARIANTS.items():
        cap = min(L * RPL, 168)
        if n > cap or (W and n > 32):
            continue
        if cap > 4 * max(n, 16):
            continue
        os.environ["CUADMM_JACOBI_CLASSES"] = f"{cap}:{v}"
        try:
            p = cu.Plan(blk, device=0)
            for _ in range(2): p.project_device(dx.data_ptr(), dy.data_ptr())
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): p.project_device(dx.data_ptr(), dy.data_ptr())
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            res.setdefault(str(n), {})[str(v)] = ms / count * 1e3   # us per block (throughput)
            print(f"n={n} count={count} variant {v} {VARIANTS[v]}: {ms:.3f} ms  {ms/count*1e3:.3f} us/block", flush=True)
            p.close()
        except Exception as ex:
            print("fail", n, v, ex)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "tune.json"), "w"), indent=1)
