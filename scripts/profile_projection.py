"""Profiling driver for the projection stage alone: C2b block mix, warm-started plan, drifting input; the projection
inside the cudaProfilerStart/Stop window is the (WARM+1)-th of the sequence (use with ncu --profile-from-start off)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cuadmm_b200 as cu
from cuadmm_b200.synthetic import c2b_blocks, random_svec
blk = c2b_blocks(2000, 6, 60, 0)[:int(os.environ.get("NB", "2000"))]
drift = float(os.environ.get("DRIFT", "1e-4")); warm = int(os.environ.get("WARM", "3"))
x = random_svec(blk, seed=0); d = random_svec(blk, seed=1)
p = cu.Plan(blk, device=0)
dy = torch.empty(len(x), dtype=torch.float64, device="cuda")
for t in range(warm + 1):
    dx = torch.from_numpy(x + t * drift * d).cuda()
    torch.cuda.synchronize()
    if t == warm:
        torch.cuda.profiler.start()
    p.project_device(dx.data_ptr(), dy.data_ptr())
    torch.cuda.synchronize()
    if t == warm:
        torch.cuda.profiler.stop()
print("ok", p.last_ms)
