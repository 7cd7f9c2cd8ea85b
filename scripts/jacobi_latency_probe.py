"""Projection time of a SMALL batch of blocks (one rank's share of C2b on 8 GPUs: 250 blocks) for several Jacobi
size-class tables — in that regime the stage is bound by the latency of one block, not by throughput."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cuadmm_b200 as cu
from cuadmm_b200.synthetic import c2b_blocks, random_svec
nb = int(os.environ.get("NB", "250"))
blk = c2b_blocks(2000, 6, 60, 0)[:nb]
x = random_svec(blk, seed=0); d = random_svec(blk, seed=1)
for table in sys.argv[1:]:
    if table != "default":
        os.environ["CUADMM_JACOBI_CLASSES"] = table
    else:
        os.environ.pop("CUADMM_JACOBI_CLASSES", None)
    res = {}
    for drift in (1e-2, 1e-4, 0.0):
        p = cu.Plan(blk, device=0)
        ms = []
        for t in range(6):
            dx = torch.from_numpy(x + t * drift * d).cuda(); dy = torch.empty_like(dx)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(); p.project_device(dx.data_ptr(), dy.data_ptr()); e1.record(); torch.cuda.synchronize()
            if t >= 1:
                ms.append(e0.elapsed_time(e1))
        res[str(drift)] = round(float(np.mean(ms)), 4)
        p.close()
    print(json.dumps({"nb": nb, "table": table, "ms": res}), flush=True)
