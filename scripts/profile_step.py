"""Profiling driver: init the bench workload, then run a few sGS-ADMM iterations inside a
cudaProfilerStart/Stop window (use with `ncu --profile-from-start off`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cuadmm_b200 as cu
from cuadmm_b200.synthetic import c2b_blocks, chain_sdp
nblk = int(os.environ.get("NBLK", "2000")); con = int(os.environ.get("CON", "700000")); iters = int(os.environ.get("ITERS", "2"))
P = chain_sdp(c2b_blocks(nblk), con, seed=0)
s = cu.Solver(verbose=False)
s.init(15, 30, P["vec_len"], P["con_num"], P["col_ptrs"], P["row_ids"], P["vals"], P["b_idx"], P["b_val"], P["C_idx"], P["C_val"], P["blk"])
s.solve(1, -1.0, 500, 50, 100, 1 << 30, 1.05)
s.run_iterations(5, sgs=True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
r = s.run_iterations(iters, sgs=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", iters, "iterations", r)
