import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from util_problems import load_fixture, make_solver, parse_log
name = sys.argv[1]; max_iter = int(sys.argv[2]); tol = float(sys.argv[3]); switch = int(sys.argv[4])
P = load_fixture(name)
t = time.time(); s = make_solver(P, verbose=True); print("init s", time.time() - t, flush=True)
if s.h and hasattr(s, "times"): pass
t = time.time(); s.solve(max_iter, tol, 0, 50, 100, switch); dt = time.time() - t
print("iters", s.info_iter_num, "solve s", dt, "ms/iter", 1e3 * dt / max(s.info_iter_num, 1), "launches", s.launches)
