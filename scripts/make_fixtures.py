"""Builds the committed fixtures under tests/golden/ from the reference tree (run in the build
container, where /root/reference is mounted; the GPU box never sees /root/reference).
  * <name>.npz : blk, con_num, At triplets (as shipped in At.txt), b, C of bundled example problems
  * <name>_<mode>.log : the reference's own committed solver logs (iteration-indexed trajectories)
"""
import os, shutil, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

def pack(src, name):
    blk = [int(l.split()[-1]) for l in open(src + "/blk.txt") if l.strip()]
    con = int(float(open(src + "/con_num.txt").read().split()[0]))
    At = np.loadtxt(src + "/At.txt", ndmin=2)
    def sv(p):
        try: a = np.loadtxt(p, ndmin=2)
        except Exception: a = np.zeros((0, 3))
        return a.reshape(-1, 3) if a.size else np.zeros((0, 3))
    b, C = sv(src + "/b.txt"), sv(src + "/C.txt")
    np.savez_compressed(os.path.join(OUT, name + ".npz"), blk=np.array(blk, np.int32), con_num=con,
                        At_row=At[:, 0].astype(np.int32), At_col=At[:, 1].astype(np.int32), At_val=At[:, 2],
                        b_idx=b[:, 0].astype(np.int32), b_val=b[:, 2], C_idx=C[:, 0].astype(np.int32), C_val=C[:, 2])
    print(name, len(blk), con, len(At), os.path.getsize(os.path.join(OUT, name + ".npz")))

pack(REF + "/examples/SPOT/data/TXT/PushT_N=10_MOMENT", "pusht_n10")
pack(REF + "/examples/plato/TXT/ros_2000", "ros_2000")
pack(REF + "/examples/plato/TXT/rose13", "rose13")
pack(REF + "/examples/dimacs/data/TXT/truss8", "truss8")
pack(REF + "/examples/plato/TXT/biggs", "biggs")
pack(REF + "/examples/dimacs/data/TXT/hinf12", "hinf12")
for src, dst in [("examples/benchmarks/PushT_N=10_MOMENT/sGS-cuADMM.log", "pusht_n10_sgs.log"),
                 ("examples/benchmarks/PushT_N=10_MOMENT/cuADMM.log", "pusht_n10_admm.log"),
                 ("examples/benchmarks/ros_2000/sGS-cuADMM.log", "ros_2000_sgs.log"),
                 ("examples/benchmarks/ros_2000/cuADMM.log", "ros_2000_admm.log"),
                 ("examples/plato/logs/rose13.log", "rose13.log")]:
    shutil.copy(os.path.join(REF, src), os.path.join(OUT, dst))
