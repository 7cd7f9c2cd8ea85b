"""Regenerates At (not shipped as At.txt) of bundled reference examples from the bundled .mat files and packs
them as tests/golden/<name>.npz next to the reference's committed logs (run in the build container, where
/root/reference is mounted; the GPU box never reads /root/reference).

  PlanarHand_N=1_MOMENT, PushBox_N=50_MOMENT : examples/SPOT/data/MOSEK/<name>.mat, MOSEK `prob` struct:
      bara.{subi,subj,subk,subl,val} = lower-triangular triplets (k >= l) of constraint matrix subi in block subj.
      svec row = off[j] + c(c+1)/2 + r with r = min(k,l)-1, c = max(k,l)-1 (upper triangle by columns),
      off-diagonal values * sqrt(2) (examples/sedumi_to_txt.m:41-49 via SDPT3's svec), b = blc.
      b.txt / C.txt / blk.txt / con_num.txt ship under examples/SPOT/data/TXT/<name>/ and are used to CHECK the
      conversion (C regenerated from barc must equal the shipped C.txt).
  pendulum N=80_licols : examples/pendulum/MATLAB/N=80_licols.mat, SDP.sdpt3.At = per-block svec matrices, stacked
      (examples/sedumi_to_txt.m:41-43 from_cell_to_array).
"""
import os, shutil, sys
import numpy as np
import scipy.io as sio
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
SQRT2 = 1.414213562373095   # the reference's SQRT2; the MATLAB side used sqrt(2): one ulp apart, below the 16 printed digits' effect


def read_sparse_txt(p):
    try:
        a = np.loadtxt(p, ndmin=2)
    except Exception:
        a = np.zeros((0, 3))
    return a.reshape(-1, 3) if a.size else np.zeros((0, 3))


def svec_rows(off, subj, subk, subl):
    r = np.minimum(subk, subl) - 1
    c = np.maximum(subk, subl) - 1
    return off[subj - 1] + c * (c + 1) // 2 + r


def from_mosek(mat, txt_dir, name):
    prob = sio.loadmat(mat, squeeze_me=True, struct_as_record=False)["prob"]
    blk = np.asarray(prob.bardim, np.int64)
    shipped_blk = np.array([int(l.split()[-1]) for l in open(txt_dir + "/blk.txt") if l.strip()], np.int64)
    assert np.array_equal(blk, shipped_blk), "bardim differs from the shipped blk.txt"
    off = np.concatenate([[0], np.cumsum(blk * (blk + 1) // 2)])
    m = int(float(open(txt_dir + "/con_num.txt").read().split()[0]))
    a = prob.bara
    subi, subj, subk, subl = [np.asarray(getattr(a, f), np.int64) for f in ("subi", "subj", "subk", "subl")]
    val = np.asarray(a.val, np.float64)
    rows = svec_rows(off, subj, subk, subl)
    v = np.where(subk == subl, val, val * np.sqrt(2.0))
    At = sp.coo_matrix((v, (rows, subi - 1)), shape=(int(off[-1]), m)).tocsc()
    At.sum_duplicates(); At.sort_indices()
    # check the recipe on C: barc -> svec must reproduce the shipped C.txt
    c = prob.barc
    cj, ck, cl = [np.asarray(getattr(c, f), np.int64) for f in ("subj", "subk", "subl")]
    cv = np.asarray(c.val, np.float64)
    Cvec = np.zeros(int(off[-1]))
    np.add.at(Cvec, svec_rows(off, cj, ck, cl), np.where(ck == cl, cv, cv * np.sqrt(2.0)))
    Cs = read_sparse_txt(txt_dir + "/C.txt")
    Cship = np.zeros(int(off[-1])); Cship[Cs[:, 0].astype(np.int64)] = Cs[:, 2]
    assert np.allclose(Cvec, Cship, rtol=1e-14, atol=1e-15), "svec recipe does not reproduce the shipped C.txt"
    bs = read_sparse_txt(txt_dir + "/b.txt")
    bvec = np.zeros(m); bvec[bs[:, 0].astype(np.int64)] = bs[:, 2]
    assert np.allclose(np.asarray(prob.blc, float), bvec) and np.allclose(np.asarray(prob.buc, float), bvec), "blc/buc differ from b.txt"
    save(name, blk, m, At, bs, Cs)


def from_sdpt3(mat, txt_dir, name):
    S = sio.loadmat(mat, squeeze_me=True, struct_as_record=False)["SDP"].sdpt3
    blk = np.array([int(l.split()[-1]) for l in open(txt_dir + "/blk.txt") if l.strip()], np.int64)
    parts = [sp.csc_matrix(x) for x in np.atleast_1d(S.At)]
    At = sp.vstack(parts).tocsc()
    m = int(float(open(txt_dir + "/con_num.txt").read().split()[0]))
    assert At.shape == (int((blk * (blk + 1) // 2).sum()), m), At.shape
    At.sum_duplicates(); At.sort_indices()
    save(name, blk, m, At, read_sparse_txt(txt_dir + "/b.txt"), read_sparse_txt(txt_dir + "/C.txt"))


def save(name, blk, m, At, b, C):
    coo = At.tocoo()
    # values as %.16f text would carry them (examples/sedumi_to_txt.m:69-72): round-trip through that format
    vals = np.array([float("%.16f" % v) for v in coo.data])
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, blk=blk.astype(np.int32), con_num=m, At_row=coo.row.astype(np.int32), At_col=coo.col.astype(np.int32),
                        At_val=vals, b_idx=b[:, 0].astype(np.int32), b_val=b[:, 2], C_idx=C[:, 0].astype(np.int32), C_val=C[:, 2])
    print(name, "blocks", len(blk), "m", m, "nnz(At)", At.nnz, "bytes", os.path.getsize(path))


if __name__ == "__main__":
    from_mosek(REF + "/examples/SPOT/data/MOSEK/PlanarHand_N=1_MOMENT.mat", REF + "/examples/SPOT/data/TXT/PlanarHand_N=1_MOMENT", "planarhand_n1")
    from_mosek(REF + "/examples/SPOT/data/MOSEK/PushBox_N=50_MOMENT.mat", REF + "/examples/SPOT/data/TXT/PushBox_N=50_MOMENT", "pushbox_n50")
    from_sdpt3(REF + "/examples/pendulum/MATLAB/N=80_licols.mat", REF + "/examples/pendulum/TXT/N=80_licols", "pendulum_n80")
    for src, dst in [("examples/benchmarks/PlanarHand_N=1_MOMENT/sGS-cuADMM.log", "planarhand_n1_sgs.log"),
                     ("examples/benchmarks/PlanarHand_N=1_MOMENT/cuADMM.log", "planarhand_n1_admm.log")]:
        shutil.copy(os.path.join(REF, src), os.path.join(OUT, dst))
    # pendulum log: 100,000 iterations; keep the header and every printed row up to iteration 3000
    lines = open(os.path.join(REF, "examples/pendulum/N=80_licols.log")).read().splitlines()
    keep = []
    for l in lines:
        parts = l.split("|")
        if len(parts) > 3 and parts[0].strip().isdigit() and int(parts[0]) > 3000:
            break
        keep.append(l)
    open(os.path.join(OUT, "pendulum_n80.log"), "w").write("\n".join(keep) + "\n")
