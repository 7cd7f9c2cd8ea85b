"""Shared helpers for the GPU tests: load committed fixtures, build CSC input, parse reference logs."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def load_fixture(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    blk = z["blk"].astype(np.int32)
    con_num = int(z["con_num"])
    vec_len = int(sum(int(n) * (int(n) + 1) // 2 for n in blk))
    rows, cols, vals = z["At_row"].astype(np.int64), z["At_col"].astype(np.int64), z["At_val"].astype(np.float64)
    order = np.lexsort((rows, cols))                 # COO_to_CSC: sort by (col,row)
    rows, cols, vals = rows[order], cols[order], vals[order]
    col_ptrs = np.zeros(con_num + 1, np.int64)
    np.add.at(col_ptrs, cols + 1, 1)
    col_ptrs = np.cumsum(col_ptrs).astype(np.int32)
    return dict(blk=blk, vec_len=vec_len, con_num=con_num, col_ptrs=col_ptrs, row_ids=rows.astype(np.int32), vals=vals,
                b_idx=z["b_idx"].astype(np.int32), b_val=z["b_val"].astype(np.float64),
                C_idx=z["C_idx"].astype(np.int32), C_val=z["C_val"].astype(np.float64))


def write_txt(P, prefix):
    """the SDPT3-style TXT files cuadmm_exe / Problem::from_txt read"""
    os.makedirs(prefix, exist_ok=True)
    with open(os.path.join(prefix, "blk.txt"), "w") as f:
        for n in P["blk"]:
            f.write(f"s {int(n)}\n")
    open(os.path.join(prefix, "con_num.txt"), "w").write(f"{P['con_num']}\n")
    cols = np.repeat(np.arange(P["con_num"]), np.diff(P["col_ptrs"]))
    with open(os.path.join(prefix, "At.txt"), "w") as f:
        for r, c, v in zip(P["row_ids"], cols, P["vals"]):
            f.write(f"{int(r)} {int(c)} {v:.17g}\n")
    with open(os.path.join(prefix, "b.txt"), "w") as f:
        for i, v in zip(P["b_idx"], P["b_val"]):
            f.write(f"{int(i)} 0 {v:.17g}\n")
    with open(os.path.join(prefix, "C.txt"), "w") as f:
        for i, v in zip(P["C_idx"], P["C_val"]):
            f.write(f"{int(i)} 0 {v:.17g}\n")


_LINE = re.compile(r"^\s*(\d+) \| (\S+) (\S+) \|\s+(\S+)\s+(\S+) (\S+) \|\s+(\S+) \| (\S+) \|")


def parse_log(path_or_text):
    text = open(path_or_text).read() if os.path.exists(path_or_text) else path_or_text
    rows = []
    for line in text.splitlines():
        m = _LINE.match(line)
        if m:
            rows.append(dict(it=int(m.group(1)), errRp=float(m.group(2)), errRd=float(m.group(3)), pobj=float(m.group(4)),
                             dobj=float(m.group(5)), relgap=float(m.group(6)), sig=float(m.group(8))))
    return rows


def synthetic_sdp(blk, m, seed=0, nnz_per_con=3, hub=False):
    """Feasible-by-construction SDP on the given blocks: constraints with a few random svec entries in
    at most two blocks, b = A svec(X*) for a random PSD X*, C = svec(S*) + A^T y* with S* complementary
    to X* (SURVEY 8d, config C4 recipe)."""
    import oracle_np as onp
    rng = np.random.default_rng(seed)
    blk = np.asarray(blk, np.int32)
    off = onp.svec_offsets(blk)
    vec_len = int(off[-1])
    xs, ss = [], []
    for n in blk:
        n = int(n)
        Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        r = max(1, n // 3)
        lam = np.zeros(n); lam[:r] = rng.uniform(0.5, 2.0, r)
        mu = np.zeros(n); mu[r:] = rng.uniform(0.5, 2.0, n - r)
        xs.append(onp.svec((Q * lam) @ Q.T)); ss.append(onp.svec((Q * mu) @ Q.T))
    xstar, sstar = np.concatenate(xs), np.concatenate(ss)
    rows, cols, vals = [], [], []
    for i in range(m):
        k = 1 + rng.poisson(nnz_per_con - 1)
        b1, b2 = rng.integers(0, len(blk), 2)
        idx = set()
        for _ in range(k):
            bsel = b1 if rng.random() < 0.5 else b2
            idx.add(int(rng.integers(off[bsel], off[bsel + 1])))
        if hub and i % 7 == 0:
            idx.add(0)
        for j in sorted(idx):
            rows.append(j); cols.append(i); vals.append(rng.standard_normal())
    rows, cols, vals = np.array(rows), np.array(cols), np.array(vals)
    import scipy.sparse as sp
    A = sp.csr_matrix((vals, (cols, rows)), shape=(m, vec_len))
    b = A @ xstar
    ystar = rng.standard_normal(m)
    C = sstar + A.T @ ystar
    Acsr = A.tocsr(); Acsr.sort_indices()
    nzb = np.nonzero(b)[0]; nzc = np.nonzero(C)[0]
    return dict(blk=blk, vec_len=vec_len, con_num=m, col_ptrs=Acsr.indptr.astype(np.int32),
                row_ids=Acsr.indices.astype(np.int32), vals=Acsr.data.astype(np.float64),
                b_idx=nzb.astype(np.int32), b_val=b[nzb], C_idx=nzc.astype(np.int32), C_val=C[nzc],
                xstar=xstar, pstar=float(C @ xstar))


from cuadmm_b200.synthetic import chain_sdp, maxcut_sdp, c2b_blocks  # noqa: E402,F401


def make_solver(P, verbose=False, sig=1.0, X=None, y=None, S=None):
    import cuadmm_b200 as cu
    s = cu.Solver(verbose=verbose)
    s.init(15, 30, P["vec_len"], P["con_num"], P["col_ptrs"], P["row_ids"], P["vals"], P["b_idx"], P["b_val"],
           P["C_idx"], P["C_val"], P["blk"], X, y, S, sig)
    return s
