"""Synthetic SDP instances in the reference's input format (numpy only; data generation, no compute
path).  Used by bench.py, smoke() and the tests for the BASELINE.json configurations that have no
bundled data (C2 "~2,000 PSD blocks of size 6-60", C3 max-cut, C4 10k mixed blocks)."""
import numpy as np

_SQRT2 = 1.414213562373095   # the reference's SQRT2 (include/cuadmm/kernels.h:173-181)


def svec_offsets(blk):
    blk = np.asarray(blk, dtype=np.int64)
    return np.concatenate([[0], np.cumsum(blk * (blk + 1) // 2)])


def _svec(M):
    n = M.shape[0]
    cols, rows = np.tril_indices(n)
    v = M[rows, cols]
    return np.where(rows == cols, v, v * _SQRT2)


def c2b_blocks(nblk=2000, lo=6, hi=60, seed=0):
    """BASELINE.json configs[1]: nblk blocks, n_k ~ UniformInt[lo, hi], numpy default_rng(seed)"""
    return np.random.default_rng(seed).integers(lo, hi + 1, nblk).astype(np.int32)


def c4_blocks(seed=0, counts=((10, 6000), (50, 3000), (200, 900), (800, 100))):
    """BASELINE.json configs[3]: 10,000 blocks of mixed sizes in shuffled order"""
    blk = np.concatenate([np.full(c, n, np.int32) for n, c in counts])
    np.random.default_rng(seed).shuffle(blk)
    return blk


def c5_blocks(seed=0, n_big=8000, k_big=4, k_small=500, lo=6, hi=32):
    """BASELINE.json configs[4]: 4 PSD blocks of n=8,000 plus 500 small blocks n ~ UniformInt[6, 32]"""
    small = np.random.default_rng(seed).integers(lo, hi + 1, k_small).astype(np.int32)
    return np.concatenate([np.full(k_big, n_big, np.int32), small])


def _complementary_pair(n, rng, thin):
    """svec(X*), svec(S*) with X* S* = 0, both PSD.  thin: rank-32 factors from one thin QR (O(n^2 r) instead of O(n^3))"""
    if thin and n > 128:
        r = 32
        Q, _ = np.linalg.qr(rng.standard_normal((n, 2 * r)))
        V, W = Q[:, :r], Q[:, r:]
        return _svec((V * rng.uniform(0.5, 2.0, r)) @ V.T), _svec((W * rng.uniform(0.5, 2.0, r)) @ W.T)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    r = max(1, n // 3)
    lam = np.zeros(n); lam[:r] = rng.uniform(0.5, 2.0, r)
    mu = np.zeros(n); mu[r:] = rng.uniform(0.5, 2.0, n - r)
    return _svec((Q * lam) @ Q.T), _svec((Q * mu) @ Q.T)


def random_svec(blk, seed=0):
    """random symmetric blocks (G + G^T)/2, G ~ N(0,1), eigenvalues straddling 0"""
    rng = np.random.default_rng(seed)
    parts = []
    for n in blk:
        G = rng.standard_normal((int(n), int(n)))
        parts.append(_svec((G + G.T) / 2))
    return np.concatenate(parts) if parts else np.zeros(0)


def chain_sdp(blk, m, seed=0, extra_frac=0.1, coeffs="gauss"):
    """Moment-relaxation-like SDP (structure of the SPOT / pendulum examples): the blocks form a chain
    (time steps); most constraints are 2-entry equalities between an svec entry of block j and one of
    block j or j+1, a fraction `extra_frac` has 3-5 entries; every svec entry is used by ~1-2
    constraints, so A A^T factors with little fill, and ~2 % of the constraints are redundant (as in
    the bundled SPOT data).  Feasible by construction: b = A svec(X*), C = svec(S*) + A^T y* with
    X* S* = 0, so the optimal value <C, X*> is known."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    blk = np.asarray(blk, np.int32)
    nb = len(blk)
    off = svec_offsets(blk)
    vec_len = int(off[-1])
    xs, ss = [], []
    for n in blk:
        n = int(n)
        Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        r = max(1, n // 3)
        lam = np.zeros(n); lam[:r] = rng.uniform(0.5, 2.0, r)
        mu = np.zeros(n); mu[r:] = rng.uniform(0.5, 2.0, n - r)
        xs.append(_svec((Q * lam) @ Q.T)); ss.append(_svec((Q * mu) @ Q.T))
    xstar, sstar = np.concatenate(xs), np.concatenate(ss)
    w = np.diff(off).astype(float)
    home = rng.choice(nb, size=m, p=w / w.sum())
    home.sort()
    k = np.where(rng.random(m) < extra_frac, rng.integers(3, 6, m), 2)
    tot = int(k.sum())
    con = np.repeat(np.arange(m), k)
    hb = np.repeat(home, k)
    nxt = np.minimum(hb + (rng.random(tot) < 0.35), nb - 1)
    ent = off[nxt] + (rng.random(tot) * (off[nxt + 1] - off[nxt])).astype(np.int64)
    val = rng.standard_normal(tot)
    if coeffs == "unit":
        # moment-consistency style rows (x_i - x_j = b, as in the SPOT data): +-1 coefficients.  A A^T is then a signed
        # graph Laplacian: redundant constraints (cycles) give pivots at rounding level and everything else is
        # polynomially conditioned, whereas N(0,1) coefficients along long chains make A A^T numerically singular with
        # a continuum of pivots between 1e-16 and 1e-11 (two valid factorisations then disagree by ~1e-5 in A^T y)
        first = np.concatenate([[True], con[1:] != con[:-1]])
        val = np.where(first, 1.0, -1.0)
    A = sp.csr_matrix((val, (con, ent)), shape=(m, vec_len))
    A.sum_duplicates(); A.sort_indices()
    b = A @ xstar
    ystar = rng.standard_normal(m)
    C = sstar + A.T @ ystar
    nzb = np.nonzero(b)[0]; nzc = np.nonzero(C)[0]
    return dict(blk=blk, vec_len=vec_len, con_num=m, col_ptrs=A.indptr.astype(np.int32),
                row_ids=A.indices.astype(np.int32), vals=A.data.astype(np.float64),
                b_idx=nzb.astype(np.int32), b_val=b[nzb], C_idx=nzc.astype(np.int32), C_val=C[nzc],
                xstar=xstar, pstar=float(C @ xstar))


def random_sdp(blk, m, seed=0, mean_extra=4.0, cheap_factors=False):
    """BASELINE.json configs[3] as defined in SURVEY 8d: m constraints, each with k ~ 1 + Poisson(mean_extra)
    non-zeros at uniformly random svec positions of at most 2 blocks (the two blocks drawn with probability
    proportional to their svec length, i.e. positions are uniform over the svec vector), values N(0,1);
    b = A svec(X*) for a random PSD X*, C = svec(S*) + A^T y* with X* S* = 0 (optimum known)."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    blk = np.asarray(blk, np.int32)
    nb = len(blk)
    off = svec_offsets(blk)
    vec_len = int(off[-1])
    xs, ss = [], []
    for n in blk:
        if cheap_factors:
            xk, sk = _complementary_pair(int(n), rng, True)
            xs.append(xk); ss.append(sk)
            continue
        n = int(n)
        Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        r = max(1, n // 3)
        lam = np.zeros(n); lam[:r] = rng.uniform(0.5, 2.0, r)
        mu = np.zeros(n); mu[r:] = rng.uniform(0.5, 2.0, n - r)
        xs.append(_svec((Q * lam) @ Q.T)); ss.append(_svec((Q * mu) @ Q.T))
    xstar, sstar = np.concatenate(xs), np.concatenate(ss)
    w = np.diff(off).astype(float)
    two = rng.choice(nb, size=(m, 2), p=w / w.sum())
    k = 1 + rng.poisson(mean_extra, m)
    tot = int(k.sum())
    con = np.repeat(np.arange(m), k)
    bsel = two[con, (rng.random(tot) < 0.5).astype(np.int64)]
    ent = off[bsel] + (rng.random(tot) * (off[bsel + 1] - off[bsel])).astype(np.int64)
    val = rng.standard_normal(tot)
    A = sp.csr_matrix((val, (con, ent)), shape=(m, vec_len))
    A.sum_duplicates(); A.sort_indices()
    b = A @ xstar
    ystar = rng.standard_normal(m)
    C = sstar + A.T @ ystar
    nzb = np.nonzero(b)[0]; nzc = np.nonzero(C)[0]
    return dict(blk=blk, vec_len=vec_len, con_num=m, col_ptrs=A.indptr.astype(np.int32),
                row_ids=A.indices.astype(np.int32), vals=A.data.astype(np.float64),
                b_idx=nzb.astype(np.int32), b_val=b[nzb], C_idx=nzc.astype(np.int32), C_val=C[nzc],
                xstar=xstar, pstar=float(C @ xstar))


def maxcut_sdp(n, p=0.01, seed=0):
    """BASELINE.json configs[2]: Max-Cut SDP on G(n, p), unit weights, C = -1/4 Laplacian, diag(X) = 1
    (examples/max-cut/genMAXCUT.m:20-31).  One block, m = n constraints, A A^T = I."""
    rng = np.random.default_rng(seed)
    iu = np.triu_indices(n, 1)
    mask = rng.random(len(iu[0])) < p
    r, c = iu[0][mask], iu[1][mask]
    deg = np.bincount(np.concatenate([r, c]), minlength=n).astype(float)
    idx = np.concatenate([c * (c + 1) // 2 + r, np.arange(n) * (np.arange(n) + 1) // 2 + np.arange(n)])
    val = np.concatenate([np.full(len(r), 0.25 * _SQRT2), -0.25 * deg])
    order = np.argsort(idx)
    diag = (np.arange(n) * (np.arange(n) + 1) // 2 + np.arange(n)).astype(np.int32)
    return dict(blk=np.array([n], np.int32), vec_len=n * (n + 1) // 2, con_num=n,
                col_ptrs=np.arange(n + 1, dtype=np.int32), row_ids=diag, vals=np.ones(n),
                b_idx=np.arange(n, dtype=np.int32), b_val=np.ones(n),
                C_idx=idx[order].astype(np.int32), C_val=val[order])
