"""oracle_np.py — TEST INFRASTRUCTURE ONLY (never imported by the product).

numpy/scipy restatement of the floating-point half of the reference's hot path, each function
citing the reference lines it follows.  Integer logic lives in oracle_host.c.  Pinned by
tests/test_oracle_golden.py against the reference's own known-answer tests and golden vectors.

Third-party pieces the reference links but does not vendor:
  * LAPACK dsyevd (MATLAB's lapack.h, include/cuadmm/eig_cpu.h:12,42; no pinned version):
    here OpenBLAS 0.3.30 dsyevd through scipy.linalg.lapack.dsyevd.
  * SuiteSparse CHOLMOD (hard-wired /usr/local/lib/libcholmod.so, CMakeLists.txt:42-44; no pinned
    version): restated as "solve (A A^T + eps I) y = rhs to working precision" with a sparse LU
    (scipy splu) — labelled CHOLMOD-substitute wherever it is timed.
"""
import math
import threading
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from scipy.linalg import lapack

def _newton_sqrt2():
    """include/cuadmm/kernels.h:173-181: fixed point of x <- (x + 2/x)/2 from 2.0.  It is
    1.414213562373095, one ulp below math.sqrt(2) (the reference's own test only checks 4 ulp)."""
    prev, curr = 0.0, 2.0
    while curr != prev:
        prev, curr = curr, 0.5 * (curr + 2.0 / curr)
    return curr


SQRT2 = _newton_sqrt2()
SQRT2INV = 1.0 / SQRT2


def tri(n):
    return n * (n + 1) // 2


def svec_offsets(blk):
    blk = np.asarray(blk, dtype=np.int64)
    return np.concatenate([[0], np.cumsum(blk * (blk + 1) // 2)])


def smat(v, n):
    """svec (upper triangle by columns, off-diagonals * sqrt2) -> dense symmetric.
    Same arithmetic as vector_to_matrices_kernel (src/kernels/vec_mat_conversion.cu:11-34)."""
    M = np.zeros((n, n))
    iu = np.triu_indices(n)
    # column-major upper triangle order: for column c, rows 0..c
    cols, rows = np.tril_indices(n)  # (c, r) with r <= c enumerated c-major
    vals = np.where(rows == cols, v, v * SQRT2INV)
    M[rows, cols] = vals
    M[cols, rows] = vals
    return M


def svec(M):
    """dense symmetric -> svec; matrices_to_vector_kernel (vec_mat_conversion.cu:36-57)."""
    n = M.shape[0]
    cols, rows = np.tril_indices(n)
    v = M[rows, cols]
    return np.where(rows == cols, v, v * SQRT2)


def eig_dsyevd(M):
    """single_eig_lapack (include/cuadmm/eig_cpu.h:31-51): dsyevd('V','U'), ascending W,
    eigenvectors in the columns."""
    w, v, info = lapack.dsyevd(np.asfortranarray(M), compute_v=1, lower=0)
    if info != 0:
        raise RuntimeError(f"dsyevd info={info}")
    return w, v


def project_block(M):
    """clamp (src/kernels/dense_scalar.cu:41-47) + column scale (diagonal_batch.cu:11-22) +
    Q diag(w+) Q^T (include/cuadmm/cublas.h:18-35)."""
    w, Q = eig_dsyevd(M)
    wp = np.maximum(w, 0.0)
    return (Q * wp) @ Q.T, w


def project_svec(blk, Xb, want_eig=False):
    """The whole projection stage src/solver.cu:531-647 with LAPACK eig on every block."""
    off = svec_offsets(blk)
    out = np.empty_like(Xb)
    eigs = []
    for k, n in enumerate(blk):
        P, w = project_block(smat(Xb[off[k]:off[k + 1]], int(n)))
        out[off[k]:off[k + 1]] = svec(P)
        eigs.append(w)
    if want_eig:
        return out, (np.concatenate(eigs) if eigs else np.zeros(0))
    return out


def project_svec_rank(blk, Xb, eig_rank):
    """Fixed-rank projection: the stage src/solver.cu:531-647 with max_dense_vector_zero replaced by
    max_dense_vector_zero_mask (src/kernels/dense_scalar.cu:51-56) and the mask of get_eig_rank_mask
    (src/utils/get_eig_rank_mask.cu:16-38): W <- max(W, 0) * mask, mask = 1 on the eig_rank largest eigenvalues
    (the wiring is present but commented out in src/duo_solver.cu:843-850)."""
    off = svec_offsets(blk)
    out = np.empty_like(Xb)
    for k, n in enumerate(blk):
        n = int(n)
        w, Q = eig_dsyevd(smat(Xb[off[k]:off[k + 1]], n))
        mask = np.zeros(n)
        mask[max(0, n - eig_rank):] = 1.0
        wp = np.maximum(w, 0.0) * mask
        out[off[k]:off[k + 1]] = svec((Q * wp) @ Q.T)
    return out


def thread_ranges(count, nthreads):
    """SDPDuoSolver's split of blocks over CPU eig threads / GPUs (src/duo_solver.cu:270-295,346-371):
    contiguous ranges; every worker gets floor(count/T), the last takes the rest, then (T > 2) the
    last hands one block at a time to workers 0,1,... while it holds >= 2 more than them."""
    T = max(1, nthreads)
    base = count // T
    per = [base] * T
    per[T - 1] = count - (T - 1) * base
    if T > 2:
        i = 0
        while i < T - 1 and per[T - 1] - per[i] >= 2:
            per[i] += 1
            per[T - 1] -= 1
            i += 1
    ranges, start = [], 0
    for c in per:
        ranges.append((start, start + c))
        start += c
    return ranges


def project_svec_threads(blk, Xb, nthreads):
    """Baseline B: the reference's LAPACK thread pool (src/duo_solver.cu:578-619) over all blocks.
    scipy's LAPACK releases the GIL, so Python threads parallelise like the reference's
    std::threads (OPENBLAS_NUM_THREADS=1 keeps each dsyevd single-threaded)."""
    off = svec_offsets(blk)
    out = np.empty_like(Xb)
    nblk = len(blk)

    def work(lo, hi):
        for k in range(lo, hi):
            P, _ = project_block(smat(Xb[off[k]:off[k + 1]], int(blk[k])))
            out[off[k]:off[k + 1]] = svec(P)

    ths = [threading.Thread(target=work, args=r) for r in thread_ranges(nblk, nthreads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return out


_CB = None


def cpu_baseline_lib():
    """oracle/_build/libcpu_baseline.so (cpu_baseline.cpp): the reference's CPU projection path in C++ —
    dsyevd on a std::thread pool (src/duo_solver.cu:346-371, 578-619) — bound to scipy's bundled OpenBLAS."""
    global _CB
    if _CB is None:
        import ctypes as C
        import glob
        import os
        import scipy
        here = os.path.dirname(os.path.abspath(__file__))
        lib = C.CDLL(os.path.join(here, "_build", "libcpu_baseline.so"))
        cands = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so"))
        if not cands or lib.cb_init(os.path.abspath(cands[0]).encode()) != 0:
            raise RuntimeError("cpu_baseline: cannot bind LAPACK dsyevd / BLAS dgemm from scipy's OpenBLAS")
        lib.cb_project.restype = C.c_int
        _CB = lib
    return _CB


def project_svec_cpp(blk, Xb, nthreads, want_eig=False):
    """Baseline B in C++ (cpu_baseline.cpp cb_project): same result as project_svec_threads, without the Python
    per-block overhead — this is what the CPU arm of bench.py times."""
    import ctypes as C
    lib = cpu_baseline_lib()
    blk = np.ascontiguousarray(blk, np.int32)
    Xb = np.ascontiguousarray(Xb, np.float64)
    out = np.empty_like(Xb)
    eig = np.zeros(int(blk.sum())) if want_eig else None
    bad = lib.cb_project(blk.ctypes.data_as(C.POINTER(C.c_int)), len(blk), Xb.ctypes.data_as(C.POINTER(C.c_double)),
                         out.ctypes.data_as(C.POINTER(C.c_double)),
                         eig.ctypes.data_as(C.POINTER(C.c_double)) if want_eig else None, int(nthreads))
    if bad != 0:
        raise RuntimeError(f"cpu_baseline: dsyevd failed on {bad} blocks")
    return (out, eig) if want_eig else out


# ---------------------------------------------------------------------------------------------
# sparse pieces
# ---------------------------------------------------------------------------------------------
def get_normA(col_ptrs, vals):
    """get_normA_kernel (src/kernels/sparse_matrix_norm.cu:11-31): serial sums, floor 1.0,
    in-place DIVISION."""
    vals = np.array(vals, dtype=np.float64)
    m = len(col_ptrs) - 1
    normA = np.zeros(m)
    for i in range(m):
        s = 0.0
        for p in range(col_ptrs[i], col_ptrs[i + 1]):
            s += vals[p] * vals[p]
        nrm = max(1.0, math.sqrt(s))
        normA[i] = nrm
        vals[col_ptrs[i]:col_ptrs[i + 1]] /= nrm
    return normA, vals


def get_normA_fast(col_ptrs, vals):
    """vectorised variant for big inputs (sums by np.add.reduceat — pairwise, so only ~1 ulp from
    the serial sum; used by the ADMM oracle, not by the bit-exact tests)."""
    vals = np.array(vals, dtype=np.float64)
    m = len(col_ptrs) - 1
    cnt = np.diff(col_ptrs)
    s = np.zeros(m)
    nz = cnt > 0
    if len(vals):
        s[nz] = np.add.reduceat(vals * vals, np.asarray(col_ptrs[:-1])[nz])
    normA = np.maximum(1.0, np.sqrt(s))
    vals /= np.repeat(normA, cnt)
    return normA, vals


class AATSolver:
    """CholeskySolverCPU (include/cuadmm/cholesky_cpu.h:62-155) restated: factor A A^T + eps I
    once, solve per call.  CHOLMOD-substitute: SuperLU with symmetric-mode MMD ordering."""

    def __init__(self, A_csr, eps=1e-15):
        m = A_csr.shape[0]
        M = (A_csr @ A_csr.T + eps * sp.eye(m, format="csr")).tocsc()
        self.lu = spla.splu(M, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0,
                            options=dict(SymmetricMode=True))

    def solve(self, rhs):
        return self.lu.solve(rhs)


# ---------------------------------------------------------------------------------------------
# the ADMM iteration, restating SDPSolver::init / solve (src/solver.cu:27-342, 355-822)
# ---------------------------------------------------------------------------------------------
class ADMMOracle:
    def __init__(self, vec_len, con_num, At_col_ptrs, At_row_ids, At_vals, b_idx, b_val, C_idx, C_val, blk,
                 X=None, y=None, S=None, sig=1.0, project=None):
        self.vec_len, self.m = vec_len, con_num
        self.blk = np.asarray(blk, dtype=np.int64)
        normA, vals = get_normA_fast(np.asarray(At_col_ptrs), At_vals)           # solver.cu:79-80
        self.normA = normA
        # At is CSC (vec_len x m) == A in CSR (m x vec_len)
        self.A = sp.csr_matrix((vals, np.asarray(At_row_ids), np.asarray(At_col_ptrs)), shape=(con_num, vec_len))
        self.At = self.A.T.tocsr()
        self.lin = AATSolver(self.A, 1e-15)                                      # solver.cu:91-96
        b = np.zeros(con_num); b[np.asarray(b_idx, dtype=np.int64)] = b_val
        C = np.zeros(vec_len); C[np.asarray(C_idx, dtype=np.int64)] = C_val
        self.X = np.zeros(vec_len) if X is None else np.array(X, dtype=np.float64)
        self.y = np.zeros(con_num) if y is None else np.array(y, dtype=np.float64)
        self.S = np.zeros(vec_len) if S is None else np.array(S, dtype=np.float64)
        self.sig = sig
        self.norm_borg = 1 + np.linalg.norm(b)                                   # solver.cu:177-178
        self.norm_Corg = 1 + np.linalg.norm(C)
        b = b / normA                                                            # :181
        self.y = self.y * normA                                                  # :182
        self.bscale = 1 + np.linalg.norm(b)                                      # :184-186
        self.Cscale = 1 + np.linalg.norm(C)
        self.objscale = self.bscale * self.Cscale
        self.b = b / self.bscale
        self.C = C / self.Cscale
        self.X /= self.bscale; self.S /= self.Cscale; self.y /= self.Cscale      # :189-191
        self.Aty = self.At @ self.y                                              # :205-218
        self.Rp = self.b - self.A @ self.X
        self.SmC = self.S - self.C
        self.Rd = self.Aty + self.SmC
        self._residuals()
        self.prim_win = 0; self.dual_win = 0
        self.sigmax, self.sigmin = 1e3, 1e-3
        self.project = project or (lambda v: project_svec(self.blk, v))
        self.hist = dict(pobj=[], dobj=[], errRp=[], errRd=[], relgap=[], sig=[])

    def _residuals(self):
        self.errRp = np.linalg.norm(self.normA * self.Rp * self.bscale) / self.norm_borg
        self.errRd = np.linalg.norm(self.Rd * self.Cscale) / self.norm_Corg
        self.maxfeas = max(self.errRp, self.errRd)
        self.pobj = float(self.C @ self.X) * self.objscale
        self.dobj = float(self.b @ self.y) * self.objscale
        self.relgap = abs(self.pobj - self.dobj) / (1 + abs(self.pobj) + abs(self.dobj))

    def solve(self, max_iter, stop_tol, sig_update_threshold=500, sig_update_stage_1=50, sig_update_stage_2=100,
              switch_admm=11000, sigscale=1.05):
        it_done = 0
        best = None
        best_KKT = np.inf if switch_admm < 1 else None
        stop_it = max_iter + 1
        for it in range(1, max_iter + 2):
            if max(self.maxfeas, self.relgap) < stop_tol or it > max_iter:       # solver.cu:419-428
                stop_it = it
                # the reference still runs step 1 before breaking (:478-500 precede the break at :567-576)
                rhsy = self.Rp / self.sig - self.A @ self.SmC
                self.y = self.lin.solve(rhsy)
                if it > switch_admm and best is not None:                       # :568-573 restore best
                    self.X, self.y, self.S = best
                break
            rhsy = self.Rp / self.sig - self.A @ self.SmC                       # :478-482
            self.y = self.lin.solve(rhsy)                                       # :487-500
            Rd1 = self.At @ self.y - self.C                                     # :514-521
            Xb = self.X + self.sig * Rd1                                        # :527
            Xproj = self.project(Xb)                                            # :531-647
            self.S = (Xproj - self.X) / self.sig - Rd1                          # :652-656
            self.SmC = self.S - self.C                                          # :672-675
            if it == switch_admm:                                               # :681-690
                sig_update_stage_2 = sig_update_stage_2 // 2
                sigscale = sigscale * 1.23
                best_KKT = max(self.maxfeas, self.relgap)
                best = (self.X.copy(), self.y.copy(), self.S.copy())
            if it < switch_admm:                                                # :693-729
                rhsy = self.Rp / self.sig - self.A @ self.SmC
                self.y = self.lin.solve(rhsy)
                Rd1 = self.At @ self.y - self.C
            if it > switch_admm:                                                # :732-741
                if best_KKT > max(self.maxfeas, self.relgap):
                    best = (self.X.copy(), self.y.copy(), self.S.copy())
                    best_KKT = max(self.maxfeas, self.relgap)
            self.Rd = Rd1 + self.S                                              # :746
            tau = 1.95 if it < switch_admm else 1.618                           # :747-754
            if self.errRd < stop_tol:
                tau = max(1.618, tau / 1.1)
            self.X = self.X + tau * self.sig * self.Rd                          # :757
            self.Rp = self.b - self.A @ self.X                                  # :764-768
            self._residuals()                                                   # :772-779
            feasratio = self.errRp / self.errRd if self.errRd != 0 else np.inf
            if feasratio < 1:
                self.prim_win += 1
            else:
                self.dual_win += 1
            if ((it <= sig_update_threshold and it % sig_update_stage_1 == 1) or
                    (it > sig_update_threshold and it % sig_update_stage_2 == 1)):  # :787-799
                if self.prim_win > 1.2 * self.dual_win:
                    self.prim_win = 0
                    self.sig = min(self.sigmax, self.sig * sigscale)
                elif self.dual_win > 1.2 * self.prim_win:
                    self.dual_win = 0
                    self.sig = max(self.sigmin, self.sig / sigscale)
            for k, v in (("pobj", self.pobj), ("dobj", self.dobj), ("errRp", self.errRp), ("errRd", self.errRd),
                         ("relgap", self.relgap), ("sig", self.sig)):
                self.hist[k].append(v)
            it_done += 1
        # unscale, solver.cu:814-816
        X = self.X * self.bscale
        y = self.y / self.normA * self.Cscale
        S = self.S * self.Cscale
        return X, y, S, it_done


# ---------------------------------------------------------------------------------------------
# TXT problem reader (Problem::from_txt, src/problem.cu:11-83) for tests
# ---------------------------------------------------------------------------------------------
def read_problem_txt(prefix):
    blk = []
    for line in open(prefix + "blk.txt"):
        t = line.split()
        if len(t) == 2 and t[0].isalpha():
            blk.append(int(t[1]))
        elif len(t) == 1:
            blk.append(int(t[0]))
    con_num = int(float(open(prefix + "con_num.txt").read().split()[0]))
    vec_len = int(sum(n * (n + 1) // 2 for n in blk))
    At = np.loadtxt(prefix + "At.txt", ndmin=2)
    rows, cols, vals = At[:, 0].astype(np.int64), At[:, 1].astype(np.int64), At[:, 2]
    order = np.lexsort((rows, cols))                                            # COO_to_CSC
    rows, cols, vals = rows[order], cols[order], vals[order]
    col_ptrs = np.zeros(con_num + 1, dtype=np.int32)
    np.add.at(col_ptrs, cols + 1, 1)
    col_ptrs = np.cumsum(col_ptrs).astype(np.int32)

    def spvec(path):
        try:
            a = np.loadtxt(path, ndmin=2)
        except Exception:
            a = np.zeros((0, 3))
        if a.size == 0:
            return np.zeros(0, np.int32), np.zeros(0)
        return a[:, 0].astype(np.int32), a[:, 2].astype(np.float64)

    b_idx, b_val = spvec(prefix + "b.txt")
    C_idx, C_val = spvec(prefix + "C.txt")
    return dict(blk=np.asarray(blk, np.int32), vec_len=vec_len, con_num=con_num, col_ptrs=col_ptrs,
                row_ids=rows.astype(np.int32), vals=vals.astype(np.float64), b_idx=b_idx, b_val=b_val,
                C_idx=C_idx, C_val=C_val)
