// spmv.h — CSR sparse matrix on the device + y = alpha*A*x + beta*y kernels with fused epilogues.
#pragma once
#include "common.h"

struct cuadmm_spmv_s {
    int device = -1;
    int64_t rows = 0, cols = 0, nnz = 0;
    cuadmm::DevBuf<int32_t> rowptr, colind;
    cuadmm::DevBuf<double> val;
    int group = 1;   // lanes cooperating on one row (1,2,4,8,16,32), chosen from nnz/row
    cuadmm::DevBuf<double> d_x, d_y;  // staging for the *_host entry point
    int64_t alg_bytes() const { return 12 * nnz + 4 * (rows + 1) + 8 * rows + 8 * cols; }
};

namespace cuadmm {

// Fused epilogues of the ADMM iteration (all optional; see solver.cu for the call sites).
// The row result r = (A x)_i is combined as out_i = alpha * r + beta * y_i (+ extras).
struct SpmvEpilogue {
    // mode 0: plain                              y = alpha*r + beta*y
    // mode 1: rhsy = Rp/sig - A*SmC              y = aux1[i] / sig - r              (src/solver.cu:478-482)
    // mode 2: Rd1 = At*y - C; Xb = X + sig*Rd1   y(=Rd1) = r - aux1[i]; out2[i] = aux2[i] + sig*y   (:514-527)
    // mode 3: Rd1 = r - C; Rd = Rd1 + S; X += tau*sig*Rd; partial sums of |Rd|^2 and <C,X>   (:721-758,775-776)
    // mode 4: Rp = b - A*X; partial sums of |normA*Rp|^2                                       (:764-772)
    // mode 5: sharded solver, fused with the reduction over ranks: the partial row alpha*r is stored straight into the
    //         staging area of the rank that reduces row i (push[i / slice], peer memory over NVLink; see peer.h)
    int mode = 0;
    const double* aux1 = nullptr;
    const double* aux2 = nullptr;
    const double* aux3 = nullptr;
    double* out2 = nullptr;
    double* out3 = nullptr;
    const double* scal = nullptr;   // device scalars (sig, tau, ...) — see solver state layout
    double* partial = nullptr;      // per-CTA partial sums (2 per CTA), deterministic two-stage reduction
    double* push[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // mode 5: staging area of every rank
    unsigned slice = 1;             // mode 5: rows reduced by one rank
    int rank = 0;                   // mode 5: this rank
};

void spmv_launch(const cuadmm_spmv_s& A, double alpha, const double* x, double beta, double* y,
                 const SpmvEpilogue& epi, cudaStream_t stream, int* grid_out = nullptr);
int spmv_grid(const cuadmm_spmv_s& A);
cuadmm_spmv_s* spmv_create(int64_t rows, int64_t cols, int64_t nnz, const int32_t* h_rowptr,
                         const int32_t* h_colind, const double* h_val, int device);
// kernels launched on this matrix return immediately once *flag != 0 (device-resident stop flag)
void spmv_set_done_flag(cuadmm_spmv_s& A, const int* flag);

}  // namespace cuadmm
