// cpu_baseline.cpp — TEST / MEASUREMENT INFRASTRUCTURE ONLY (never linked or loaded by the product).
//
// C++ restatement of the reference's CPU projection path ("baseline B"): LAPACK dsyevd on a pool of
// std::threads, then clamp and Q diag(w+) Q^T, for every block of an svec vector.  Follows
//   * svec -> dense      vector_to_matrices_kernel        src/kernels/vec_mat_conversion.cu:11-34
//   * dsyevd('V','U')    single_eig_lapack                include/cuadmm/eig_cpu.h:31-51
//                        workspace sizes lwork = 1 + 6n + 2n^2, liwork = 3 + 5n   src/duo_solver.cu:373-377
//   * thread pool        equal-count contiguous ranges, remainder handed out from the last thread
//                                                         src/duo_solver.cu:346-371, 578-619
//   * clamp              max_dense_vector_zero            src/kernels/dense_scalar.cu:41-47
//   * rebuild            dense_matrix_mul_diag_batch + gemm (Q diag(w+)) Q^T   src/kernels/diagonal_batch.cu:11-22,
//                                                         include/cuadmm/cublas.h:18-35
//   * dense -> svec      matrices_to_vector_kernel        src/kernels/vec_mat_conversion.cu:36-57
// LAPACK/BLAS are third-party, not vendored by the reference (it links MATLAB's, CMakeLists.txt:25-27,75): here the
// OpenBLAS that ships inside scipy (scipy.libs/libscipy_openblas*.so, symbols scipy_dsyevd_ / scipy_dgemm_, LP64),
// dlopen'ed from the path the caller passes; one BLAS thread per call, the pool is the only parallelism.
// Pinned by tests/test_cpu_baseline.py: the reference's 4x4 known-answer matrix (test/eig_cpu_test.hpp:7-66) and
// agreement with oracle_np.project_svec.
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <thread>
#include <vector>

namespace {

typedef void (*dsyevd_t)(const char* jobz, const char* uplo, const int* n, double* a, const int* lda, double* w, double* work,
                         const int* lwork, int* iwork, const int* liwork, int* info);
typedef void (*dgemm_t)(const char* ta, const char* tb, const int* m, const int* n, const int* k, const double* alpha,
                        const double* a, const int* lda, const double* b, const int* ldb, const double* beta, double* c, const int* ldc);
typedef void (*setthreads_t)(int);

dsyevd_t p_dsyevd = nullptr;
dgemm_t p_dgemm = nullptr;

double sqrt2_newton() {   // include/cuadmm/kernels.h:173-181
    double prev = 0.0, curr = 2.0;
    while (curr != prev) { prev = curr; curr = 0.5 * (curr + 2.0 / curr); }
    return curr;
}
const double SQRT2 = sqrt2_newton();
const double SQRT2INV = 1.0 / SQRT2;

// one block: svec -> dense, eig, clamp, rebuild, dense -> svec; returns LAPACK info
int project_block(int n, const double* in, double* out, double* eig_out, std::vector<double>& mat, std::vector<double>& tmp,
                  std::vector<double>& prod, std::vector<double>& w, std::vector<double>& work, std::vector<int>& iwork) {
    const size_t nn = (size_t)n * n;
    if (mat.size() < nn) { mat.resize(nn); tmp.resize(nn); prod.resize(nn); }
    if ((int)w.size() < n) w.resize(n);
    const int lwork = 1 + 6 * n + 2 * n * n, liwork = 3 + 5 * n;
    if ((int)work.size() < lwork) work.resize(lwork);
    if ((int)iwork.size() < liwork) iwork.resize(liwork);
    // column-major, both triangles (vec_mat_conversion.cu:23-31)
    size_t idx = 0;
    for (int c = 0; c < n; ++c)
        for (int r = 0; r <= c; ++r, ++idx) {
            const double v = (r == c) ? in[idx] : in[idx] * SQRT2INV;
            mat[(size_t)c * n + r] = v;
            mat[(size_t)r * n + c] = v;
        }
    int info = 0;
    p_dsyevd("V", "U", &n, mat.data(), &n, w.data(), work.data(), &lwork, iwork.data(), &liwork, &info);
    if (eig_out) memcpy(eig_out, w.data(), sizeof(double) * n);
    for (int j = 0; j < n; ++j) {
        const double wj = w[j] > 0.0 ? w[j] : 0.0;                       // dense_scalar.cu:41-47
        for (int i = 0; i < n; ++i) tmp[(size_t)j * n + i] = mat[(size_t)j * n + i] * wj;   // diagonal_batch.cu:18-19
    }
    const double one = 1.0, zero = 0.0;
    p_dgemm("N", "T", &n, &n, &n, &one, tmp.data(), &n, mat.data(), &n, &zero, prod.data(), &n);   // cublas.h:18-35
    idx = 0;
    for (int c = 0; c < n; ++c)
        for (int r = 0; r <= c; ++r, ++idx) {
            const double v = prod[(size_t)c * n + r];
            out[idx] = (r == c) ? v : v * SQRT2;                          // vec_mat_conversion.cu:48-54
        }
    return info;
}

}  // namespace

extern "C" {

// dlopen the BLAS/LAPACK library; returns 0 on success
int cb_init(const char* openblas_path) {
    void* h = dlopen(openblas_path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) return -1;
    p_dsyevd = (dsyevd_t)dlsym(h, "scipy_dsyevd_");
    if (!p_dsyevd) p_dsyevd = (dsyevd_t)dlsym(h, "dsyevd_");
    p_dgemm = (dgemm_t)dlsym(h, "scipy_dgemm_");
    if (!p_dgemm) p_dgemm = (dgemm_t)dlsym(h, "dgemm_");
    if (!p_dsyevd || !p_dgemm) return -2;
    setthreads_t st = (setthreads_t)dlsym(h, "scipy_openblas_set_num_threads");
    if (!st) st = (setthreads_t)dlsym(h, "openblas_set_num_threads");
    if (st) st(1);
    return 0;
}

// thread ranges of src/duo_solver.cu:346-371; ptrs has nthreads + 1 entries
void cb_thread_ranges(int count, int nthreads, int* ptrs) {
    const int T = nthreads < 1 ? 1 : nthreads;
    std::vector<int> per(T, count / T);
    per[T - 1] = count - (T - 1) * (count / T);
    if (T > 2) {
        int i = 0;
        while (i < T - 1 && per[T - 1] - per[i] >= 2) { per[i] += 1; per[T - 1] -= 1; ++i; }
    }
    ptrs[0] = 0;
    for (int t = 0; t < T; ++t) ptrs[t + 1] = ptrs[t] + per[t];
}

// Xproj = Pi+(Xb) block by block on `nthreads` std::threads; eig (may be NULL) gets the eigenvalues, ascending per block.
// Returns the number of blocks whose dsyevd reported info != 0.
int cb_project(const int* blk, int nblk, const double* Xb, double* Xproj, double* eig, int nthreads) {
    if (!p_dsyevd) return -1;
    std::vector<int64_t> off(nblk + 1, 0), eoff(nblk + 1, 0);
    for (int k = 0; k < nblk; ++k) { off[k + 1] = off[k] + (int64_t)blk[k] * (blk[k] + 1) / 2; eoff[k + 1] = eoff[k] + blk[k]; }
    const int T = nthreads < 1 ? 1 : nthreads;
    std::vector<int> ptrs(T + 1);
    cb_thread_ranges(nblk, T, ptrs.data());
    std::vector<int> bad(T, 0);
    std::vector<std::thread> pool;
    for (int t = 0; t < T; ++t) {
        pool.emplace_back([&, t] {
            std::vector<double> mat, tmp, prod, w, work;
            std::vector<int> iwork;
            for (int k = ptrs[t]; k < ptrs[t + 1]; ++k)
                if (project_block(blk[k], Xb + off[k], Xproj + off[k], eig ? eig + eoff[k] : nullptr, mat, tmp, prod, w, work, iwork) != 0) ++bad[t];
        });
    }
    for (auto& th : pool) th.join();
    int nbad = 0;
    for (int t = 0; t < T; ++t) nbad += bad[t];
    return nbad;
}

}  // extern "C"
