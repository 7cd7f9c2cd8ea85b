// ref_driver.cu — TEST/BENCH INFRASTRUCTURE ONLY.
//
// Thin extern "C" driver over the UNMODIFIED reference sources, which oracle/Makefile compiles
// where they lie under /root/reference into oracle/_ref/libcuadmm_ref.so (git-ignored).  It
// exposes (1) the reference's host integer logic (analyze_blk, MatrixSizes, get_maps, COO_to_CSC,
// read_blk, get_inverse_permutation) so the C restatement in oracle_host.c and the product can be
// checked bit-exactly against the real thing, and (2) "Baseline A": the reference's own cuSOLVER
// projection stage, transcribed call by call from src/solver.cu:531-647 on top of the reference's
// own wrappers (include/cuadmm/{cusolver,cublas,kernels}.h), used as the GPU oracle and as the
// in-run speed baseline.  No reference source is copied into this repository.
#include <sstream>
#include <iostream>
#include <chrono>
#include <algorithm>
#include "cuadmm/memory.h"
#include "cuadmm/utils.h"
#include "cuadmm/matrix_sizes.h"
#include "cuadmm/kernels.h"
#include "cuadmm/cusolver.h"
#include "cuadmm/cublas.h"
#include "cuadmm/cusparse.h"
#include "cuadmm/io.h"

namespace {
struct CoutSilencer {
    std::streambuf* old;
    std::ostringstream sink;
    CoutSilencer() { old = std::cout.rdbuf(sink.rdbuf()); }
    ~CoutSilencer() { std::cout.rdbuf(old); }
};
}

extern "C" {

int ref_is_large_mat(int n, int cnt) { return is_large_mat(n, cnt) ? 1 : 0; }

// analyze_blk + MatrixSizes::init; returns number of distinct sizes
int ref_analyze(const int* blk, int nblk, int* sizes, int* nums, int* is_large,
                int* totals /*6*/, int* large_start, int* large_W_start, int* small_start, int* small_W_start,
                int* n_large_groups, int* n_small_groups) {
    CoutSilencer q;
    HostDenseVector<int> hblk(nblk);
    memcpy(hblk.vals, blk, sizeof(int) * nblk);
    std::vector<int> s, c;
    analyze_blk(hblk, s, c);
    MatrixSizes ms;
    ms.init(s, c);
    for (size_t i = 0; i < s.size(); ++i) { sizes[i] = s[i]; nums[i] = c[i]; is_large[i] = ms.is_large(s[i]) ? 1 : 0; }
    totals[0] = ms.large_mat_num; totals[1] = ms.sum_large_mat_size; totals[2] = ms.total_large_mat_size;
    totals[3] = ms.small_mat_num; totals[4] = ms.sum_small_mat_size; totals[5] = ms.total_small_mat_size;
    for (size_t i = 0; i < ms.large_mat_start_indices.size(); ++i) { large_start[i] = ms.large_mat_start_indices[i]; large_W_start[i] = ms.large_W_start_indices[i]; }
    for (size_t i = 0; i < ms.small_mat_start_indices.size(); ++i) { small_start[i] = ms.small_mat_start_indices[i]; small_W_start[i] = ms.small_W_start_indices[i]; }
    *n_large_groups = (int)ms.large_mat_sizes.size();
    *n_small_groups = (int)ms.small_mat_sizes.size();
    return (int)s.size();
}

void ref_get_maps(const int* blk, int nblk, int vec_len, int* map_B, int* map_M1, int* map_M2) {
    CoutSilencer q;
    HostDenseVector<int> hblk(nblk);
    memcpy(hblk.vals, blk, sizeof(int) * nblk);
    std::vector<int> s, c;
    analyze_blk(hblk, s, c);
    MatrixSizes ms;
    ms.init(s, c);
    std::vector<int> B, M1, M2;
    get_maps(hblk, vec_len, B, M1, M2, ms);
    memcpy(map_B, B.data(), sizeof(int) * vec_len);
    memcpy(map_M1, M1.data(), sizeof(int) * vec_len);
    memcpy(map_M2, M2.data(), sizeof(int) * vec_len);
}

void ref_get_maps_duo(const int* blk, int nblk, int LARGE, int SMALL, int vec_len, int* map_B, int* map_M1, int* map_M2) {
    HostDenseVector<int> hblk(nblk);
    memcpy(hblk.vals, blk, sizeof(int) * nblk);
    std::vector<int> B, M1, M2;
    get_maps_duo(hblk, LARGE, SMALL, vec_len, B, M1, M2);
    memcpy(map_B, B.data(), sizeof(int) * vec_len);
    memcpy(map_M1, M1.data(), sizeof(int) * vec_len);
    memcpy(map_M2, M2.data(), sizeof(int) * vec_len);
}

void ref_inverse_permutation(const int* perm, int n, int* perm_inv) {
    std::vector<int> p(perm, perm + n), inv;
    get_inverse_permutation(inv, p);
    memcpy(perm_inv, inv.data(), sizeof(int) * n);
}

void ref_coo_to_csc(int* col_ptrs, int* col_ids, int* row_ids, double* vals, int nnz, int col_num) {
    CoutSilencer q;
    std::vector<int> cp(col_num + 1, 0), ci(col_ids, col_ids + nnz), ri(row_ids, row_ids + nnz);
    std::vector<double> v(vals, vals + nnz);
    COO_to_CSC(cp, ci, ri, v, nnz, col_num);
    memcpy(col_ptrs, cp.data(), sizeof(int) * (col_num + 1));
    memcpy(col_ids, ci.data(), sizeof(int) * nnz);
    memcpy(row_ids, ri.data(), sizeof(int) * nnz);
    memcpy(vals, v.data(), sizeof(double) * nnz);
}

int ref_read_blk(const char* path, char* types, int* vals, int cap) {
    std::vector<std::tuple<char, int>> v;
    read_blk(path, v);
    int n = (int)v.size();
    for (int i = 0; i < n && i < cap; ++i) { types[i] = std::get<0>(v[i]); vals[i] = std::get<1>(v[i]); }
    return n;
}

double ref_sqrt2() { return SQRT2; }
double ref_sqrt2inv() { return SQRT2INV; }

// ------------------------------------------------------------------------------------------
// Baseline A: the reference's cuSOLVER projection stage (src/solver.cu:531-647), on GPU 0.
// ------------------------------------------------------------------------------------------
struct RefProj {
    int vec_len = 0, nstream = 0;
    MatrixSizes sizes;
    std::vector<int> blk_sizes, blk_nums;
    DeviceDenseVector<int> map_B, map_M1, map_M2;
    DeviceDenseVector<double> Xb, Xproj;
    DeviceDenseVector<double> large_mat, large_W, small_mat, small_W;
    DeviceDenseVector<double> large_mat_tmp, small_mat_tmp, large_mat_P, small_mat_P;
    DeviceDenseVector<int> large_info, small_info;
    std::vector<DeviceStream> eig_stream_arr;
    std::vector<DeviceSolverDnHandle> cusolverH_eig_large_arr;
    DeviceSolverDnHandle cusolverH_eig_small;
    DeviceBlasHandle cublasH;
    SingleEigParameter eig_param_single;
    BatchEigParameter eig_param_batch;
    std::vector<size_t> eig_large_buffer_size, cpu_eig_large_buffer_size, eig_small_buffer_size;
    DeviceDenseVector<double> eig_large_buffer, eig_small_buffer;
    HostDenseVector<double> cpu_eig_large_buffer;
    // 64-bit offsets (the reference keeps these in int and overflows for n >~ 6000; "patched")
    std::vector<size_t> large_buf_start, large_cpu_buf_start, small_buf_start;
    cudaEvent_t e0, e1;
};

void* ref_proj_create(const int* blk, int nblk, int eig_stream_num_per_gpu) {
    CoutSilencer q;
    RefProj* R = new RefProj();
    HostDenseVector<int> cpu_blk(nblk);
    memcpy(cpu_blk.vals, blk, sizeof(int) * nblk);
    int vec_len = 0;
    for (int i = 0; i < nblk; ++i) vec_len += blk[i] * (blk[i] + 1) / 2;
    R->vec_len = vec_len;
    analyze_blk(cpu_blk, R->blk_sizes, R->blk_nums);
    R->sizes.init(R->blk_sizes, R->blk_nums);
    std::vector<int> B, M1, M2;
    get_maps(cpu_blk, vec_len, B, M1, M2, R->sizes);
    R->map_B.allocate(GPU0, vec_len); R->map_M1.allocate(GPU0, vec_len); R->map_M2.allocate(GPU0, vec_len);
    cudaMemcpy(R->map_B.vals, B.data(), sizeof(int) * vec_len, H2D);
    cudaMemcpy(R->map_M1.vals, M1.data(), sizeof(int) * vec_len, H2D);
    cudaMemcpy(R->map_M2.vals, M2.data(), sizeof(int) * vec_len, H2D);
    R->Xb.allocate(GPU0, vec_len); R->Xproj.allocate(GPU0, vec_len);
    R->cublasH.set_gpu_id(GPU0); R->cublasH.activate();

    // src/solver.cu:230-283
    R->large_mat.allocate(GPU0, std::max(R->sizes.total_large_mat_size, 1));
    R->large_W.allocate(GPU0, std::max(R->sizes.sum_large_mat_size, 1));
    R->large_info.allocate(GPU0, std::max(R->sizes.large_mat_num, 1));
    R->nstream = eig_stream_num_per_gpu;
    R->eig_stream_arr = std::vector<DeviceStream>(R->nstream);
    R->cusolverH_eig_large_arr = std::vector<DeviceSolverDnHandle>(R->nstream);
    for (int i = 0; i < R->nstream; ++i) {
        R->eig_stream_arr[i].set_gpu_id(GPU0); R->eig_stream_arr[i].activate();
        R->cusolverH_eig_large_arr[i].set_gpu_id(GPU0); R->cusolverH_eig_large_arr[i].activate(R->eig_stream_arr[i]);
    }
    size_t nl = R->sizes.large_mat_sizes.size();
    R->eig_large_buffer_size.assign(nl, 0); R->cpu_eig_large_buffer_size.assign(nl, 0);
    R->large_buf_start.assign(1, 0); R->large_cpu_buf_start.assign(1, 0);
    for (size_t i = 0; i < nl; ++i) {
        single_eig_get_buffersize_cusolver(
            R->cusolverH_eig_large_arr[i % R->nstream], R->eig_param_single, R->large_mat, R->large_W,
            R->sizes.large_mat_sizes[i], &R->eig_large_buffer_size[i], &R->cpu_eig_large_buffer_size[i],
            R->sizes.large_mat_offset(i, 0), R->sizes.large_W_offset(i, 0));
        R->large_buf_start.push_back(R->large_buf_start[i] + (size_t)R->sizes.large_mat_nums[i] * R->eig_large_buffer_size[i]);
        R->large_cpu_buf_start.push_back(R->large_cpu_buf_start[i] + (size_t)R->sizes.large_mat_nums[i] * R->cpu_eig_large_buffer_size[i]);
    }
    // DeviceDenseVector<double>::allocate(.., size, true) takes a byte count in the reference; sizes are int there
    R->eig_large_buffer.allocate(GPU0, (int)(R->large_buf_start.back() / sizeof(double) + 1));
    R->cpu_eig_large_buffer.allocate((int)(R->large_cpu_buf_start.back() / sizeof(double) + 1));

    // src/solver.cu:285-312
    R->cusolverH_eig_small.set_gpu_id(GPU0); R->cusolverH_eig_small.activate();
    R->small_mat.allocate(GPU0, std::max(R->sizes.total_small_mat_size, 1));
    R->small_W.allocate(GPU0, std::max(R->sizes.sum_small_mat_size, 1));
    R->small_info.allocate(GPU0, std::max(R->sizes.small_mat_num, 1));
    R->small_buf_start.assign(1, 0);
    for (size_t i = 0; i < R->sizes.small_mat_sizes.size(); ++i) {
        R->eig_small_buffer_size.push_back(batch_eig_get_buffersize_cusolver(
            R->cusolverH_eig_small, R->eig_param_batch, R->small_mat, R->small_W,
            R->sizes.small_mat_sizes[i], R->sizes.small_mat_nums[i],
            R->sizes.small_mat_offset(i), R->sizes.small_W_offset(i)));
        R->small_buf_start.push_back(R->small_buf_start[i] + R->eig_small_buffer_size[i]);
    }
    R->eig_small_buffer.allocate(GPU0, (int)(R->small_buf_start.back() / sizeof(double) + 1));
    // src/solver.cu:315-318
    R->large_mat_tmp.allocate(GPU0, std::max(R->sizes.total_large_mat_size, 1));
    R->small_mat_tmp.allocate(GPU0, std::max(R->sizes.total_small_mat_size, 1));
    R->large_mat_P.allocate(GPU0, std::max(R->sizes.total_large_mat_size, 1));
    R->small_mat_P.allocate(GPU0, std::max(R->sizes.total_small_mat_size, 1));
    cudaEventCreate(&R->e0); cudaEventCreate(&R->e1);
    cudaDeviceSynchronize();
    return R;
}

// one projection pass, device-resident input already in R->Xb  (src/solver.cu:531-647)
static void ref_proj_step(RefProj* R) {
    vector_to_matrices(R->Xb, R->large_mat, R->small_mat, R->map_B, R->map_M1, R->map_M2);
    // the reference has no sync between the default-stream scatter and the non-blocking eig
    // streams; the transcription adds it so the baseline is correct
    CHECK_CUDA( cudaDeviceSynchronize() );
    int counter = 0;
    for (size_t i = 0; i < R->sizes.large_mat_sizes.size(); ++i) {
        for (int j = 0; j < R->sizes.large_mat_nums[i]; ++j) {
            int stream_id = counter % R->nstream;
            single_eig_cusolver(
                R->cusolverH_eig_large_arr[stream_id], R->eig_param_single,
                R->large_mat, R->large_W, R->eig_large_buffer, R->cpu_eig_large_buffer, R->large_info,
                R->sizes.large_mat_sizes[i], R->eig_large_buffer_size[i], R->cpu_eig_large_buffer_size[i],
                R->sizes.large_mat_offset(i, j), R->sizes.large_W_offset(i, j),
                R->large_buf_start[i] + R->eig_large_buffer_size[i] * (size_t)j,
                R->large_cpu_buf_start[i] + R->cpu_eig_large_buffer_size[i] * (size_t)j,
                counter);
            counter++;
        }
    }
    for (int s = 0; s < R->nstream; ++s) CHECK_CUDA( cudaStreamSynchronize(R->eig_stream_arr[s].stream) );
    int info_offset = 0;
    for (size_t i = 0; i < R->sizes.small_mat_sizes.size(); ++i) {
        batch_eig_cusolver(
            R->cusolverH_eig_small, R->eig_param_batch, R->small_mat, R->small_W,
            R->eig_small_buffer, R->small_info,
            R->sizes.small_mat_sizes[i], R->sizes.small_mat_nums[i], R->eig_small_buffer_size[i],
            R->sizes.small_mat_offset(i), R->sizes.small_W_offset(i),
            R->small_buf_start[i], 0, info_offset);
        info_offset += R->sizes.small_mat_nums[i];
    }
    if (R->large_W.size > 0) max_dense_vector_zero(R->large_W);
    if (R->small_W.size > 0) max_dense_vector_zero(R->small_W);
    for (size_t i = 0; i < R->sizes.large_mat_sizes.size(); ++i)
        dense_matrix_mul_diag_batch(R->large_mat_tmp, R->large_mat, R->large_W,
            R->sizes.large_mat_sizes[i], R->sizes.large_mat_nums[i],
            R->sizes.large_mat_offset(i, 0), R->sizes.large_W_offset(i, 0));
    for (size_t i = 0; i < R->sizes.small_mat_sizes.size(); ++i)
        dense_matrix_mul_diag_batch(R->small_mat_tmp, R->small_mat, R->small_W,
            R->sizes.small_mat_sizes[i], R->sizes.small_mat_nums[i],
            R->sizes.small_mat_offset(i), R->sizes.small_W_offset(i));
    for (size_t i = 0; i < R->sizes.large_mat_sizes.size(); ++i)
        dense_matrix_mul_trans_batch(R->cublasH, R->large_mat_P, R->large_mat_tmp, R->large_mat,
            R->sizes.large_mat_sizes[i], R->sizes.large_mat_nums[i], R->sizes.large_mat_offset(i, 0));
    for (size_t i = 0; i < R->sizes.small_mat_sizes.size(); ++i)
        dense_matrix_mul_trans_batch(R->cublasH, R->small_mat_P, R->small_mat_tmp, R->small_mat,
            R->sizes.small_mat_sizes[i], R->sizes.small_mat_nums[i], R->sizes.small_mat_offset(i));
    matrices_to_vector(R->Xproj, R->large_mat_P, R->small_mat_P, R->map_B, R->map_M1, R->map_M2);
}

// host in / host out; returns average wall ms per projection over `reps` passes (after 1 warm-up
// when reps > 1), measured host-side around device syncs because the stage itself syncs the host.
double ref_proj_run(void* h, const double* Xb_host, double* Xproj_host, int reps) {
    RefProj* R = (RefProj*)h;
    cudaMemcpy(R->Xb.vals, Xb_host, sizeof(double) * R->vec_len, H2D);
    if (reps > 1) { ref_proj_step(R); cudaDeviceSynchronize(); }
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r) ref_proj_step(R);
    cudaDeviceSynchronize();
    auto t1 = std::chrono::steady_clock::now();
    if (Xproj_host) cudaMemcpy(Xproj_host, R->Xproj.vals, sizeof(double) * R->vec_len, D2H);
    return std::chrono::duration<double, std::milli>(t1 - t0).count() / reps;
}

void ref_proj_destroy(void* h) { delete (RefProj*)h; }

}  // extern "C"
