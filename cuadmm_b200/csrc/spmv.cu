// spmv.cu — hand-written CSR SpMV for the constraint operators A (m x vec_len) and At
// (vec_len x m), replacing cusparseSpMV(CSR_ALG1) (include/cuadmm/cusparse.h:70-83) and, through
// the fused epilogues, the axpy / axpby / elementwise launches around each call site in
// src/solver.cu (478-482, 514-527, 695-699, 721-758, 764-777).
//
// Shape of the data: SPOT/moment matrices have ~2-5 non-zeros per constraint row, At has many
// empty rows and a handful of very long ones (a shared moment entry used by thousands of
// constraints).  So: G lanes per row (G from the mean row length) for ordinary rows, and one warp
// per "long" row (> kLongRow non-zeros; > 16 for the large one-lane-per-row operators) appended to the same grid.  HBM-bound: 12 B per non-zero
// (value + int32 column), coalesced across the lanes of a group.
#include "spmv.h"
#include <algorithm>
#include <cstdlib>

namespace cuadmm {

static constexpr int kLongRow = 256;
static constexpr int kThreads = 256;

// operands of the epilogue of row i.  They do not depend on the product, so the main loop requests them together with the
// row pointers: issued after the dot product they were a fourth dependent round trip (40 % of the kernel's stall samples).
struct SpmvPre { double a1, a2, a3; };
template <int MODE>
__device__ __forceinline__ SpmvPre spmv_prefetch(const SpmvEpilogue& e, double beta, const double* y, int64_t i) {
    SpmvPre p = {0.0, 0.0, 0.0};
    switch (MODE) {
        case 0: if (beta != 0.0) p.a1 = y[i]; break;
        case 1: p.a1 = e.aux1[i]; break;
        case 2: p.a1 = e.aux1[i]; p.a2 = e.aux2[i]; break;
        case 3: p.a1 = e.aux1[i]; p.a2 = e.aux2[i]; p.a3 = e.out2[i]; break;
        case 4: p.a1 = e.aux1[i]; p.a2 = e.aux2[i]; p.a3 = e.aux3[i]; break;
        default: break;
    }
    return p;
}

template <int MODE>
__device__ __forceinline__ void spmv_store(const SpmvEpilogue& e, double alpha, double beta, double* y,
                                           int64_t i, double r, const SpmvPre& pre, double& acc0, double& acc1) {
    switch (MODE) {
        case 0:
            y[i] = (beta == 0.0) ? alpha * r : alpha * r + beta * pre.a1;
            break;
        case 1: {  // rhsy = Rp / sig - A*SmC
            const double sig = e.scal[0];
            y[i] = pre.a1 / sig - r;
            break;
        }
        case 2: {  // Rd1 = At*y - C ; Xb = X + sig * Rd1
            const double sig = e.scal[0];
            const double rd1 = r - pre.a1;
            y[i] = rd1;
            e.out2[i] = pre.a2 + sig * rd1;
            break;
        }
        case 3: {  // Rd1 = At*y - C ; Rd = Rd1 + S ; X += tau*sig*Rd ; sums |Rd|^2, <C,X>
            const double sig = e.scal[0], tau = e.scal[1];
            const double c = pre.a1;
            const double rd = (r - c) + pre.a2;
            y[i] = rd;                                   // Rd
            const double xn = pre.a3 + (tau * sig) * rd;
            e.out2[i] = xn;                              // X
            acc0 = fma(rd, rd, acc0);
            acc1 = fma(c, xn, acc1);
            break;
        }
        case 4: {  // Rp = b - A*X ; sums |normA .* Rp|^2 and <b, y>
            const double b = pre.a1;
            const double rp = b - r;
            y[i] = rp;
            const double t = pre.a2 * rp;                // normA[i] * Rp[i]  (bscale applied by the caller)
            acc0 = fma(t, t, acc0);
            acc1 = fma(b, pre.a3, acc1);                 // <b, y>
            break;
        }
        case 5: {  // partial row of this rank -> staging slot [rank] of the rank that reduces row i
            const unsigned q = (unsigned)i / e.slice;
            e.push[q][(size_t)e.rank * e.slice + ((unsigned)i - q * e.slice)] = alpha * r;
            break;
        }
    }
}

#ifndef CUADMM_SPMV_MINB
#define CUADMM_SPMV_MINB 5
#endif
template <int G, int MODE>
__global__ void __launch_bounds__(kThreads, CUADMM_SPMV_MINB) spmv_csr_kernel(
        int64_t rows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
        const double* __restrict__ val, const double* __restrict__ x, double* y, double alpha, double beta,
        SpmvEpilogue e, const int32_t* __restrict__ long_rows, int n_long, int main_blocks, int long_thr,
        const int* __restrict__ done_flag) {
    if (done_flag && *done_flag) return;
    __shared__ double red[2][kThreads / 32];
    double acc0 = 0.0, acc1 = 0.0;
    const int lane = threadIdx.x & 31;
    if ((int)blockIdx.x < main_blocks) {
        const int sub = threadIdx.x % G;
        const int64_t gid = ((int64_t)blockIdx.x * kThreads + threadIdx.x) / G;
        const int64_t ngroups = (int64_t)main_blocks * kThreads / G;
        const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane / G * G));
        // U rows per group in flight: the loads of one row are a chain of three dependent round trips
        // (row pointer -> column/value -> x), and with ~2 entries per row there is nothing else to
        // overlap them with, so independent rows are interleaved by hand.
        // One lane per row (G == 1, the large operators): two rows with their epilogue operands prefetched — measured on the
        // bench operator (scripts/stage_split_probe.py, SpMV + vector kernels per iteration): 0.146 ms as before (U = 4,
        // operands loaded after the product), 0.119 ms so; three or four rows spill at the 48 registers that keep five
        // CTAs per SM resident and were slower (0.126-0.135).  Lane groups (small, latency-bound operators) keep four rows
        // and load the operands of the one storing lane late.
        constexpr int U = (G == 1) ? 2 : 4;
        constexpr bool PRE = (G == 1);
        for (int64_t i0 = gid; i0 < rows; i0 += ngroups * U) {
            int p0[U], p1[U];
            bool mine[U];
            double r[U];
            SpmvPre pre[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = i0 + (int64_t)u * ngroups;
                p0[u] = 0; p1[u] = 0;
                pre[u] = SpmvPre{0.0, 0.0, 0.0};
                if (i < rows) {
                    p0[u] = rowptr[i]; p1[u] = rowptr[i + 1];
                    if (PRE && sub == 0) pre[u] = spmv_prefetch<MODE>(e, beta, y, i);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = i0 + (int64_t)u * ngroups;
                mine[u] = i < rows && p1[u] - p0[u] <= long_thr;      // long rows belong to the warps below
                if (!mine[u]) p1[u] = p0[u];
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int p = p0[u] + sub;
                r[u] = 0.0;
                if (p < p1[u]) r[u] = fma(val[p], x[colind[p]], 0.0);
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                for (int p = p0[u] + sub + G; p < p1[u]; p += G) r[u] = fma(val[p], x[colind[p]], r[u]);
#pragma unroll
            for (int u = 0; u < U; ++u) {
#pragma unroll
                for (int o = G / 2; o > 0; o >>= 1) r[u] += __shfl_xor_sync(gmask, r[u], o);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = i0 + (int64_t)u * ngroups;
                if (sub == 0 && mine[u]) {
                    if (!PRE) pre[u] = spmv_prefetch<MODE>(e, beta, y, i);
                    spmv_store<MODE>(e, alpha, beta, y, i, r[u], pre[u], acc0, acc1);
                }
            }
        }
    } else {
        // long rows: one warp per row
        const int w = ((int)blockIdx.x - main_blocks) * (kThreads / 32) + (threadIdx.x >> 5);
        const int nw = ((int)gridDim.x - main_blocks) * (kThreads / 32);
        for (int li = w; li < n_long; li += nw) {
            const int64_t i = long_rows[li];
            const int p0 = rowptr[i], p1 = rowptr[i + 1];
            SpmvPre pre = {0.0, 0.0, 0.0};
            if (lane == 0) pre = spmv_prefetch<MODE>(e, beta, y, i);
            double r = 0.0;
            for (int p = p0 + lane; p < p1; p += 32) r = fma(val[p], x[colind[p]], r);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
            if (lane == 0) spmv_store<MODE>(e, alpha, beta, y, i, r, pre, acc0, acc1);
        }
    }
    if ((MODE == 3 || MODE == 4) && e.partial) {
        // deterministic two-stage reduction: fixed tree inside the CTA, fixed order across CTAs later
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
            acc1 += __shfl_xor_sync(0xffffffffu, acc1, o);
        }
        if (lane == 0) { red[0][threadIdx.x >> 5] = acc0; red[1][threadIdx.x >> 5] = acc1; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double s0 = 0.0, s1 = 0.0;
            for (int k = 0; k < kThreads / 32; ++k) { s0 += red[0][k]; s1 += red[1][k]; }
            e.partial[2 * blockIdx.x] = s0;
            e.partial[2 * blockIdx.x + 1] = s1;
        }
    }
}

struct SpmvAux {
    DevBuf<int32_t> long_rows;
    int n_long = 0;
    int long_thr = kLongRow;       // rows with more entries than this are summed by a whole warp
    int main_blocks = 0, long_blocks = 0;
};

}  // namespace cuadmm

using namespace cuadmm;

// the aux data (long-row list, grid shape) hangs off the handle
struct cuadmm_spmv_full : cuadmm_spmv_s {
    SpmvAux aux;
    const int* done_flag = nullptr;
};

namespace cuadmm {

int spmv_grid(const cuadmm_spmv_s& A_) {
    const cuadmm_spmv_full& A = static_cast<const cuadmm_spmv_full&>(A_);
    return A.aux.main_blocks + A.aux.long_blocks;
}

void spmv_set_done_flag(cuadmm_spmv_s& A_, const int* flag) {
    static_cast<cuadmm_spmv_full&>(A_).done_flag = flag;
}

void spmv_launch(const cuadmm_spmv_s& A_, double alpha, const double* x, double beta, double* y,
                 const SpmvEpilogue& epi, cudaStream_t stream, int* grid_out) {
    const cuadmm_spmv_full& A = static_cast<const cuadmm_spmv_full&>(A_);
    const int grid = A.aux.main_blocks + A.aux.long_blocks;
    if (grid_out) *grid_out = grid;
    if (A.rows == 0 || grid == 0) return;
#define CUADMM_SPMV_LAUNCH(G, MODE)                                                                     \
    spmv_csr_kernel<G, MODE><<<grid, kThreads, 0, stream>>>(A.rows, A.rowptr.p, A.colind.p, A.val.p, x, y, \
        alpha, beta, epi, A.aux.long_rows.p, A.aux.n_long, A.aux.main_blocks, A.aux.long_thr, A.done_flag)
#define CUADMM_SPMV_CASE(G)                                                                            \
    case G:                                                                                            \
        switch (epi.mode) {                                                                            \
            case 0: CUADMM_SPMV_LAUNCH(G, 0); break;                                                   \
            case 1: CUADMM_SPMV_LAUNCH(G, 1); break;                                                   \
            case 2: CUADMM_SPMV_LAUNCH(G, 2); break;                                                   \
            case 3: CUADMM_SPMV_LAUNCH(G, 3); break;                                                   \
            case 4: CUADMM_SPMV_LAUNCH(G, 4); break;                                                   \
            case 5: CUADMM_SPMV_LAUNCH(G, 5); break;                                                   \
            default: throw Error(CUADMM_EINVAL, "bad spmv epilogue mode");                             \
        }                                                                                              \
        break;
    switch (A.group) {
        CUADMM_SPMV_CASE(1)
        CUADMM_SPMV_CASE(2)
        CUADMM_SPMV_CASE(4)
        CUADMM_SPMV_CASE(8)
        CUADMM_SPMV_CASE(16)
        CUADMM_SPMV_CASE(32)
        default: throw Error(CUADMM_EINVAL, "bad spmv group size");
    }
#undef CUADMM_SPMV_LAUNCH
#undef CUADMM_SPMV_CASE
    CUADMM_CUDA(cudaGetLastError());
}

cuadmm_spmv_s* spmv_create(int64_t rows, int64_t cols, int64_t nnz, const int32_t* h_rowptr,
                         const int32_t* h_colind, const double* h_val, int device) {
    CUADMM_REQUIRE(rows >= 0 && cols >= 0 && nnz >= 0, "negative dimension");
    CUADMM_REQUIRE(nnz <= INT32_MAX, "nnz exceeds int32 row pointers");
    CUADMM_REQUIRE(h_rowptr != nullptr, "rowptr is null");
    CUADMM_REQUIRE(h_rowptr[0] == 0 && h_rowptr[rows] == nnz, "rowptr does not span nnz");
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) {
        cudaGetLastError();
        throw Error(CUADMM_ENODEVICE, "no CUDA device available; SpMV has no CPU fallback");
    }
    CUADMM_REQUIRE(device >= 0 && device < cnt, "device index out of range");
    std::unique_ptr<cuadmm_spmv_full> A(new cuadmm_spmv_full());
    A->device = device; A->rows = rows; A->cols = cols; A->nnz = nnz;
    DeviceGuard g(device);
    A->rowptr.alloc(rows + 1); A->rowptr.upload(h_rowptr, rows + 1);
    A->colind.alloc(std::max<int64_t>(nnz, 1)); A->colind.upload(h_colind, nnz);
    A->val.alloc(std::max<int64_t>(nnz, 1)); A->val.upload(h_val, nnz);
    std::vector<int32_t> longs;
    int64_t short_nnz = 0, short_rows = 0;
    for (int64_t i = 0; i < rows; ++i) {
        CUADMM_REQUIRE(h_rowptr[i + 1] >= h_rowptr[i], "rowptr not monotone");
        const int len = h_rowptr[i + 1] - h_rowptr[i];
        if (len > 0 && len <= kLongRow) { short_nnz += len; ++short_rows; }
    }
    for (int64_t p = 0; p < nnz; ++p) CUADMM_REQUIRE(h_colind[p] >= 0 && h_colind[p] < cols, "column index out of range");
    const double mean = short_rows ? (double)short_nnz / (double)short_rows : 1.0;
    int G = 1;
    while (G < 32 && (double)G < mean) G <<= 1;   // smallest power of two >= mean row length
    // measured on B200 (profiles/spmv_probe_r01.json, the bench operator, 2.2 and 1.5 entries per row): one lane per
    // row beats 2 or 4 lanes (12.4 us vs 20.4 us for A) — with rows this short the sub-warp reduction and the idle
    // lanes of the epilogue cost more than the uncoalesced value/column loads.  Only for matrices large enough to be
    // throughput-bound (several rows per resident thread); small ones (PushT: 28k rows) are latency-bound and keep
    // the lanes of a row working in parallel (measured: 0.405 vs 0.446 ms per iteration)
    if (mean <= 4.0 && rows >= 262144) G = 1;
    if (const char* e = getenv("CUADMM_SPMV_G")) {   // tuning override: 1, 2, 4, ... 32
        const int g = atoi(e);
        if (g >= 1 && g <= 32 && (g & (g - 1)) == 0) G = g;
    }
    A->group = G;
    // one lane per row: a row of k entries is a chain of ~k/4 dependent load pairs for that lane while the 31 others idle,
    // and the slowest lane of the last wave is the tail of the kernel (At of the bench operator: mean 1.1, max 41) — rows
    // beyond 16 entries go to the warp-per-row part there.  Smaller matrices keep the round-1 rule (and summation order).
    int long_thr = kLongRow;
    if (G == 1 && rows >= 262144) long_thr = 16;
    if (const char* e = getenv("CUADMM_SPMV_LONG")) { const int t = atoi(e); if (t >= 1 && t <= kLongRow) long_thr = t; }
    A->aux.long_thr = long_thr;
    for (int64_t i = 0; i < rows; ++i)
        if (h_rowptr[i + 1] - h_rowptr[i] > long_thr) longs.push_back((int32_t)i);
    A->aux.n_long = (int)longs.size();
    if (!longs.empty()) { A->aux.long_rows.upload(longs); }
    int sm = 148;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device);
    const int64_t need = (rows * G + kThreads - 1) / kThreads;
    // one resident wave: the grid is capped at the CTAs per SM the kernel can keep resident (5 at 48 registers);
    // 8 per SM made 1.6 waves and cost 20-25 % (same probe).  CUADMM_SPMV_CAP overrides.
    const int cap_default = CUADMM_SPMV_MINB;     // every instantiation is compiled for this many resident CTAs per SM
    int cap = cap_default;
    if (const char* e = getenv("CUADMM_SPMV_CAP")) { const int c = atoi(e); if (c >= 1 && c <= 32) cap = c; }
    A->aux.main_blocks = (int)std::min<int64_t>(std::max<int64_t>(need, 1), (int64_t)sm * cap);
    A->aux.long_blocks = longs.empty() ? 0 : (int)std::min<int64_t>(((int64_t)longs.size() + 7) / 8, (int64_t)sm);
    CUADMM_CUDA(cudaStreamSynchronize(0));
    return A.release();
}

}  // namespace cuadmm

extern "C" {

int cuadmm_spmv_create(int64_t rows, int64_t cols, int64_t nnz, const int32_t* h_rowptr, const int32_t* h_colind,
                       const double* h_val, int device, cuadmm_spmv_t** out) {
    return guarded([&] {
        CUADMM_REQUIRE(out != nullptr, "out is null");
        *out = nullptr;
        *out = spmv_create(rows, cols, nnz, h_rowptr, h_colind, h_val, device);
    });
}

void cuadmm_spmv_destroy(cuadmm_spmv_t* A) { delete static_cast<cuadmm_spmv_full*>(A); }

int cuadmm_spmv(cuadmm_spmv_t* A, double alpha, const double* d_x, double beta, double* d_y, void* stream) {
    return guarded([&] {
        CUADMM_REQUIRE(A && d_x && d_y, "null argument");
        DeviceGuard g(A->device);
        SpmvEpilogue e;
        spmv_launch(*A, alpha, d_x, beta, d_y, e, (cudaStream_t)stream);
    });
}

int cuadmm_spmv_host(cuadmm_spmv_t* A, double alpha, const double* h_x, double beta, double* h_y) {
    return guarded([&] {
        CUADMM_REQUIRE(A && h_x && h_y, "null argument");
        DeviceGuard g(A->device);
        if (A->d_x.n != A->cols) A->d_x.alloc(A->cols);
        if (A->d_y.n != A->rows) A->d_y.alloc(A->rows);
        A->d_x.upload(h_x, A->cols);
        A->d_y.upload(h_y, A->rows);
        SpmvEpilogue e;
        spmv_launch(*A, alpha, A->d_x.p, beta, A->d_y.p, e, 0);
        A->d_y.download(h_y, A->rows);
        CUADMM_CUDA(cudaStreamSynchronize(0));
    });
}

// get_normA on host arrays (src/kernels/sparse_matrix_norm.cu:11-31): serial sum per constraint,
// floor 1.0, in-place division.  Init-time; the solver runs the same arithmetic.
int cuadmm_normA_host(int64_t con_num, const int32_t* At_col_ptrs, double* At_vals, double* normA) {
    return guarded([&] {
        CUADMM_REQUIRE(At_col_ptrs && At_vals && normA, "null argument");
        for (int64_t i = 0; i < con_num; ++i) {
            double norm = 0.0;
            for (int p = At_col_ptrs[i]; p < At_col_ptrs[i + 1]; ++p) norm += At_vals[p] * At_vals[p];
            norm = std::max(1.0, sqrt(norm));
            normA[i] = norm;
            for (int p = At_col_ptrs[i]; p < At_col_ptrs[i + 1]; ++p) At_vals[p] /= norm;
        }
    });
}

// CSC -> CSR of the same matrix (cusparseCsr2cscEx2 in the reference, include/cuadmm/cusparse.h:35-49):
// counting sort, row-major output with ascending column ids inside a row.
int cuadmm_csc_to_csr_host(int64_t nrows, int64_t ncols, int64_t nnz, const int32_t* col_ptrs, const int32_t* row_ids,
                           const double* vals, int32_t* row_ptrs, int32_t* col_ids, double* out_vals) {
    return guarded([&] {
        CUADMM_REQUIRE(col_ptrs && row_ptrs, "null argument");
        std::vector<int64_t> cnt(nrows + 1, 0);
        for (int64_t p = 0; p < nnz; ++p) {
            CUADMM_REQUIRE(row_ids[p] >= 0 && row_ids[p] < nrows, "row index out of range");
            cnt[row_ids[p] + 1]++;
        }
        for (int64_t i = 0; i < nrows; ++i) cnt[i + 1] += cnt[i];
        for (int64_t i = 0; i <= nrows; ++i) row_ptrs[i] = (int32_t)cnt[i];
        std::vector<int64_t> next(cnt.begin(), cnt.end() - 1);
        for (int64_t c = 0; c < ncols; ++c)
            for (int p = col_ptrs[c]; p < col_ptrs[c + 1]; ++p) {
                const int64_t q = next[row_ids[p]]++;
                col_ids[q] = (int32_t)c;
                out_vals[q] = vals[p];
            }
    });
}

}  // extern "C"
