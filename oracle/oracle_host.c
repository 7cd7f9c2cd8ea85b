/*
 * oracle_host.c — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the reference's host-side integer logic on the hot path, each
 * function citing the reference lines it follows.  Pinned against the reference's own
 * golden vectors (tests/test_oracle_golden.py) and, where /root/reference is present,
 * against the reference sources compiled unmodified into oracle/_ref (oracle/Makefile).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this.
 *
 * int is 32-bit like the reference's `int`; overflow behaviour is deliberately the same.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* src/matrix_sizes.cu:14-19 */
int oracle_is_large_mat(int mat_size, int mat_num) {
    if (mat_size > 32) return 1;
    return ((double)mat_size - 17.0 > (double)mat_num * 1.4) ? 1 : 0;
}

static int cmp_int(const void* a, const void* b) {
    int x = *(const int*)a, y = *(const int*)b;
    return (x > y) - (x < y);
}

/* src/utils/analyze_blk.cu:63-99 — distinct sizes ascending (std::set order) and counts.
 * Returns the number of distinct sizes; blk_sizes / blk_nums must hold nblk entries. */
int oracle_analyze_blk(const int* blk, int nblk, int* blk_sizes, int* blk_nums) {
    int* tmp = (int*)malloc(sizeof(int) * (size_t)(nblk > 0 ? nblk : 1));
    memcpy(tmp, blk, sizeof(int) * (size_t)nblk);
    qsort(tmp, (size_t)nblk, sizeof(int), cmp_int);
    int ns = 0;
    for (int i = 0; i < nblk; ++i)
        if (i == 0 || tmp[i] != tmp[i - 1]) blk_sizes[ns++] = tmp[i];
    free(tmp);
    for (int j = 0; j < ns; ++j) blk_nums[j] = 0;
    for (int i = 0; i < nblk; ++i)
        for (int j = 0; j < ns; ++j)
            if (blk[i] == blk_sizes[j]) blk_nums[j] += 1;
    return ns;
}

/* src/utils/analyze_blk.cu:22-60 — returns 0 on success, -1 when #sizes != 2 (reference asserts) */
int oracle_analyze_blk_duo(const int* blk, int nblk, int* LARGE, int* SMALL, int* mom_mat_num, int* loc_mat_num) {
    int mn = blk[0], mx = blk[0], types = 0;
    int* s = (int*)malloc(sizeof(int) * (size_t)nblk);
    int* c = (int*)malloc(sizeof(int) * (size_t)nblk);
    types = oracle_analyze_blk(blk, nblk, s, c);
    free(s); free(c);
    for (int i = 0; i < nblk; ++i) { if (blk[i] < mn) mn = blk[i]; if (blk[i] > mx) mx = blk[i]; }
    if (types != 2) return -1;
    *LARGE = mx; *SMALL = mn; *mom_mat_num = 0; *loc_mat_num = 0;
    for (int i = 0; i < nblk; ++i) { if (blk[i] == mx) *mom_mat_num += 1; else *loc_mat_num += 1; }
    return 0;
}

/* MatrixSizes::init, src/matrix_sizes.cu:22-68.  Output arrays sized for ns(+1) entries.
 * totals[6] = large_mat_num, sum_large_mat_size, total_large_mat_size, small_mat_num,
 *             sum_small_mat_size, total_small_mat_size.  Returns nl | (nsm << 16). */
int oracle_matrix_sizes(const int* blk_sizes, const int* blk_nums, int ns,
                        int* large_sizes, int* large_nums, int* large_mat_start, int* large_W_start,
                        int* small_sizes, int* small_nums, int* small_mat_start, int* small_W_start,
                        int* totals) {
    int nl = 0, nsm = 0;
    int large_mat_num = 0, sum_large = 0, total_large = 0, small_mat_num = 0, sum_small = 0, total_small = 0;
    large_mat_start[0] = 0; large_W_start[0] = 0; small_mat_start[0] = 0; small_W_start[0] = 0;
    for (int i = 0; i < ns; ++i) {
        int mat_size = blk_sizes[i], mat_num = blk_nums[i];
        if (oracle_is_large_mat(mat_size, mat_num)) {
            large_mat_num += mat_num;
            sum_large += mat_size * mat_num;
            total_large += mat_num * mat_size * mat_size;
            large_sizes[nl] = mat_size; large_nums[nl] = mat_num; ++nl;
            large_mat_start[nl] = total_large; large_W_start[nl] = sum_large;
        } else {
            sum_small += mat_size * mat_num;
            small_mat_num += mat_num;
            total_small += mat_num * mat_size * mat_size;
            small_sizes[nsm] = mat_size; small_nums[nsm] = mat_num; ++nsm;
            small_mat_start[nsm] = total_small; small_W_start[nsm] = sum_small;
        }
    }
    totals[0] = large_mat_num; totals[1] = sum_large; totals[2] = total_large;
    totals[3] = small_mat_num; totals[4] = sum_small; totals[5] = total_small;
    return nl | (nsm << 16);
}

/* get_maps, src/utils/get_maps.cu:80-134 (offset getters src/matrix_sizes.cu:116-150) */
void oracle_get_maps(const int* blk, int nblk, int* map_B, int* map_M1, int* map_M2) {
    int* sizes = (int*)malloc(sizeof(int) * (size_t)(nblk + 1));
    int* nums = (int*)malloc(sizeof(int) * (size_t)(nblk + 1));
    int ns = oracle_analyze_blk(blk, nblk, sizes, nums);
    int* ls = (int*)malloc(sizeof(int) * (size_t)(ns + 1)); int* ln = (int*)malloc(sizeof(int) * (size_t)(ns + 1));
    int* lm = (int*)malloc(sizeof(int) * (size_t)(ns + 2)); int* lw = (int*)malloc(sizeof(int) * (size_t)(ns + 2));
    int* ss = (int*)malloc(sizeof(int) * (size_t)(ns + 1)); int* sn = (int*)malloc(sizeof(int) * (size_t)(ns + 1));
    int* sm = (int*)malloc(sizeof(int) * (size_t)(ns + 2)); int* sw = (int*)malloc(sizeof(int) * (size_t)(ns + 2));
    int totals[6];
    int packed = oracle_matrix_sizes(sizes, nums, ns, ls, ln, lm, lw, ss, sn, sm, sw, totals);
    int nl = packed & 0xffff, nsm = packed >> 16;
    int* lenc = (int*)calloc((size_t)(nl + 1), sizeof(int));
    int* senc = (int*)calloc((size_t)(nsm + 1), sizeof(int));
    int idx = 0;
    for (int k = 0; k < nblk; ++k) {
        int s = blk[k], b, gi = 0, same, off;
        int is_large = 0;
        for (int j = 0; j < nl; ++j) if (ls[j] == s) { is_large = 1; gi = j; break; }
        if (is_large) {
            b = 0; same = lenc[gi]++;
            off = lm[gi] + same * s * s;          /* large_mat_offset */
        } else {
            for (int j = 0; j < nsm; ++j) if (ss[j] == s) { gi = j; break; }
            b = 1; same = senc[gi]++;
            off = sm[gi] + same * s * s;          /* small_mat_offset */
        }
        for (int i = 1; i <= s; ++i)
            for (int j = 1; j <= i; ++j) {
                map_B[idx] = b;
                map_M1[idx] = off + s * (i - 1) + j - 1;   /* count horizontally */
                map_M2[idx] = off + s * (j - 1) + i - 1;   /* count vertically */
                ++idx;
            }
    }
    free(sizes); free(nums); free(ls); free(ln); free(lm); free(lw); free(ss); free(sn); free(sm); free(sw);
    free(lenc); free(senc);
}

/* get_maps_duo, src/utils/get_maps.cu:22-68 */
void oracle_get_maps_duo(const int* blk, int nblk, int LARGE, int* map_B, int* map_M1, int* map_M2) {
    int idx = 0, k_Xmom = 0, k_Xloc = 0;
    for (int k = 0; k < nblk; ++k) {
        int s = blk[k], b;
        if (s == LARGE) { b = 0; ++k_Xmom; } else { b = 1; ++k_Xloc; }
        for (int i = 1; i <= s; ++i)
            for (int j = 1; j <= i; ++j) {
                map_B[idx] = b;
                int kk = (s == LARGE) ? k_Xmom : k_Xloc;
                map_M1[idx] = s * s * (kk - 1) + s * (i - 1) + j - 1;
                map_M2[idx] = s * s * (kk - 1) + s * (j - 1) + i - 1;
                ++idx;
            }
    }
}

/* vector_to_matrices_kernel / matrices_to_vector_kernel, src/kernels/vec_mat_conversion.cu:11-57.
 * SQRT2: include/cuadmm/kernels.h:173-181 (Newton fixed point), SQRT2INV = 1.0/SQRT2. */
static double sqrt_newton(double x, double curr, double prev) {
    while (curr != prev) { double next = 0.5 * (curr + x / curr); prev = curr; curr = next; }
    return curr;
}
double oracle_sqrt2(void) { return sqrt_newton(2.0, 2.0, 0.0); }

void oracle_vector_to_matrices(const double* Xb, double* mom_mat, double* loc_mat,
                               const int* map_B, const int* map_M1, const int* map_M2, int vec_len) {
    const double SQRT2 = oracle_sqrt2(), SQRT2INV = 1.0 / SQRT2;
    for (int idx = 0; idx < vec_len; ++idx) {
        int b = map_B[idx], m1 = map_M1[idx], m2 = map_M2[idx];
        int if_diag = (m1 == m2);
        double* M = (b == 0) ? mom_mat : loc_mat;
        M[m1] = Xb[idx] * (SQRT2INV + (double)if_diag * (1 - SQRT2INV));
        M[m2] = M[m1];
    }
}

void oracle_matrices_to_vector(double* Xb, const double* mom_mat, const double* loc_mat,
                               const int* map_B, const int* map_M1, const int* map_M2, int vec_len) {
    const double SQRT2 = oracle_sqrt2();
    for (int idx = 0; idx < vec_len; ++idx) {
        int b = map_B[idx], m1 = map_M1[idx], m2 = map_M2[idx];
        int if_diag = (m1 == m2);
        const double* M = (b == 0) ? mom_mat : loc_mat;
        Xb[idx] = M[m1] * (SQRT2 + (double)if_diag * (1 - SQRT2));
    }
}

/* dense_matrix_mul_diag_batch_kernel, src/kernels/diagonal_batch.cu:11-22 */
void oracle_mul_diag_batch(double* out, const double* in, const double* vec, int mat_size, int total_len) {
    for (int idx = 0; idx < total_len; ++idx) {
        int k = idx / (mat_size * mat_size);
        int i = (idx % (mat_size * mat_size)) / mat_size;
        out[idx] = in[idx] * vec[k * mat_size + i];
    }
}

/* get_normA_kernel, src/kernels/sparse_matrix_norm.cu:11-31 (serial sum in column order,
 * floor 1.0, in-place division) */
void oracle_get_normA(const int* At_col_ptrs, double* At_vals, double* normA, int con_num) {
    for (int idx = 0; idx < con_num; ++idx) {
        double norm = 0.0;
        for (int i = At_col_ptrs[idx]; i < At_col_ptrs[idx + 1]; ++i) norm += At_vals[i] * At_vals[i];
        norm = fmax(1.0, sqrt(norm));
        normA[idx] = norm;
        for (int i = At_col_ptrs[idx]; i < At_col_ptrs[idx + 1]; ++i) At_vals[i] /= norm;
    }
}

/* perform_permutation_kernel, src/kernels/permutation.cu:12-17: a scatter */
void oracle_perform_permutation(double* vec1, const double* vec2, const int* perm, int size) {
    for (int i = 0; i < size; ++i) vec1[perm[i]] = vec2[i];
}

/* get_inverse_permutation, src/utils/inverse_permutation.cu:17-30 (argsort of perm) */
void oracle_inverse_permutation(int* perm_inv, const int* perm, int size) {
    for (int i = 0; i < size; ++i) perm_inv[perm[i]] = i;
}

/* COO_to_CSC, src/utils/io.cu:187-243: sort by (col,row); col_ptrs filled by the reference's
 * scan (col_ptrs[0] stays 0; trailing pointers = nnz).  Stable merge sort on (col,row). */
typedef struct { int col, row; double val; } trip_t;
static int cmp_trip(const void* a, const void* b) {
    const trip_t* x = (const trip_t*)a; const trip_t* y = (const trip_t*)b;
    if (x->col != y->col) return (x->col > y->col) - (x->col < y->col);
    return (x->row > y->row) - (x->row < y->row);
}
void oracle_coo_to_csc(int* col_ptrs, int* col_ids, int* row_ids, double* vals, int nnz, int col_num) {
    trip_t* t = (trip_t*)malloc(sizeof(trip_t) * (size_t)(nnz > 0 ? nnz : 1));
    for (int i = 0; i < nnz; ++i) { t[i].col = col_ids[i]; t[i].row = row_ids[i]; t[i].val = vals[i]; }
    qsort(t, (size_t)nnz, sizeof(trip_t), cmp_trip);
    for (int i = 0; i <= col_num; ++i) col_ptrs[i] = 0;
    int id = 0;
    for (int i = 1; i < nnz; ++i) {
        if (t[i - 1].col < t[i].col) {
            int tmp = t[i - 1].col;
            while (tmp < t[i].col) { id = id + 1; col_ptrs[id] = i; tmp = tmp + 1; }
        }
    }
    id = id + 1;
    while (id <= col_num) { col_ptrs[id] = nnz; id = id + 1; }
    for (int i = 0; i < nnz; ++i) { col_ids[i] = t[i].col; row_ids[i] = t[i].row; vals[i] = t[i].val; }
    free(t);
}

/* y = alpha*A*x + beta*y for CSR (semantics of SpMV_cusparse, include/cuadmm/cusparse.h:70-83),
 * serial row sums in column order */
void oracle_spmv_csr(int rows, const int* rowptr, const int* colind, const double* val,
                     double alpha, const double* x, double beta, double* y) {
    for (int i = 0; i < rows; ++i) {
        double acc = 0.0;
        for (int p = rowptr[i]; p < rowptr[i + 1]; ++p) acc += val[p] * x[colind[p]];
        y[i] = alpha * acc + (beta == 0.0 ? 0.0 : beta * y[i]);
    }
}

/* get_eig_rank_mask (src/utils/get_eig_rank_mask.cu:16-38): zero everywhere, 1 on the last eig_rank entries of every
 * block of mat_size (eigenvalues are ascending, so those are the eig_rank largest) */
void oracle_eig_rank_mask(int* mask, int batch_size, int mat_size, int eig_rank) {
    for (int i = 0; i < batch_size * mat_size; ++i) mask[i] = 0;
    for (int i = 0; i < batch_size; ++i)
        for (int j = 0; j < eig_rank; ++j) mask[i * mat_size + (mat_size - 1 - j)] = 1;
}
