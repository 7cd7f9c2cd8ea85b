// problem.h — host-side SDP problem in the reference's TXT conventions (include/cuadmm/problem.h).
#pragma once
#include "common.h"

namespace cuadmm {

struct Problem {
    std::vector<double> X_vals, y_vals, S_vals;              // optional warm start
    std::vector<int32_t> At_csc_col_ptrs, At_csc_row_ids;    // At (vec_len x con_num) in CSC
    std::vector<double> At_csc_vals;
    int64_t At_nnz = 0;
    std::vector<int32_t> b_indices; std::vector<double> b_vals; int64_t b_nnz = 0;
    std::vector<int32_t> C_indices; std::vector<double> C_vals; int64_t C_nnz = 0;
    std::vector<char> blk_types; std::vector<int32_t> blk_vals;
    int64_t vec_len = 0, mat_num = 0, con_num = 0;
    std::vector<std::string> warnings;
    void from_txt(const std::string& prefix, bool warm_start = false);
};

// sort COO triplets by (col,row) into CSC (src/utils/io.cu:187-243)
void coo_to_csc(int64_t ncols, const std::vector<int32_t>& rows, const std::vector<int32_t>& cols,
                const std::vector<double>& vals, std::vector<int32_t>& col_ptrs, std::vector<int32_t>& row_ids,
                std::vector<double>& out_vals);

}  // namespace cuadmm

struct cuadmm_problem {
    cuadmm::Problem prob;
};
