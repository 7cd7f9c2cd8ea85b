// plan.h — the projection plan: host block layout + device descriptors + launch classes.
#pragma once
#include "common.h"
#include "blocks.h"
#include "project_jacobi.cuh"

namespace cuadmm { struct DensePart; }
cuadmm::DensePart* dense_part_create(int device, const std::vector<int32_t>& blk, const std::vector<int64_t>& svec_off,
                                     const std::vector<int64_t>& which);
void dense_part_destroy(cuadmm::DensePart* p);
int dense_part_project(cuadmm::DensePart* D, const double* Xb, double* Xproj, cudaStream_t st,
                       const cuadmm::ProjEpilogue* epi, const int* done_flag);

struct cuadmm_plan {
    int device = -1;
    cuadmm::BlockLayout layout;

    // launch classes: blocks grouped by size range, each class = one kernel launch
    struct Class {
        int kind;          // index into the kernel table
        int nmax;          // largest n in the class
        int32_t count;     // blocks in the class
        int64_t first;     // first descriptor (in d_desc)
        size_t smem;       // dynamic shared memory per CTA
        int grid;
    };
    std::vector<Class> classes;
    std::vector<cuadmm::BlkDesc> h_desc;       // class-major, n descending inside a class
    cuadmm::DevBuf<cuadmm::BlkDesc> d_desc;
    cuadmm::DevBuf<double> d_scratch;          // global-memory Jacobi variant / large path
    cuadmm::DevBuf<double> d_eig;              // debug eigenvalue output (sum n)
    cuadmm::DevBuf<int32_t> d_sweeps;
    cuadmm::DevBuf<double> d_in, d_out;        // staging for the *_host entry points
    // pooled-layout descriptors for svec<->smat
    cuadmm::DevBuf<int64_t> d_svec_off, d_mat_off;
    cuadmm::DevBuf<int32_t> d_blk;
    cuadmm::DevBuf<uint8_t> d_pool;

    // warm start of the Jacobi kernels: per block the orthonormal basis the last projection ended in
    cuadmm::DevBuf<double> d_Q;
    bool warm_start = true;
    void reset_warm_start();
    double threshold = 1e-11;     // columns of G count as orthogonal when every |cos| <= threshold (tested on the state after each sweep)
    int max_sweeps = 40;
    bool force_global = false;    // every block above the shared-memory classes on the global-memory Jacobi kernel (eigenvalue debug plan)
    std::unique_ptr<cuadmm_plan> eig_plan;   // lazily built shadow plan: eigenvalues of the blocks on the dense sign path (debug entry only)
    int rank_limit = 0;           // > 0: fixed-rank projection (cuadmm_plan_set_rank_limit)
    bool use_gram = true;         // CUADMM_JACOBI_GRAM=0: round-1 stopping rule (a whole sweep without a cosine above the threshold)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaStream_t> side_streams;    // classes run concurrently on these
    std::vector<cudaEvent_t> side_events;
    cudaEvent_t fork_event = nullptr;
    cuadmm::DensePart* dense = nullptr;        // blocks n > 168: GEMM-only sign-function path (dense_proj.cu)
    const int* done_flag = nullptr;            // device stop flag honoured by the kernels (solver)
    double last_ms = 0.0;
    int64_t last_launches = 0;

    ~cuadmm_plan();
    void build_device();
    // core launcher; epi may be null.  Returns the number of kernel launches.
    int project(const double* d_Xb, double* d_Xproj, cudaStream_t stream,
                const cuadmm::ProjEpilogue* epi, bool want_eig);
};
