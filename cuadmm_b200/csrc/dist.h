// dist.h — NCCL plumbing for the sharded solver.  libnccl is dlopen'ed (no link-time dependency): in a
// torchrun process this resolves to the NCCL torch already loaded, in cuadmm_exe to the system one.
#pragma once
#include "common.h"

namespace cuadmm {

struct NcclComm {
    void* comm = nullptr;
    int rank = 0, world = 1;
    ~NcclComm();
    void init(int rank, int world, const char id[128], int device);
    // sum all-reduce of `count` doubles on `stream` (send == recv: in place)
    void allreduce_sum(const double* send, double* recv, int64_t count, cudaStream_t stream);
};

void nccl_unique_id(char out[128]);

}  // namespace cuadmm
