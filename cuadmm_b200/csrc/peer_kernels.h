// peer_kernels.h — launchers of the peer-memory kernels (peer.cu).
#pragma once
#include "peer.h"

namespace cuadmm {

int peer_reduce_grid(const PeerComm& pc, int64_t slice);
// mode 0: out = sum over ranks of the staged partial rows; mode 1: out = b - sum plus the residual scalars
void peer_reduce_launch(const PeerComm& pc, int mode, int64_t count, int64_t slice, const double* stage, const PeerPtrs& out,
                        const double* b, const double* normA, const double* y, const double* part_rd, int n_rd, double* cta_part,
                        const int* done_flag, cudaStream_t stream);
void peer_scatter_full_launch(const PeerComm& pc, int64_t nloc, const double* local, const int64_t* loc2glob, const PeerPtrs& full,
                              cudaStream_t stream);

}  // namespace cuadmm
