"""cuadmm_b200 — Python binding (ctypes) over the C ABI in include/cuadmm_b200.h.

The product is the C++/CUDA library ``cuadmm_b200/lib/libcuadmm_b200.so`` (sm_100a only) and the
``cuadmm_exe`` front end; this package only loads the library for tests and the bench harness.
There is no Python/CPU implementation of any compute path: if the library is missing the import
fails loudly, and compute entry points fail with CUADMM_ENODEVICE when no GPU is present.
"""
from .capi import (  # noqa: F401
    lib, LIB_PATH, CuadmmError, Plan, SpMV, YSolve, Solver, Problem, device_count, version,
    normA_host, csc_to_csr_host, Shard, nccl_unique_id, unique_id,
)
