"""cuadmm_b200 — Python binding (ctypes) over the C ABI in include/cuadmm_b200.h.

The product is the C++/CUDA library ``cuadmm_b200/lib/libcuadmm_b200.so`` (sm_100a only) and the
``cuadmm_exe`` front end; this package only loads the library for tests and the bench harness.
There is no Python/CPU implementation of any compute path: if the library is missing the import
fails loudly, and compute entry points fail with CUADMM_ENODEVICE when no GPU is present.

The shared library is mapped on first use of a binding (``cuadmm_b200.Solver`` etc.), not when a
data-only submodule such as ``cuadmm_b200.synthetic`` is imported — the CPU reference arm of bench.py
generates its workload without ever loading the CUDA library.
"""
import os as _os

LIB_PATH = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "lib", "libcuadmm_b200.so")
if not _os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `make` at the repo root (or "
        "`python -c 'import __graft_entry__ as g; g.build()'`). There is no pure-Python fallback.")

_API = ("lib", "CuadmmError", "Plan", "SpMV", "YSolve", "Solver", "Problem", "device_count", "version",
        "normA_host", "csc_to_csr_host", "Shard", "nccl_unique_id", "unique_id", "solve_matlab_like", "eig_rank_mask")


def __getattr__(name):
    if name in _API:
        from . import capi
        return getattr(capi, name)
    raise AttributeError(f"module 'cuadmm_b200' has no attribute {name!r}")


def __dir__():
    return sorted(list(globals()) + list(_API))
