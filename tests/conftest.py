import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_built():
    lib = os.path.join(ROOT, "cuadmm_b200", "lib", "libcuadmm_b200.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-C", ROOT, "-j8"], stdout=subprocess.DEVNULL)
    ohost = os.path.join(ROOT, "oracle", "_build", "liboracle_host.so")
    if not os.path.exists(ohost) or not os.path.exists(os.path.join(ROOT, "oracle", "_build", "libcpu_baseline.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)


_ensure_built()

i32p = C.POINTER(C.c_int)
f64p = C.POINTER(C.c_double)


def ip(a):
    return a.ctypes.data_as(i32p)


def dp(a):
    return a.ctypes.data_as(f64p)


@pytest.fixture(scope="session")
def ohost():
    """the plain-C oracle (oracle/oracle_host.c)"""
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "liboracle_host.so"))
    lib.oracle_sqrt2.restype = C.c_double
    return lib


@pytest.fixture(scope="session")
def oref():
    """the unmodified reference compiled into oracle/_ref (host entry points only on CPU)"""
    path = os.path.join(ROOT, "oracle", "_ref", "libcuadmm_ref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built (no /root/reference here and no prebuilt copy)")
    try:
        lib = C.CDLL(path)
    except OSError as e:
        pytest.skip(f"oracle/_ref not loadable: {e}")
    lib.ref_sqrt2.restype = C.c_double
    lib.ref_sqrt2inv.restype = C.c_double
    lib.ref_proj_create.restype = C.c_void_p
    lib.ref_proj_run.restype = C.c_double
    return lib


def oracle_maps(ohost, blk):
    blk = np.ascontiguousarray(blk, np.int32)
    L = int(sum(int(n) * (int(n) + 1) // 2 for n in blk))
    B = np.zeros(L, np.int32); M1 = np.zeros(L, np.int32); M2 = np.zeros(L, np.int32)
    ohost.oracle_get_maps(ip(blk), len(blk), ip(B), ip(M1), ip(M2))
    return B, M1, M2


def random_svec(blk, seed, scale=1.0):
    """random symmetric blocks (G + G^T)/2, G ~ N(0,1), as one svec vector (SURVEY 8d inputs)"""
    import oracle_np as onp
    rng = np.random.default_rng(seed)
    parts = []
    for n in blk:
        G = rng.standard_normal((int(n), int(n)))
        parts.append(onp.svec((G + G.T) / 2 * scale))
    return np.concatenate(parts) if parts else np.zeros(0)


def has_gpu():
    import cuadmm_b200
    return cuadmm_b200.device_count() > 0
