"""ctypes access to oracle/_ref/libref_rlap.so — the UNMODIFIED reference C++ path
(rlap/csrc/{factorizers,reader,preconditioner}.cc) compiled against the container-only
Eigen stand-in. TEST INFRASTRUCTURE ONLY: may be imported from tests/, bench.py's
cpu_baseline / --impl reference leg and __graft_entry__.smoke() — never from rlap_b200/.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libref_rlap.so")
_lib = None


def available() -> bool:
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(_LIB_PATH)
        L.ref_set_seeds.argtypes = [ctypes.c_uint64, ctypes.c_int, ctypes.c_uint64]
        L.ref_set_seeds.restype = None
        L.ref_approximate_cholesky.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
            ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.POINTER(ctypes.c_double))]
        L.ref_approximate_cholesky.restype = ctypes.c_int64
        L.ref_free.argtypes = [ctypes.POINTER(ctypes.c_double)]
        L.ref_free.restype = None
        _lib = L
    return _lib


def approximate_cholesky(edge_info: np.ndarray, num_nodes: int, num_remove: int, o_v: str, o_n: str,
                         sample_seed: int = 5489, rd_seed=None) -> np.ndarray:
    """edge_info: [E,3] float64 (row, col, weight) as rlap/ops.py:47 packs it.
    rd_seed=None keeps the reference's real std::random_device; an int injects a
    reproducible stream for the permutation / o_n="random" shuffles."""
    ei = np.ascontiguousarray(edge_info, dtype=np.float64)
    assert ei.ndim == 2 and ei.shape[1] == 3
    L = lib()
    L.ref_set_seeds(sample_seed, 0 if rd_seed is None else 1, 0 if rd_seed is None else int(rd_seed))
    out = ctypes.POINTER(ctypes.c_double)()
    rows = L.ref_approximate_cholesky(ei.ctypes.data, ei.shape[0], num_nodes, num_remove,
                                      o_v.encode(), o_n.encode(), ctypes.byref(out))
    res = np.ctypeslib.as_array(out, shape=(max(rows, 1), 3))[:rows].copy()
    L.ref_free(out)
    return res
