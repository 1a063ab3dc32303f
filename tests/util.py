"""Shared helpers for the test-suite (graphs, dense Laplacians, exact Schur complements)."""
import numpy as np

from rlap_b200 import graphs

COMBOS = [(ov, on) for ov in ("random", "degree", "coarsen") for on in ("asc", "desc", "random")]


def sym_weights(ei: np.ndarray, lo=0.5, hi=1.5) -> np.ndarray:
    """deterministic symmetric weights in [lo, hi): a hash of the unordered pair"""
    a, b = np.minimum(ei[0], ei[1]), np.maximum(ei[0], ei[1])
    h = (a * 7919 + b * 104729) % 1000
    return (lo + (hi - lo) * h / 1000.0).astype(np.float32)


def edge_info(ei: np.ndarray, w=None) -> np.ndarray:
    E = ei.shape[1]
    w = np.ones(E) if w is None else np.asarray(w, dtype=np.float64)
    return np.concatenate([ei.T.astype(np.float64), w.reshape(-1, 1)], axis=1)


def laplacian(rows, cols, w, n) -> np.ndarray:
    A = np.zeros((n, n))
    np.add.at(A, (np.asarray(rows, dtype=np.int64), np.asarray(cols, dtype=np.int64)), np.asarray(w, dtype=np.float64))
    return np.diag(A.sum(0)) - A


def exact_schur(L: np.ndarray, keep: np.ndarray) -> np.ndarray:
    n = L.shape[0]
    F = np.setdiff1d(np.arange(n), keep)
    if F.size == 0:
        return L[np.ix_(keep, keep)]
    return L[np.ix_(keep, keep)] - L[np.ix_(keep, F)] @ np.linalg.solve(L[np.ix_(F, F)], L[np.ix_(F, keep)])


def small_cases():
    """(name, edge_index, n, graph_ptr or None, num_remove) for the parity tests"""
    cases = []
    cases.append(("ba100_m50", graphs.barabasi_albert(100, 50, seed=1), 100, None, 50))
    cases.append(("ba300_m3", graphs.barabasi_albert(300, 3, seed=2), 300, None, 150))
    cases.append(("sbm_cora", graphs.sbm(2708, 7, 5278, seed=0), 2708, None, 812))
    ei, ptr = graphs.proteins_like_batch(40, seed=3)
    cases.append(("proteins40_union", ei, int(ptr[-1]), None, int(ptr[-1]) // 2))
    cases.append(("proteins40_pergraph", ei, int(ptr[-1]), ptr, (np.diff(ptr) // 2)))
    return cases
