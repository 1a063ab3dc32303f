"""The chained-elimination statistic of scripts/rlap_vc_spectral.py:14-58 on the sequential oracle: keyed mode (what the
CUDA path computes bit for bit) against ref mode (the reference). This is the statistic that exposes the tie order of
o_n = asc / desc: ties by neighbour id put it 7 - 11 % low, round 1's Philox key 3 % low, the exact
std::sort order within the noise. CPU twin of tests/test_gpu_adapters.py::test_chained_elimination_statistics_match_reference
(same graph, same seeds, 16 runs)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as sla

from rlap_b200 import graphs
from tests import util


def _top_sv(rows, cols, n):
    A = sp.csr_matrix((np.ones(rows.shape[0]), (rows, cols)), shape=(n, n))
    A.data[:] = 1.0                                   # unweighted, duplicates collapse
    return float(sla.eigsh(A.asfptype(), k=1, which="LA", return_eigenvectors=False, tol=1e-9)[0])   # symmetric, non-negative


def _chain(fn, ei, n, steps, t):
    rows, cols, w = ei[0].astype(np.int64), ei[1].astype(np.int64), None
    out = []
    for k in range(steps):
        r, c, ww = fn(k, rows, cols, w, n, t)
        nodes = np.unique(np.concatenate([r, c]))
        rows, cols, w, n = np.searchsorted(nodes, r), np.searchsorted(nodes, c), ww, nodes.shape[0]
        out.append((_top_sv(rows, cols, n), n, rows.shape[0]))
    return np.array(out, dtype=np.float64).T


@pytest.mark.parametrize("o_n", ["asc", "desc"])
def test_keyed_chained_statistic_matches_reference(oracle_port, o_n):
    n, steps, t, R = 1000, 10, 50, 16
    ei = graphs.barabasi_albert(n, 5, seed=3)

    def keyed(r):
        def fn(k, rows, cols, w, n_, t_):
            ptr, col, ww = oracle_port.ingest(np.stack([rows, cols]), w, n_)
            a, b, x = oracle_port.keyed_schur(ptr, col, ww, t_, "random", o_n, seed=100 * r + k, view=0)
            return a.astype(np.int64), b.astype(np.int64), x
        return fn

    def ref(r):
        def fn(k, rows, cols, w, n_, t_):
            info = util.edge_info(np.stack([rows, cols]), None if w is None else w.astype(np.float64))
            o = oracle_port.ref_approximate_cholesky(info, n_, t_, "random", o_n, sample_seed=7 + 31 * r + k, rd_seed=1000 * r + k)
            return o[:, 0].astype(np.int64), o[:, 1].astype(np.int64), o[:, 2]
        return fn

    got = np.mean([_chain(keyed(r), ei, n, steps, t) for r in range(R)], axis=0)
    want = np.mean([_chain(ref(r), ei, n, steps, t) for r in range(R)], axis=0)
    assert np.all(got[1] == want[1]), (got[1], want[1])                        # node counts
    assert np.all(np.abs(got[2] - want[2]) <= 0.01 * want[2]), (got[2], want[2])   # edge counts
    assert np.all(np.abs(got[0] - want[0]) <= 0.03 * want[0]), (got[0] / want[0] - 1)   # top singular value
