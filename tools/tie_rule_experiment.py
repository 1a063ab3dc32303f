"""CPU experiment behind DESIGN.md §3.3: the chained-elimination statistic of scripts/rlap_vc_spectral.py (BA-1000, 10 x 5 %,
o_v = random) on the sequential oracle, keyed mode (what the CUDA path computes bit for bit) against ref mode (the reference).

    python tools/tie_rule_experiment.py RUNS {asc|desc|random} {keyed|ref} [FIRST_RUN]

prints the per-step means of the top singular value (with standard errors) and of the edge count; the per-run arrays go
to /tmp/tie2_<which>_<o_n>_<first>.npy. Runs r of both sides use the seeds of tests/test_gpu_adapters.py."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import port
from rlap_b200 import graphs
from tests import util

def relabel(rows, cols, w):
    nodes = np.unique(np.concatenate([rows, cols]))
    return nodes, np.searchsorted(nodes, rows), np.searchsorted(nodes, cols), w

def chain(fn, ei, n, steps, t):
    rows, cols = ei[0].astype(np.int64), ei[1].astype(np.int64)
    w = None
    sv, nn, ne = [], [], []
    for k in range(steps):
        r, c, ww = fn(k, rows, cols, w, n, t)
        nodes, r2, c2, ww = relabel(r, c, ww)
        n = nodes.shape[0]
        rows, cols, w = r2, c2, ww
        A = np.zeros((n, n), dtype=np.float32); A[rows, cols] = 1
        sv.append(np.linalg.norm(A, 2)); nn.append(n); ne.append(rows.shape[0])
    return np.array([sv, nn, ne], dtype=np.float64)

def keyed_fn(seed, o_n):
    def fn(k, rows, cols, w, n, t):
        ptr, col, ww = port.ingest(np.stack([rows, cols]), w, n)
        r, c, x = port.keyed_schur(ptr, col, ww, t, "random", o_n, seed=seed + k, view=0)
        return r.astype(np.int64), c.astype(np.int64), x
    return fn

def ref_fn(rr, o_n):
    def fn(k, rows, cols, w, n, t):
        out = port.ref_approximate_cholesky(util.edge_info(np.stack([rows, cols]), None if w is None else w.astype(np.float64)), n, t, "random", o_n,
                                            sample_seed=7 + 31 * rr + k, rd_seed=1000 * rr + k)
        return out[:, 0].astype(np.int64), out[:, 1].astype(np.int64), out[:, 2]
    return fn

if __name__ == "__main__":
    R = int(sys.argv[1]); o_n = sys.argv[2]; which = sys.argv[3]; r0 = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    n = 1000
    ei = graphs.barabasi_albert(n, 5, seed=3)
    acc = []
    for r in range(r0, r0 + R):
        f = keyed_fn(100 * r, o_n) if which == "keyed" else ref_fn(r, o_n)
        acc.append(chain(f, ei, n, 10, 50))
    acc = np.array(acc)
    m = acc.mean(0); s = acc.std(0) / np.sqrt(R)
    np.save(f"/tmp/tie2_{which}_{o_n}_{r0}.npy", acc)
    print(which, o_n, "sv", np.round(m[0], 3)); print("  se", np.round(s[0], 3)); print("  edges", m[2])
