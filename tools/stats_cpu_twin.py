"""CPU twin of tests/test_gpu_stats.py: the keyed oracle stands in for the CUDA path (bit-identical by the parity tests),
to see where the statistical tests land under a change of the keyed specification before spending GPU time."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import port
from rlap_b200 import graphs
from tests import util
from tests.test_gpu_stats import _curve_from_views, _ref_views, KS, KS_TIGHT

def keyed_views(ei, n, t, o_v, o_n, K, seed, shared):
    ptr, col, w = port.ingest(ei, None, n)
    rows, cols, ws, vp = [], [], [], [0]
    for s in range(K):
        r, c, x = port.keyed_schur(ptr, col, w, t, o_v, o_n, seed=seed, view=s, flags=2 if shared else 0)
        rows.append(r.astype(np.int64)); cols.append(c.astype(np.int64)); ws.append(x.astype(np.float64)); vp.append(vp[-1] + r.shape[0])
    return np.concatenate(rows), np.concatenate(cols), np.concatenate(ws), np.array(vp)

def curve_test(o_v, o_n, t):
    n = 100
    ei = graphs.barabasi_albert(n, 50, seed=1)
    L0 = util.laplacian(ei[0], ei[1], np.ones(ei.shape[1]), n)
    K = max(KS)
    ei_gpu, back = ei, np.arange(n)
    if o_v == "random":
        sigma_ref = port.ref_random_order(n, 4)
        pi_gpu = np.argsort(port.rank_perm(2024, 0, 0, n))
        f = np.empty(n, dtype=np.int64); f[sigma_ref] = pi_gpu
        ei_gpu = f[ei]
        back = np.empty(n, dtype=np.int64); back[f] = np.arange(n)
    row, col, w, vp = keyed_views(np.ascontiguousarray(ei_gpu), n, t, o_v, o_n, K, 2024, True)
    cg = _curve_from_views(back[row], back[col], w, vp, n, L0, KS)
    cr = _curve_from_views(*_ref_views(port, util.edge_info(ei), n, t, o_v, o_n, K), n, L0, KS)
    print("curve", o_v, o_n, t, "ratio", np.round(cg / cr, 3), flush=True)

def tight_test(o_n):
    n, t, R = 100, 50, 2
    ei = graphs.barabasi_albert(n, 50, seed=1)
    L0 = util.laplacian(ei[0], ei[1], np.ones(ei.shape[1]), n)
    K = max(KS_TIGHT)
    sigma_ref = port.ref_random_order(n, 4)
    info = util.edge_info(ei)
    cg, cr = [], []
    for rep in range(R):
        seed = 3000 + rep
        pi_gpu = np.argsort(port.rank_perm(seed, 0, 0, n))
        f = np.empty(n, dtype=np.int64); f[sigma_ref] = pi_gpu
        back = np.empty(n, dtype=np.int64); back[f] = np.arange(n)
        row, col, w, vp = keyed_views(np.ascontiguousarray(f[ei]), n, t, "random", o_n, K, seed, True)
        cg.append(_curve_from_views(back[row], back[col], w, vp, n, L0, KS_TIGHT))
        rows, cols, ws, rvp = [], [], [], [0]
        for s in range(K):
            o = port.ref_approximate_cholesky(info, n, t, "random", o_n,
                                              sample_seed=((rep * K + s + 1) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF, rd_seed=4)
            rows.append(o[:, 0].astype(np.int64)); cols.append(o[:, 1].astype(np.int64)); ws.append(o[:, 2]); rvp.append(rvp[-1] + o.shape[0])
        cr.append(_curve_from_views(np.concatenate(rows), np.concatenate(cols), np.concatenate(ws), np.array(rvp), n, L0, KS_TIGHT))
    cg, cr = np.mean(cg, 0), np.mean(cr, 0)
    sq = np.sqrt(np.array(KS_TIGHT, dtype=np.float64))
    fg, fr = cg[-1] * sq[-1] / (cg[0] * sq[0]), cr[-1] * sq[-1] / (cr[0] * sq[0])
    print("tight", o_n, "ratio", np.round(cg / cr, 3), "floors", round(fg, 3), round(fr, 3), flush=True)

def spectrum_test(o_v, o_n):
    n, t, K = 2708, 812, 64
    ei = graphs.sbm(n, 7, 5278, seed=0)
    row, col, w, vp = keyed_views(ei, n, t, o_v, o_n, K, 77, False)
    rr, rc, rw, rvp = _ref_views(port, util.edge_info(ei), n, t, o_v, o_n, K)
    if o_v == "random":
        parts = [port.ref_approximate_cholesky(util.edge_info(ei), n, t, o_v, o_n, sample_seed=1 + s, rd_seed=100 + s) for s in range(K)]
        rr = np.concatenate([p[:, 0] for p in parts]).astype(np.int64); rc = np.concatenate([p[:, 1] for p in parts]).astype(np.int64)
        rw = np.concatenate([p[:, 2] for p in parts]); rvp = np.concatenate([[0], np.cumsum([p.shape[0] for p in parts])])
    def stats(r, c, wt, p):
        out = []
        for s in range(K):
            a, b, x = r[p[s]:p[s + 1]], c[p[s]:p[s + 1]], wt[p[s]:p[s + 1]]
            keep = np.unique(b)
            A = np.zeros((n, n), dtype=np.float32); A[a, b] = 1.0
            import scipy.sparse.linalg as sla, scipy.sparse as sp
            sv = float(sla.svds(sp.csr_matrix(A[np.ix_(keep, keep)]), k=1, return_singular_vectors=False)[0])
            out.append((a.shape[0], x.sum(), x.max(), sv))
        return np.array(out)
    sg, sr = stats(row, col, w, vp), stats(rr, rc, rw, rvp)
    mg, mr = sg.mean(0), sr.mean(0)
    sd = np.maximum(sr.std(0), sg.std(0)) / np.sqrt(K) + 1e-12
    ok = np.all(np.abs(mg - mr) < 3 * np.sqrt(2) * sd + 0.02 * np.abs(mr))
    print("spectrum", o_v, o_n, "rel gap", np.round((mg - mr) / mr, 4), "in sigma", np.round((mg - mr) / (np.sqrt(2) * sd), 2), "pass", ok, flush=True)

if __name__ == "__main__":
    what = sys.argv[1]
    if what == "curve":
        for o_v, o_n, t in ([("degree", "asc", 50)] if len(sys.argv) > 2 else [("random", "asc", 1), ("random", "asc", 10), ("random", "asc", 50), ("random", "desc", 50), ("degree", "asc", 50)]):
            curve_test(o_v, o_n, t)
    elif what == "tight":
        for o_n in ("asc", "desc"):
            tight_test(o_n)
    else:
        for o_v, o_n in [("random", "asc"), ("random", "desc"), ("degree", "asc")]:
            spectrum_test(o_v, o_n)
