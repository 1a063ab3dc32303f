"""Statistical parity of the CUDA path with the reference's sampling (north_star, third criterion):
the mean Laplacian over K views converges to the exact Schur complement like the reference's does
(reference = oracle ref mode, which is bit-identical to the unmodified C++), and edge-count / weight /
top-singular-value statistics agree within a stated tolerance."""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu

KS = (4, 16, 64, 256)


def _curve_from_views(row, col, w, vp, n, L0, Ks):
    groups, out = {}, []
    for s in range(max(Ks)):
        r, c, wt = row[vp[s]:vp[s + 1]], col[vp[s]:vp[s + 1]], w[vp[s]:vp[s + 1]]
        keep = tuple(np.unique(c).tolist())
        g = groups.setdefault(keep, [np.zeros((n, n)), 0, None])
        g[0] += util.laplacian(r, c, wt, n)
        g[1] += 1
        if g[2] is None:
            ex = np.zeros((n, n))
            k = np.array(keep, dtype=np.int64)
            ex[np.ix_(k, k)] = util.exact_schur(L0, k)
            g[2] = ex
        if s + 1 in Ks:
            num = sum(np.linalg.norm(v[0] / v[1] - v[2]) ** 2 * v[1] for v in groups.values())
            den = sum(np.linalg.norm(v[2]) ** 2 * v[1] for v in groups.values())
            out.append(np.sqrt(num / den))
    return np.array(out)


def _modal_views(row, col, w, vp):
    """the views whose surviving set is the most frequent one. o_v = degree on BA-100 removes the same 50 vertices in
    all but the odd view (0 - 2 of 256, on either side); such a singleton group enters the pooled error with its
    full one-sample error and moves the K = 256 point by 40 %, whichever side it falls on. The curves are compared on
    the common set instead."""
    keeps = [tuple(np.unique(col[vp[s]:vp[s + 1]]).tolist()) for s in range(len(vp) - 1)]
    modal = max(set(keeps), key=keeps.count)
    sel = [s for s, k in enumerate(keeps) if k == modal]
    nvp = np.concatenate([[0], np.cumsum([vp[s + 1] - vp[s] for s in sel])])
    cat = lambda a: np.concatenate([a[vp[s]:vp[s + 1]] for s in sel])
    return cat(row), cat(col), cat(w), nvp


def _ref_views(oracle_port, info, n, t, o_v, o_n, K):
    rows, cols, ws, vp = [], [], [], [0]
    for s in range(K):
        o = oracle_port.ref_approximate_cholesky(info, n, t, o_v, o_n,
                                                 sample_seed=((s + 1) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF, rd_seed=4)
        rows.append(o[:, 0].astype(np.int64)); cols.append(o[:, 1].astype(np.int64)); ws.append(o[:, 2])
        vp.append(vp[-1] + o.shape[0])
    return np.concatenate(rows), np.concatenate(cols), np.concatenate(ws), np.array(vp)


@pytest.mark.parametrize("o_v,o_n,t", [("random", "asc", 1), ("random", "asc", 10), ("random", "asc", 50),
                                       ("random", "desc", 50), ("random", "random", 50), ("degree", "asc", 50),
                                       ("degree", "random", 50), ("coarsen", "asc", 10)])
def test_error_curve_matches_reference(oracle_port, o_v, o_n, t):
    """relative Frobenius error of the K-view mean vs the exact Schur complement: GPU and reference curves agree
    point by point (+-35 %) and decay like 1/sqrt(K) over K = 4..256"""
    import rlap_b200
    from rlap_b200 import graphs
    n = 100
    ei = graphs.barabasi_albert(n, 50, seed=1)
    L0 = util.laplacian(ei[0], ei[1], np.ones(ei.shape[1]), n)
    K = max(KS)
    ei_gpu, back = ei, np.arange(n)
    if o_v == "random":
        # make both sides eliminate the same vertices in the same order: relabel the graph so that the keyed
        # permutation (shared by all views) visits the images of the reference's pop order
        sigma_ref = oracle_port.ref_random_order(n, 4)                         # vertex eliminated at step i
        pi_gpu = np.argsort(oracle_port.rank_perm(2024, 0, 0, n))              # vertex of rank i
        f = np.empty(n, dtype=np.int64)
        f[sigma_ref] = pi_gpu
        ei_gpu = f[ei]
        back = np.empty(n, dtype=np.int64)
        back[f] = np.arange(n)
    g = rlap_b200.prepare(torch.from_numpy(np.ascontiguousarray(ei_gpu)).cuda(), None, n)
    (row, col, w), vp = rlap_b200.schur_views(g, t, o_v, o_n, num_views=K, seed=2024, shared_order=True, dtype=None)
    gv = (back[row.cpu().numpy()], back[col.cpu().numpy()], w.cpu().numpy().astype(np.float64), vp.numpy())
    rv = _ref_views(oracle_port, util.edge_info(ei), n, t, o_v, o_n, K)
    ks = KS
    if o_v == "degree":
        gv, rv = _modal_views(*gv), _modal_views(*rv)
        kmin = min(len(gv[3]), len(rv[3])) - 1
        assert kmin >= K - 8, (len(gv[3]) - 1, len(rv[3]) - 1)      # the odd view out is rare on both sides
        ks = KS[:-1] + (kmin,)
    cg = _curve_from_views(*gv, n, L0, ks)
    cr = _curve_from_views(*rv, n, L0, ks)
    # coarsen: one pick per vertex instead of one per neighbour and a seed-dependent elimination set -> noisier
    lo, hi = (0.5, 1.6) if o_v == "coarsen" else (0.65, 1.35)
    assert np.all(cg / cr < hi) and np.all(cg / cr > lo), (cg, cr)
    if o_v == "random":   # identical elimination sets on both sides: the K = 256 points agree within 25 %
        assert abs(cg[-1] / cr[-1] - 1.0) < 0.25, (cg, cr)
    assert cg[-1] < cg[0] / (2.0 if o_v == "coarsen" else 4.0), cg


@pytest.mark.parametrize("o_v,o_n", [("random", "asc"), ("random", "desc"), ("degree", "asc"), ("coarsen", "asc")])
def test_edge_count_weight_and_spectrum_statistics(oracle_port, o_v, o_n):
    """per-view statistics over 64 seeds on the Cora-shaped graph: number of rows, total weight, largest weight and
    the top singular value of the unweighted view (the statistic of scripts/rlap_vc_spectral.py:55-57) have the
    same mean as the reference's within 3 sigma of the seed-to-seed spread (and within 2 % for the counts)"""
    import rlap_b200
    from rlap_b200 import graphs
    n, t, K = 2708, 812, 64
    ei = graphs.sbm(n, 7, 5278, seed=0)
    g = rlap_b200.prepare(torch.from_numpy(ei).cuda(), None, n)
    (row, col, w), vp = rlap_b200.schur_views(g, t, o_v, o_n, num_views=K, seed=77, dtype=None)
    row, col, w, vp = row.cpu().numpy(), col.cpu().numpy(), w.cpu().numpy().astype(np.float64), vp.numpy()
    rr, rc, rw, rvp = _ref_views(oracle_port, util.edge_info(ei), n, t, o_v, o_n, K)
    # the reference draws its permutation from one injected stream (rd_seed) - vary it like the GPU varies views
    if o_v == "random":
        parts = [oracle_port.ref_approximate_cholesky(util.edge_info(ei), n, t, o_v, o_n, sample_seed=1 + s, rd_seed=100 + s)
                 for s in range(K)]
        rr = np.concatenate([p[:, 0] for p in parts]).astype(np.int64)
        rc = np.concatenate([p[:, 1] for p in parts]).astype(np.int64)
        rw = np.concatenate([p[:, 2] for p in parts])
        rvp = np.concatenate([[0], np.cumsum([p.shape[0] for p in parts])])

    def stats(r, c, wt, p):
        out = []
        for s in range(K):
            a, b, x = r[p[s]:p[s + 1]], c[p[s]:p[s + 1]], wt[p[s]:p[s + 1]]
            keep = np.unique(b)
            A = np.zeros((n, n), dtype=np.float32)
            A[a, b] = 1.0
            sv = float(torch.linalg.matrix_norm(torch.from_numpy(A[np.ix_(keep, keep)]).cuda(), ord=2))
            out.append((a.shape[0], x.sum(), x.max(), sv))
        return np.array(out)

    sg, sr = stats(row, col, w, vp), stats(rr, rc, rw, rvp)
    mg, mr = sg.mean(0), sr.mean(0)
    sd = np.maximum(sr.std(0), sg.std(0)) / np.sqrt(K) + 1e-12
    assert abs(mg[0] - mr[0]) / mr[0] < 0.02 and abs(mg[1] - mr[1]) / mr[1] < 0.02, (mg, mr)
    assert np.all(np.abs(mg - mr) < 3 * np.sqrt(2) * sd + 0.02 * np.abs(mr)), (mg, mr, sd)


KS_TIGHT = (16, 64, 256, 1024, 4096)


@pytest.mark.parametrize("o_n", ["asc", "desc", "random"])
def test_error_curve_and_bias_floor_match_reference_within_15_percent(oracle_port, o_n):
    """SURVEY.md App. B.2 at its own resolution: BA-100, t = 50, o_v = random (the configuration with a visible bias
    floor), K up to 4096, two independent K-batches per side averaged to halve the noise. The GPU curve follows the
    reference's point by point within +-15 %, and both leave the 1/sqrt(K) line at large K by the same factor (the
    floor: relFrob * sqrt(K) at K = 4096 is well above its value at K = 16)."""
    import rlap_b200
    from rlap_b200 import graphs
    n, t, R = 100, 50, 2
    ei = graphs.barabasi_albert(n, 50, seed=1)
    L0 = util.laplacian(ei[0], ei[1], np.ones(ei.shape[1]), n)
    K = max(KS_TIGHT)
    sigma_ref = oracle_port.ref_random_order(n, 4)
    info = util.edge_info(ei)
    cg, cr = [], []
    for rep in range(R):
        seed = 3000 + rep
        pi_gpu = np.argsort(oracle_port.rank_perm(seed, 0, 0, n))
        f = np.empty(n, dtype=np.int64)
        f[sigma_ref] = pi_gpu
        back = np.empty(n, dtype=np.int64)
        back[f] = np.arange(n)
        g = rlap_b200.prepare(torch.from_numpy(np.ascontiguousarray(f[ei])).cuda(), None, n)
        (row, col, w), vp = rlap_b200.schur_views(g, t, "random", o_n, num_views=K, seed=seed, shared_order=True, dtype=None)
        cg.append(_curve_from_views(back[row.cpu().numpy()], back[col.cpu().numpy()], w.cpu().numpy().astype(np.float64),
                                    vp.numpy(), n, L0, KS_TIGHT))
        rows, cols, ws, rvp = [], [], [], [0]
        for s in range(K):
            o = oracle_port.ref_approximate_cholesky(info, n, t, "random", o_n,
                                                     sample_seed=((rep * K + s + 1) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF,
                                                     rd_seed=4)
            rows.append(o[:, 0].astype(np.int64)); cols.append(o[:, 1].astype(np.int64)); ws.append(o[:, 2])
            rvp.append(rvp[-1] + o.shape[0])
        cr.append(_curve_from_views(np.concatenate(rows), np.concatenate(cols), np.concatenate(ws), np.array(rvp), n, L0,
                                    KS_TIGHT))
    cg, cr = np.mean(cg, 0), np.mean(cr, 0)
    ratio = cg / cr
    print(f"o_n={o_n}: K={KS_TIGHT} GPU {np.round(cg, 5)} reference {np.round(cr, 5)} ratio {np.round(ratio, 3)}")
    assert np.all(np.abs(ratio - 1.0) < 0.15), (cg, cr)
    sq = np.sqrt(np.array(KS_TIGHT, dtype=np.float64))
    floor_g, floor_r = cg[-1] * sq[-1] / (cg[0] * sq[0]), cr[-1] * sq[-1] / (cr[0] * sq[0])
    assert floor_r > 1.3 and floor_g > 1.3, (floor_g, floor_r)      # both curves flatten: the bias floor of App. B.2
    assert abs(floor_g / floor_r - 1.0) < 0.2, (floor_g, floor_r)
