"""The keyed-mode specification (oracle/rlap_oracle.cc, what the CUDA path reproduces bit for bit):
known-answer tests and invariants and its statistical agreement with the
reference's sampling rule (ref mode)."""
import numpy as np
import pytest

from rlap_b200 import graphs
from tests import util


def test_philox_known_answers(oracle_port):
    # Random123 kat_vectors, philox4x32-10
    kat = [((0, 0), (0, 0, 0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff, 0xffffffff), (0xffffffff,) * 4, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0xa4093822, 0x299f31d0), (0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for key, ctr, want in kat:
        assert tuple(int(x) for x in oracle_port.philox(key[0], key[1], *ctr)) == want


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 16, 17, 39, 100, 1000, 4097])
def test_rank_permutation_is_a_bijection(oracle_port, n):
    for seed in (0, 1, 12345):
        r = oracle_port.rank_perm(seed, 3, 7, n)
        assert sorted(r.tolist()) == list(range(n))


def test_rank_permutation_is_uniform(oracle_port):
    n, K = 6, 30000
    cnt = np.zeros((n, n))
    for s in range(K):
        cnt[np.arange(n), oracle_port.rank_perm(s, 0, 0, n)] += 1
    sigma = np.sqrt((1 / n) * (1 - 1 / n) / K)
    assert np.abs(cnt / K - 1 / n).max() < 5 * sigma


def test_ingest_semantics(oracle_port):
    # zero weights dropped, duplicates summed in order, rows ascending (reader.cc:42-61)
    ei = np.array([[1, 0, 2, 0, 1, 0, 2, 1], [0, 1, 0, 2, 0, 1, 1, 2]])
    w = np.array([1.0, 1.0, 2.0, 2.0, 0.5, 0.5, 0.0, 0.0], dtype=np.float32)
    ptr, col, ww = oracle_port.ingest(ei, w, 3)
    assert ptr.tolist() == [0, 2, 3, 4]
    assert col.tolist() == [1, 2, 0, 0]
    assert ww.tolist() == [1.5, 2.0, 1.5, 2.0]
    with pytest.raises(ValueError):
        oracle_port.ingest(np.array([[0], [3]]), None, 3)
    with pytest.raises(ValueError):
        oracle_port.ingest(np.array([[1], [1]]), None, 3)


@pytest.mark.parametrize("o_v,o_n", util.COMBOS)
def test_keyed_invariants(oracle_port, o_v, o_n):
    n, t = 300, 120
    ei = graphs.barabasi_albert(n, 4, seed=11)
    w = util.sym_weights(ei)
    ptr, col, ww = oracle_port.ingest(ei, w, n)
    r, c, wt, st, order = oracle_port.keyed_schur(ptr, col, ww, t, o_v, o_n, seed=3, return_stats=True,
                                                  return_order=True)
    elim = np.where(order >= 0)[0]
    assert elim.shape[0] == t                                    # (iii) exactly min(t, n-1) eliminated
    assert not np.isin(r, elim).any() and not np.isin(c, elim).any()   # no row touches an eliminated vertex
    A = np.zeros((n, n))
    A[r, c] = wt
    assert np.array_equal(A, A.T)                                # (iv) symmetric, weights included
    assert (wt > 0).all()
    assert np.all(np.diff(c) >= 0)                               # sorted by (col, row)
    same = np.diff(c) == 0
    assert np.all(np.diff(r)[same] > 0)


def test_keyed_edge_cases(oracle_port):
    n = 64
    ei = graphs.barabasi_albert(60, 3, seed=4)                   # vertices 60..63 isolated
    ptr, col, w = oracle_port.ingest(ei, None, n)
    r, c, wt = oracle_port.keyed_schur(ptr, col, w, 0, "degree", "asc")
    assert r.shape[0] == ei.shape[1] and np.array_equal(wt, np.ones_like(wt))   # (v) t=0: coalesced input
    for o_v in ("random", "degree", "coarsen"):
        r, c, wt = oracle_port.keyed_schur(ptr, col, w, n - 1, o_v, "asc")
        assert r.shape[0] == 0                                   # (vi) t >= n-1: empty
        r, c, wt = oracle_port.keyed_schur(ptr, col, w, 10 ** 6, o_v, "asc")
        assert r.shape[0] == 0
    # isolated vertices consume removal slots under degree order (key 0 pops first, A.3)
    r, c, wt, order = oracle_port.keyed_schur(ptr, col, w, 4, "degree", "asc", return_order=True)
    assert set(np.where(order >= 0)[0].tolist()) == {60, 61, 62, 63}
    assert r.shape[0] == ei.shape[1]


def test_mass_invariant_single_elimination(oracle_port):
    """(i) eliminating one vertex of a star-free graph: sum of fill weights = (S^2 - sum w^2) / (2S)"""
    n = 40
    ei = graphs.barabasi_albert(n, 6, seed=2)
    w = util.sym_weights(ei)
    ptr, col, ww = oracle_port.ingest(ei, w, n)
    for seed in range(5):
        r, c, wt, order = oracle_port.keyed_schur(ptr, col, ww, 1, "random", "desc", seed=seed, return_order=True)
        i = int(np.where(order >= 0)[0][0])
        wi = ww[ptr[i]:ptr[i + 1]].astype(np.float64)
        S = wi.sum()
        L0 = util.laplacian(col, np.repeat(np.arange(n), np.diff(ptr)), ww, n)
        L1 = util.laplacian(r, c, wt, n)
        keep = np.setdiff1d(np.arange(n), [i])
        # total edge weight among the survivors grows by exactly the clique mass
        m0 = -np.triu(L0[np.ix_(keep, keep)], 1).sum()
        m1 = -np.triu(L1[np.ix_(keep, keep)], 1).sum()
        assert abs((m1 - m0) - (S * S - (wi * wi).sum()) / (2 * S)) < 1e-5 * S


@pytest.mark.parametrize("o_v", ["random", "degree", "coarsen"])
def test_full_clique_is_exact_schur_complement(oracle_port, o_v):
    n = 100
    ei = graphs.barabasi_albert(n, 50, seed=1)
    w = util.sym_weights(ei)
    ptr, col, ww = oracle_port.ingest(ei, w, n)
    r, c, wt = oracle_port.keyed_schur(ptr, col, ww, 50, o_v, "asc", seed=7, flags=oracle_port.FLAG_FULL_CLIQUE)
    keep = np.unique(c)
    L0 = util.laplacian(ei[0], ei[1], w, n)
    Ls = util.laplacian(r, c, wt, n)[np.ix_(keep, keep)]
    ex = util.exact_schur(L0, keep)
    assert np.linalg.norm(Ls - ex) / np.linalg.norm(ex) < 1e-5   # fp32 weights


def _mean_error_curve(sampler, L0, n, Ks):
    """relative Frobenius error of the mean Laplacian over K seeds vs the exact Schur complement of the
    same elimination set (sets are grouped: orders that depend on the seed are averaged per set)"""
    groups = {}
    out = []
    for s in range(max(Ks)):
        r, c, wt = sampler(s)
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


@pytest.mark.parametrize("o_v,o_n,t", [("random", "asc", 1), ("random", "asc", 10), ("random", "desc", 50),
                                       ("random", "random", 50), ("degree", "asc", 50), ("coarsen", "asc", 10)])
def test_keyed_error_curve_matches_reference(oracle_port, o_v, o_n, t):
    """the K-seed mean converges to the exact Schur complement like the reference's does
    (SURVEY.md App. B.2): same curve point by point (+-25 %), ~1/sqrt(K) decay over the tested range"""
    n = 100
    ei = graphs.barabasi_albert(n, 50, seed=1)
    info = util.edge_info(ei)
    ptr, col, w = oracle_port.ingest(ei, None, n)
    L0 = util.laplacian(ei[0], ei[1], np.ones(ei.shape[1]), n)
    Ks = (4, 16, 64, 256)
    shared = oracle_port.FLAG_SHARED_ORDER

    def keyed(s):
        return oracle_port.keyed_schur(ptr, col, w, t, o_v, o_n, seed=99, view=s, flags=shared)

    def ref(s):
        # well-spread seeds: consecutive mt19937_64 seeds give correlated first draws
        o = oracle_port.ref_approximate_cholesky(info, n, t, o_v, o_n,
                                                 sample_seed=((s + 1) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF, rd_seed=4)
        return o[:, 0].astype(np.int64), o[:, 1].astype(np.int64), o[:, 2]

    ck, cr = _mean_error_curve(keyed, L0, n, Ks), _mean_error_curve(ref, L0, n, Ks)
    assert np.all(ck / cr < 1.35) and np.all(ck / cr > 0.65), (ck, cr)
    # ~ 1/sqrt(64) = 1/8 expected between K=4 and K=256; under coarsen the elimination SET depends on the sampled
    # contractions (dynamic degrees), so seeds split into groups and each group averages fewer samples
    assert ck[-1] < ck[0] / (2.0 if o_v == "coarsen" else 4.0)


def test_degree_rounds_track_the_reference_order(oracle_port):
    """the round-parallel restatement of the dynamic min-degree order eliminates a set
    with the same degree profile as the reference's sequential bucket queue and the same work counters"""
    n = 4000
    ei = graphs.barabasi_albert(n, 7, seed=0)
    info = util.edge_info(ei)
    ptr, col, w = oracle_port.ingest(ei, None, n)
    t = n // 2
    d0 = np.diff(ptr)
    for o_v in ("degree", "coarsen"):
        r, c, wt, st, order = oracle_port.keyed_schur(ptr, col, w, t, o_v, "asc", seed=1, return_stats=True,
                                                      return_order=True)
        out, cnt = oracle_port.ref_approximate_cholesky(info, n, t, o_v, "asc", return_counters=True)
        elim_k = np.where(order >= 0)[0]
        elim_r = np.setdiff1d(np.arange(n), np.unique(out[:, 1].astype(np.int64)))
        # survivors without edges do not appear in the output: compare on degree histograms instead
        hk, hr = np.bincount(d0[elim_k], minlength=64)[:64], np.bincount(d0[elim_r], minlength=64)[:64]
        assert np.abs(hk - hr).sum() <= 0.08 * t
        assert abs(st["D"] - cnt[0]) <= 0.02 * cnt[0] and abs(st["F"] - cnt[1]) <= 0.02 * cnt[1]
        assert abs(r.shape[0] - out.shape[0]) <= 0.01 * out.shape[0]
        jacc = np.intersect1d(elim_k, elim_r).shape[0] / np.union1d(elim_k, elim_r).shape[0]
        assert jacc > 0.7
