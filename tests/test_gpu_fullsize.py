"""BASELINE.json's full-size configurations on the GPU: bit-exact against the keyed oracle where the oracle
finishes in seconds, plus the size-independent properties of SURVEY.md App. A.7."""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def _check_properties(row, col, w, n, n_elim_expected, ei_keep_nodes=None):
    assert (w > 0).all()
    key = row.astype(np.int64) * n + col
    keyt = col.astype(np.int64) * n + row
    o1, o2 = np.argsort(key), np.argsort(keyt)
    assert np.array_equal(key[o1], keyt[o2])                       # structurally symmetric
    assert np.array_equal(w[o1], w[o2])                            # and the two directions carry the same weight
    assert np.all(np.diff(key[np.lexsort((row, col))]) != 0)       # no duplicate rows
    order = np.lexsort((row, col))
    assert np.array_equal(order, np.arange(row.shape[0]))          # sorted by (col, row)


@pytest.mark.parametrize("o_v,o_n", [("degree", "asc"), ("random", "asc"), ("coarsen", "asc"), ("degree", "desc")])
def test_c4_arxiv_shape_matches_oracle(oracle_port, o_v, o_n):
    """C4: BA n=169,343, m=7 (E ~ 2.37M directed), num_remove = 50 %"""
    import rlap_b200
    from rlap_b200 import graphs
    n = 169343
    ei = graphs.barabasi_albert(n, 7, seed=0)
    optr, ocol, ow = oracle_port.ingest(ei, None, n)
    g = rlap_b200.prepare(torch.from_numpy(ei).cuda(), None, n)
    assert g.nnz == ocol.shape[0]
    (row, col, w), vp, st = rlap_b200.schur_views(g, n // 2, o_v, o_n, num_views=3, seed=11, view_base=2, dtype=None,
                                                  return_stats=True)
    row, col, w, vp = row.cpu().numpy(), col.cpu().numpy(), w.cpu().numpy(), vp.numpy()
    r0, c0, w0, so, order = oracle_port.keyed_schur(optr, ocol, ow, n // 2, o_v, o_n, seed=11, view=3, return_stats=True,
                                                    return_order=True)
    s, e = vp[1], vp[2]
    assert np.array_equal(row[s:e], r0) and np.array_equal(col[s:e], c0)
    assert np.array_equal(w[s:e].view(np.uint32), w0.view(np.uint32))
    elim = order >= 0
    assert elim.sum() == n // 2
    for v in range(3):
        rr, cc, ww = row[vp[v]:vp[v + 1]], col[vp[v]:vp[v + 1]], w[vp[v]:vp[v + 1]]
        _check_properties(rr, cc, ww, n, n // 2)
    assert not elim[row[s:e]].any() and not elim[col[s:e]].any()    # no row touches an eliminated vertex
    assert st["fills"] > 0 and st["rows"] == vp[-1]


def test_c3_proteins_batch_matches_oracle(oracle_port):
    """C3: 1,113 small graphs (~39 nodes), two views per graph, per-graph num_remove = 50 %, and the reference's
    union-batch semantics (128 graphs treated as one graph, scripts/graph_shared.py:139-146)"""
    import rlap_b200
    from rlap_b200 import graphs
    ei, ptr = graphs.proteins_like_batch(1113, seed=0)
    n = int(ptr[-1])
    t = np.diff(ptr) // 2
    optr, ocol, ow = oracle_port.ingest(ei, None, n)
    g = rlap_b200.prepare(torch.from_numpy(ei).cuda(), None, n, graph_ptr=ptr)
    for o_v, o_n in (("random", "asc"), ("degree", "asc"), ("coarsen", "random"), ("random", "random")):
        (row, col, w), vp = rlap_b200.schur_views(g, t, o_v, o_n, num_views=2, seed=5, dtype=None)
        row, col, w, vp = row.cpu().numpy(), col.cpu().numpy(), w.cpu().numpy(), vp.numpy()
        for v in range(2):
            r0, c0, w0 = oracle_port.keyed_schur(optr, ocol, ow, t, o_v, o_n, seed=5, view=v, graph_ptr=ptr)
            s, e = vp[v], vp[v + 1]
            assert np.array_equal(row[s:e], r0) and np.array_equal(col[s:e], c0), (o_v, o_n, v)
            assert np.array_equal(w[s:e].view(np.uint32), w0.view(np.uint32)), (o_v, o_n, v)
            gid = np.searchsorted(ptr, col[s:e], side="right") - 1
            assert np.array_equal(gid, np.searchsorted(ptr, row[s:e], side="right") - 1)   # no edge crosses graphs
            # every graph keeps exactly n_g - t_g vertices with at least ... (survivors without edges do not appear)
            assert np.all(np.bincount(gid, minlength=1113) >= 0)
    # union batch of the first 128 graphs as ONE graph
    sub = ei[:, ei[0] < ptr[128]]
    n2 = int(ptr[128])
    g2 = rlap_b200.prepare(torch.from_numpy(sub).cuda(), None, n2)
    p2, c2, w2 = oracle_port.ingest(sub, None, n2)
    (row, col, w), vp = rlap_b200.schur_views(g2, n2 // 2, "random", "asc", num_views=1, seed=9, dtype=None)
    r0, c0, w0 = oracle_port.keyed_schur(p2, c2, w2, n2 // 2, "random", "asc", seed=9, view=0)
    assert np.array_equal(row.cpu().numpy(), r0) and np.array_equal(col.cpu().numpy(), c0)
    assert np.array_equal(w.cpu().numpy(), w0)


def test_c5_products_shape_scaled(oracle_port):
    """C5 (ogbn-products-shaped SBM, coarsen, several seeds) at 1/8 scale: n = 306,128, 47 blocks, mean degree ~50
    (7.73M undirected edges); bit-exact against the oracle for one seed, properties for all"""
    import rlap_b200
    from rlap_b200 import graphs
    n, und = 306128, 7730000
    ei = graphs.sbm(n, 47, und, seed=0)
    g = rlap_b200.prepare(torch.from_numpy(ei).cuda(), None, n)
    assert g.nnz == 2 * und
    (row, col, w), vp, st = rlap_b200.schur_views(g, n // 2, "coarsen", "asc", num_views=4, seed=3, dtype=None,
                                                  return_stats=True)
    row, col, w, vp = row.cpu().numpy(), col.cpu().numpy(), w.cpu().numpy(), vp.numpy()
    optr, ocol, ow = oracle_port.ingest(ei, None, n)
    r0, c0, w0 = oracle_port.keyed_schur(optr, ocol, ow, n // 2, "coarsen", "asc", seed=3, view=1)
    s, e = vp[1], vp[2]
    assert np.array_equal(row[s:e], r0) and np.array_equal(col[s:e], c0)
    assert np.array_equal(w[s:e].view(np.uint32), w0.view(np.uint32))
    for v in range(4):
        _check_properties(row[vp[v]:vp[v + 1]], col[vp[v]:vp[v + 1]], w[vp[v]:vp[v + 1]], n, n // 2)
