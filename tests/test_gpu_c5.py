"""C5 (BASELINE.json configs[4]): ogbn-products-shaped SBM, 2 449 029 nodes / 123.7 M directed edges, o_v = coarsen,
64 seeds, checked against the EXACT Schur complement.

An exact Schur complement of a 2.4 M-vertex graph cannot be formed, so the check is matrix free:
for a probe vector x on the surviving vertices C of a view, x^T SC x = min over y of [x; y]^T L [x; y], i.e. one
conjugate-gradient solve on L_FF (float64, on the GPU) - compared with the view's own quadratic form
x^T L~ x = sum over its rows of w (x_r - x_c)^2 / 2. The statistic is the ratio of the two, averaged over the seeds.

Coarsening is unbiased for a single elimination but only approximately over many dependent ones:
the REFERENCE's own ratio at this shape (mean degree 50.5, half the vertices removed) is not 1 but about 1.09, and that
constant - measured in the same test on a scaled replica with the reference restatement (oracle ref mode, bit-identical
to the unmodified reference) - is what the CUDA path must reproduce, on the replica and at full size."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _adjacency(ei: torch.Tensor, n: int):
    """unit-weight adjacency of a symmetric edge_index (sorted by (row, col)) as a float64 CSR tensor + degrees"""
    row, col = ei[0], ei[1]
    crow = torch.zeros(n + 1, dtype=torch.int64, device=ei.device)
    crow[1:] = torch.cumsum(torch.bincount(row, minlength=n), 0)
    A = torch.sparse_csr_tensor(crow, col, torch.ones(col.numel(), dtype=torch.float64, device=ei.device), size=(n, n))
    return A, (crow[1:] - crow[:-1]).double()


def _exact_quadratic(A, deg, keep: torch.Tensor, X: torch.Tensor, tol=1e-11, maxiter=400):
    """x_C^T SC x_C for every column of X (values outside C are ignored), SC = exact Schur complement of L = D - A onto
    the vertices flagged in `keep`. One CG per column on L_FF, run side by side."""
    mC = keep.double().unsqueeze(1)
    mF = 1.0 - mC
    xC = X * mC
    AxC = torch.sparse.mm(A, xC)
    rhs = AxC * mF                                           # -L_FC x_C = A_FC x_C
    qCC = (deg.unsqueeze(1) * xC * xC).sum(0) - (xC * AxC).sum(0)

    def op(y):
        return (deg.unsqueeze(1) * y - torch.sparse.mm(A, y)) * mF

    y = torch.zeros_like(rhs)
    r = rhs.clone()
    p = r.clone()
    rs = (r * r).sum(0)
    rs0 = rs.clone()
    for _ in range(maxiter):
        Ap = op(p)
        alpha = rs / (p * Ap).sum(0).clamp_min(1e-300)
        y += alpha * p
        r -= alpha * Ap
        rs_new = (r * r).sum(0)
        if bool((rs_new <= tol * tol * rs0).all()):
            break
        p = r + (rs_new / rs) * p
        rs = rs_new
    return qCC - (rhs * y).sum(0)


def _view_quadratic(row, col, w, X):
    d = X[row.long()] - X[col.long()]
    return 0.5 * (w.double().unsqueeze(1) * d * d).sum(0)


def _ratios_gpu(g, ei_dev, n, t, seeds, views_per_call, X, A, deg, seed=3):
    import rlap_b200
    out = []
    for v0 in range(0, seeds, views_per_call):
        (row, col, w), vp = rlap_b200.schur_views(g, t, "coarsen", "asc", num_views=views_per_call, seed=seed, view_base=v0,
                                                  dtype=None)
        for k in range(views_per_call):
            s, e = int(vp[k]), int(vp[k + 1])
            keep = torch.zeros(n, dtype=torch.bool, device=row.device)
            keep[col[s:e].long()] = True
            assert int(keep.sum()) == n - t, "the shape is connected: every survivor keeps an edge"
            out.append((_view_quadratic(row[s:e], col[s:e], w[s:e], X) / _exact_quadratic(A, deg, keep, X)).cpu().numpy())
        del row, col, w
    return np.array(out)


def test_c5_unbiasedness_statistic_matches_reference_on_replica(oracle_port):
    """scaled replica (n = 20 000, same block count and mean degree): the CUDA path's quadratic-form ratio against the
    exact Schur complement equals the reference's (oracle ref mode) within the sampling error of 16 seeds each"""
    import rlap_b200
    from rlap_b200 import graphs
    from tests import util
    n, und, K = 20000, 505000, 16
    ei = graphs.sbm(n, 47, und, seed=0)
    eid = torch.from_numpy(ei).cuda()
    A, deg = _adjacency(eid, n)
    X = torch.randn(n, 2, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
    g = rlap_b200.prepare(eid, None, n)
    r_gpu = _ratios_gpu(g, eid, n, n // 2, K, 8, X, A, deg)
    info = util.edge_info(ei)
    r_ref = []
    for s in range(K):
        o = oracle_port.ref_approximate_cholesky(info, n, n // 2, "coarsen", "asc", sample_seed=1000 + s, rd_seed=77 + s)
        row, col, w = (torch.from_numpy(np.ascontiguousarray(o[:, j])).cuda() for j in range(3))
        keep = torch.zeros(n, dtype=torch.bool, device="cuda")
        keep[col.long()] = True
        r_ref.append((_view_quadratic(row, col, w, X) / _exact_quadratic(A, deg, keep, X)).cpu().numpy())
    r_ref = np.array(r_ref)
    mg, mr = r_gpu.mean(), r_ref.mean()
    sg, sr = r_gpu.std() / np.sqrt(r_gpu.size), r_ref.std() / np.sqrt(r_ref.size)
    print(f"C5 replica: quadratic-form ratio vs exact SC: CUDA {mg:.4f} +- {sg:.4f}, reference {mr:.4f} +- {sr:.4f}")
    assert 1.0 < mr < 1.2, "the reference's coarsening bias at this shape (measured 1.09)"
    assert abs(mg - mr) <= 4.0 * np.hypot(sg, sr) + 2e-3, (mg, mr, sg, sr)


def test_c5_products_shape_full_size_64_seeds(oracle_port):
    """the stated size: n = 2 449 029, 61 859 140 undirected edges, coarsen, num_remove = 50 %, 64 seeds.
    Memory: one view needs the fill pool (2 nnz int4 = 3.96 GB), staging (0.99 GB) and ~0.15 GB of per-vertex state;
    4 views per call = 20.4 GB beside the 2 GB edge list, the 2.5 GB ingest workspace and the float64 adjacency of the
    checker (3 GB)."""
    import rlap_b200
    from rlap_b200 import graphs
    n, und = 2449029, 61859140
    t = n // 2
    eid = graphs.sbm_torch(n, 47, und, seed=0, device="cuda")
    g = rlap_b200.prepare(eid, None, n)
    assert g.nnz == 2 * und
    A, deg = _adjacency(eid, n)
    X = torch.randn(n, 2, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    r = _ratios_gpu(g, eid, n, t, 64, 4, X, A, deg)
    m, sd = r.mean(), r.std()
    print(f"C5 full size: quadratic-form ratio vs exact SC over 64 seeds x 2 probes: {m:.4f}, std {sd:.5f}, "
          f"sem {sd / np.sqrt(r.size):.5f}")
    # the reference's constant at this shape is 1.090 +- 0.002 (replica test above, 16 seeds; tests/README of the number:
    # DESIGN.md §5); a view of 2.4 M vertices averages over 100 x more stars than the replica, so its own spread is tiny
    assert abs(m - 1.090) <= 0.01, m
    assert sd <= 0.01, sd
    # one view bit for bit against the sequential oracle at the stated size
    del A
    (row, col, w), vp = rlap_b200.schur_views(g, t, "coarsen", "asc", num_views=1, seed=3, view_base=5, dtype=None)
    ei = eid.cpu().numpy()
    optr, ocol, ow = oracle_port.ingest(ei, None, n)
    r0, c0, w0 = oracle_port.keyed_schur(optr, ocol, ow, t, "coarsen", "asc", seed=3, view=5)
    assert np.array_equal(row.cpu().numpy(), r0) and np.array_equal(col.cpu().numpy(), c0)
    assert np.array_equal(w.cpu().numpy().view(np.uint32), w0.view(np.uint32))
