"""GPU parity tests proper: the CUDA path (through the C ABI, via rlap_b200.ops) against the
in-repo oracle's keyed mode on the same seeded inputs. Bit-exact: rows, columns and fp32 weights."""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def _gpu_graph(ei, w, n, gptr):
    import rlap_b200
    return rlap_b200.prepare(torch.from_numpy(ei).cuda(), None if w is None else torch.from_numpy(w).cuda(), n,
                             graph_ptr=gptr)


@pytest.mark.parametrize("weighted", [False, True])
def test_ingest_matches_oracle(oracle_port, weighted):
    rng = np.random.default_rng(0)
    for name, ei, n, gptr, t in util.small_cases():
        w = util.sym_weights(ei) if weighted else None
        # shuffle the edge order and add duplicates + explicit zeros: the coalesced result must not change
        E = ei.shape[1]
        dup = rng.integers(0, E, size=E // 10)
        both = np.concatenate([dup, dup + 0])  # keep symmetric: add each picked edge and its twin
        rev = {(int(a), int(b)): k for k, (a, b) in enumerate(zip(ei[0], ei[1]))}
        twin = np.array([rev[(int(ei[1, k]), int(ei[0, k]))] for k in dup])
        extra = np.concatenate([dup, twin])
        ei2 = np.concatenate([ei, ei[:, extra]], axis=1)
        w2 = None if w is None else np.concatenate([w, w[extra]])
        zero = ei[:, :5]
        ei2 = np.concatenate([ei2, zero], axis=1)
        w2 = np.concatenate([np.ones(ei2.shape[1] - 5, dtype=np.float32) if w2 is None else w2,
                             np.zeros(5, dtype=np.float32)])
        perm = rng.permutation(ei2.shape[1])
        ei2, w2 = np.ascontiguousarray(ei2[:, perm]), np.ascontiguousarray(w2[perm])
        optr, ocol, ow = oracle_port.ingest(ei2, w2, n)
        g = _gpu_graph(ei2, w2, n, gptr)
        assert g.nnz == ocol.shape[0], name
        assert np.array_equal(g.ptr.cpu().numpy().astype(np.int64), optr), name
        assert np.array_equal(g.col.cpu().numpy()[: g.nnz], ocol), name
        assert np.array_equal(g.w.cpu().numpy()[: g.nnz], ow), name  # bit exact fp32 sums


@pytest.mark.parametrize("o_v,o_n", util.COMBOS)
@pytest.mark.parametrize("weighted", [False, True])
def test_views_match_oracle_bit_exact(oracle_port, o_v, o_n, weighted):
    import rlap_b200
    for name, ei, n, gptr, t in util.small_cases():
        w = util.sym_weights(ei) if weighted else None
        optr, ocol, ow = oracle_port.ingest(ei, w, n)
        g = _gpu_graph(ei, w, n, gptr)
        V = 3
        (row, col, wt), vp = rlap_b200.schur_views(g, t, o_v, o_n, num_views=V, seed=1234, view_base=5, dtype=None)
        row, col, wt, vp = row.cpu().numpy(), col.cpu().numpy(), wt.cpu().numpy(), vp.numpy()
        for v in range(V):
            r0, c0, w0 = oracle_port.keyed_schur(optr, ocol, ow, t, o_v, o_n, seed=1234, view=5 + v, graph_ptr=gptr)
            s, e = vp[v], vp[v + 1]
            assert e - s == r0.shape[0], (name, v, e - s, r0.shape[0])
            assert np.array_equal(row[s:e], r0), (name, v)
            assert np.array_equal(col[s:e], c0), (name, v)
            assert np.array_equal(wt[s:e].view(np.uint32), w0.view(np.uint32)), (name, v)


@pytest.mark.parametrize("o_v", ["random", "degree", "coarsen"])
def test_full_clique_is_exact_schur_complement(oracle_port, o_v):
    """sampling replaced by full clique elimination == exact Schur complement (fp32, 1e-5 relative)"""
    import rlap_b200
    from rlap_b200 import graphs
    n = 100
    ei = graphs.barabasi_albert(n, 50, seed=1)
    w = util.sym_weights(ei)
    g = _gpu_graph(ei, w, n, None)
    (row, col, wt), vp = rlap_b200.schur_views(g, 50, o_v, "asc", seed=7, full_clique=True, dtype=None)
    row, col, wt = row.cpu().numpy(), col.cpu().numpy(), wt.cpu().numpy()
    L0 = util.laplacian(ei[0], ei[1], w, n)
    keep = np.unique(col)
    assert keep.shape[0] == 50
    Ls = util.laplacian(row, col, wt, n)[np.ix_(keep, keep)]
    ex = util.exact_schur(L0, keep)
    assert np.linalg.norm(Ls - ex) / np.linalg.norm(ex) < 1e-5
    # and bit-exact against the oracle's full-clique mode
    optr, ocol, ow = oracle_port.ingest(ei, w, n)
    r0, c0, w0 = oracle_port.keyed_schur(optr, ocol, ow, 50, o_v, "asc", seed=7, flags=oracle_port.FLAG_FULL_CLIQUE)
    assert np.array_equal(row, r0) and np.array_equal(col, c0)
    assert np.array_equal(wt.view(np.uint32), w0.view(np.uint32))


def test_reference_api_contract():
    """the reference's own test (tests/test_rlap.py:23-65): dtype float64 and a symmetric
    unweighted adjacency over exactly n - num_remove surviving nodes; plus the README call shape"""
    import rlap_b200
    from rlap_b200 import graphs
    for seed in range(3):
        ei = torch.from_numpy(graphs.barabasi_albert(100, 50, seed=seed))
        for weights in (None, torch.ones((1, ei.shape[1]))):
            out = rlap_b200.ops.approximate_cholesky(edge_index=ei, edge_weights=weights, num_nodes=100, num_remove=50,
                                                     o_v="random", o_n="asc")
            assert out.dtype == torch.double and out.device == ei.device and out.shape[1] == 3
            A = torch.zeros(100, 100)
            A[out[:, 0].long(), out[:, 1].long()] = 1
            assert torch.equal(A, A.t())
            assert torch.unique(out[:, :2]).numel() == 50
    a = torch.randn(100, 100, dtype=torch.double)
    assert torch.allclose(rlap_b200.ops.identity(a), a, atol=1e-8)


def test_edge_cases(oracle_port):
    import rlap_b200
    from rlap_b200 import graphs
    ei = graphs.barabasi_albert(60, 3, seed=4)
    n = 64  # 4 isolated vertices at the end: they consume removal slots (A.3)
    optr, ocol, ow = oracle_port.ingest(ei, None, n)
    g = _gpu_graph(ei, None, n, None)
    for t in (0, 1, 4, 63, 64, 1000):
        for o_v in ("random", "degree", "coarsen"):
            (row, col, wt), vp = rlap_b200.schur_views(g, t, o_v, "asc", seed=3, dtype=None)
            r0, c0, w0 = oracle_port.keyed_schur(optr, ocol, ow, t, o_v, "asc", seed=3)
            assert np.array_equal(row.cpu().numpy(), r0) and np.array_equal(col.cpu().numpy(), c0), (t, o_v)
            assert np.array_equal(wt.cpu().numpy(), w0), (t, o_v)
            if t == 0:  # t = 0 returns the coalesced input
                assert row.shape[0] == ei.shape[1]
            if t >= n - 1:
                assert row.shape[0] == 0
    # empty graph
    g0 = rlap_b200.prepare(torch.zeros((2, 0), dtype=torch.long).cuda(), None, 5)
    out, vp = rlap_b200.schur_views(g0, 2, "degree", "asc", seed=1)
    assert out.shape == (0, 3)
    # errors instead of exit(0)
    bad = torch.tensor([[0, 1, 2], [1, 0, 0]]).cuda()
    with pytest.raises(ValueError):
        rlap_b200.prepare(bad, None, 3)              # asymmetric
    with pytest.raises(ValueError):
        rlap_b200.prepare(torch.tensor([[0, 5], [5, 0]]).cuda(), None, 3)   # id out of range
    with pytest.raises(ValueError):
        rlap_b200.prepare(torch.tensor([[1], [1]]).cuda(), None, 3)          # self loop


def test_host_abi_matches_device_path(oracle_port):
    """rlap_approximate_cholesky_host (HOST [E,3] f64 in/out) == device path == oracle"""
    from rlap_b200 import graphs, ops
    ei = graphs.sbm(2708, 7, 5278, seed=0)
    info = util.edge_info(ei)
    out = ops.approximate_cholesky_host(info, 2708, 812, "degree", "asc", seed=99)
    optr, ocol, ow = oracle_port.ingest(ei, None, 2708)
    r0, c0, w0 = oracle_port.keyed_schur(optr, ocol, ow, 812, "degree", "asc", seed=99, view=0)
    assert out.dtype == np.float64 and out.shape == (r0.shape[0], 3)
    assert np.array_equal(out[:, 0], r0) and np.array_equal(out[:, 1], c0) and np.array_equal(out[:, 2], w0.astype(np.float64))


@pytest.mark.parametrize("o_n", ["desc", "asc"])
def test_chained_views_extreme_dynamic_range(oracle_port, o_n):
    """ten chained eliminations (each on the relabelled, weighted output of the previous one, as
    scripts/rlap_vc_spectral.py does): under desc the weights end up spanning ~1e-29 .. 3, so neighbours quantise to
    zero inside a star, fills underflow and stars of 17..140 neighbours go through every tier. Bit-exact at every
    step (regression: a live neighbour with fixed-point weight 0 used to sort behind the padding of a 32-lane tile)."""
    import rlap_b200
    from rlap_b200 import graphs
    n0 = 1000
    ei0 = graphs.barabasi_albert(n0, 5, seed=3)
    for r in (3, 4, 7):
        ei, w, nn = ei0, np.ones(ei0.shape[1], dtype=np.float32), n0
        for k in range(10):
            ptr, col, ww = oracle_port.ingest(ei, w, nn)
            r0, c0, w0 = oracle_port.keyed_schur(ptr, col, ww, 50, "random", o_n, seed=100 * r + k)
            g = rlap_b200.prepare(torch.from_numpy(np.ascontiguousarray(ei)).cuda(), torch.from_numpy(w).cuda(), nn)
            (row, c2, w2), vp = rlap_b200.schur_views(g, 50, "random", o_n, seed=100 * r + k, dtype=None)
            assert np.array_equal(row.cpu().numpy(), r0) and np.array_equal(c2.cpu().numpy(), c0), (r, k)
            assert np.array_equal(w2.cpu().numpy().view(np.uint32), w0.view(np.uint32)), (r, k)
            nodes = np.unique(np.concatenate([r0, c0]))
            ei = np.stack([np.searchsorted(nodes, r0), np.searchsorted(nodes, c0)])
            w, nn = w0, len(nodes)


def test_live_counters_are_exact():
    """compute-sanitizer is closed on this pool, so the library carries its own consistency check
    (RLAP_FLAG_CHECK_LIVE): after the elimination, the number of live entries the emission finds for every surviving
    vertex must equal the live counter the elimination maintained (a mismatch means a lost or phantom entry)."""
    import rlap_b200
    for name, ei, n, gptr, t in util.small_cases():
        for w in (None, util.sym_weights(ei, 1e-6, 10.0)):
            g = _gpu_graph(ei, w, n, gptr)
            for o_v, o_n in util.COMBOS:
                _, _, st = rlap_b200.schur_views(g, t, o_v, o_n, num_views=4, seed=3, dtype=None, return_stats=True,
                                                 check_live=True)
                assert st["check_mismatches"] == 0, (name, o_v, o_n)


def test_torch_dispatcher_op_schema():
    """torch.ops.extension_cpp.approximate_cholesky keeps the reference's schema (py_api_binder.cc:80-83) and
    dispatches on CUDA and CPU tensors"""
    import rlap_b200
    from rlap_b200 import graphs
    assert rlap_b200.register_torch_op()
    ei = graphs.barabasi_albert(100, 50, seed=0)
    info = torch.from_numpy(util.edge_info(ei))
    for t in (info, info.cuda()):
        out = torch.ops.extension_cpp.approximate_cholesky.default(edge_info=t, num_nodes=100, num_remove=50,
                                                                   o_v="random", o_n="asc")
        assert out.dtype == torch.double and out.shape[1] == 3 and out.device == t.device
        assert torch.unique(out[:, :2]).numel() == 50
    a = torch.randn(8, 8, dtype=torch.double)
    assert torch.equal(torch.ops.extension_cpp.identity.default(a=a), a)
    schema = str(torch.ops.extension_cpp.approximate_cholesky.default._schema)
    assert "Tensor edge_info, int num_nodes, int num_remove, str o_v, str o_n" in schema


def test_dirty_workspace_and_pool_retry_are_harmless(oracle_port):
    """the workspace is never assumed to be zero: results do not depend on what the allocator hands back (a failed
    run used to leave unwritten pool slots for the emission to read), and a pool overflow is retried cleanly"""
    import rlap_b200
    from rlap_b200 import graphs
    rng = np.random.default_rng(0)
    n = 100
    ei = graphs.barabasi_albert(n, 50, seed=1)
    w = util.sym_weights(ei)
    g = _gpu_graph(ei, w, n, None)
    optr, ocol, ow = oracle_port.ingest(ei, w, n)
    for o_v in ("random", "degree", "coarsen"):
        for full in (True, False):
            flags = oracle_port.FLAG_FULL_CLIQUE if full else 0
            r0, c0, w0 = oracle_port.keyed_schur(optr, ocol, ow, 50, o_v, "asc", seed=7, flags=flags)
            for rep in range(6):
                junk = torch.randint(0, 255, (int(rng.integers(1, 48)) << 20,), dtype=torch.uint8, device="cuda")
                del junk
                # a pool of 64 entries always overflows first: the retry path runs on recycled memory
                (row, col, wt), vp = rlap_b200.schur_views(g, 50, o_v, "asc", seed=7, full_clique=full, dtype=None,
                                                           pool_cap=64 if rep % 2 else 0)
                assert np.array_equal(row.cpu().numpy(), r0) and np.array_equal(col.cpu().numpy(), c0), (o_v, full, rep)
                assert np.array_equal(wt.cpu().numpy().view(np.uint32), w0.view(np.uint32)), (o_v, full, rep)
    # the same through the view groups (one concurrent launch per view): every group overflows and is retried
    for o_v in ("random", "degree", "coarsen"):
        (row, col, wt), vp = rlap_b200.schur_views(g, 50, o_v, "asc", num_views=4, seed=7, dtype=None, pool_cap=64)
        row, col, wt, vp = row.cpu().numpy(), col.cpu().numpy(), wt.cpu().numpy(), vp.numpy()
        for v in range(4):
            r0, c0, w0 = oracle_port.keyed_schur(optr, ocol, ow, 50, o_v, "asc", seed=7, view=v)
            s, e = vp[v], vp[v + 1]
            assert np.array_equal(row[s:e], r0) and np.array_equal(col[s:e], c0), (o_v, v)
            assert np.array_equal(wt[s:e].view(np.uint32), w0.view(np.uint32)), (o_v, v)


@pytest.mark.parametrize("o_v", ["degree", "coarsen"])
def test_many_desynchronised_views_match_oracle(oracle_port, o_v):
    """48 views of one graph run their bucket levels out of step (some rescan while others work from their low
    lists); every checked view must still be the oracle's, bit for bit"""
    import rlap_b200
    from rlap_b200 import graphs
    n = 20000
    ei = graphs.barabasi_albert(n, 7, seed=5)
    optr, ocol, ow = oracle_port.ingest(ei, None, n)
    g = _gpu_graph(ei, None, n, None)
    V = 48
    (row, col, wt), vp, st = rlap_b200.schur_views(g, n // 2, o_v, "asc", num_views=V, seed=99, dtype=None,
                                                   return_stats=True)
    row, col, wt, vp = row.cpu().numpy(), col.cpu().numpy(), wt.cpu().numpy(), vp.numpy()
    for v in (0, 1, 17, 31, 47):
        r0, c0, w0 = oracle_port.keyed_schur(optr, ocol, ow, n // 2, o_v, "asc", seed=99, view=v)
        s, e = vp[v], vp[v + 1]
        assert e - s == r0.shape[0], (v, e - s, r0.shape[0])
        assert np.array_equal(row[s:e], r0) and np.array_equal(col[s:e], c0), v
        assert np.array_equal(wt[s:e].view(np.uint32), w0.view(np.uint32)), v


def test_column_pointer_output_matches_packed_rows():
    """schur_views(colptr=True) + expand_cols == the col array of the packed rows (rows and weights unchanged)"""
    import rlap_b200
    from rlap_b200 import graphs, ops
    for name, ei, n, gptr, t in util.small_cases():
        g = _gpu_graph(ei, None, n, gptr)
        (row, col, wt), vp = rlap_b200.schur_views(g, t, "degree", "asc", num_views=5, seed=3, dtype=None)
        (row2, cp, wt2), vp2 = rlap_b200.schur_views(g, t, "degree", "asc", num_views=5, seed=3, dtype=None, colptr=True)
        assert torch.equal(vp, vp2) and torch.equal(row, row2) and torch.equal(wt, wt2), name
        assert cp.shape == (5, n + 1) and cp.dtype == torch.int32
        col2 = ops.expand_cols(cp.cpu(), vp2, threads=4)
        assert torch.equal(col.cpu()[: int(vp[-1])], col2), name


@pytest.mark.gpu
def test_batched_dispatcher_op(oracle_port):
    """torch.ops.extension_cpp.approximate_cholesky_batched: the views of every graph of a batch in one
    dispatcher call, bit-exact against the oracle"""
    import rlap_b200
    from rlap_b200 import graphs
    assert rlap_b200.register_torch_op()
    ei, ptr = graphs.proteins_like_batch(24, seed=7)
    n = int(ptr[-1])
    nr = np.diff(ptr) // 2
    out, vp = torch.ops.extension_cpp.approximate_cholesky_batched.default(
        edge_index=torch.from_numpy(ei).cuda(), edge_weight=None, graph_ptr=torch.from_numpy(ptr),
        num_remove=torch.from_numpy(nr), o_v="degree", o_n="desc", num_views=2, seed=77)
    assert out.dtype == torch.double and out.is_cuda and vp.shape[0] == 3
    optr, ocol, ow = oracle_port.ingest(ei, None, n)
    o = out.cpu().numpy()
    for v in range(2):
        r0, c0, w0 = oracle_port.keyed_schur(optr, ocol, ow, nr, "degree", "desc", seed=77, view=v, graph_ptr=ptr)
        s, e = int(vp[v]), int(vp[v + 1])
        assert np.array_equal(o[s:e, 0], r0) and np.array_equal(o[s:e, 1], c0)
        assert np.array_equal(o[s:e, 2].astype(np.float32).view(np.uint32), w0.view(np.uint32))


@pytest.mark.gpu
def test_two_host_threads_two_streams(oracle_port):
    """Threading row of the boundary: calls issued concurrently from two host threads, each on its own
    CUDA stream, return what the same calls return one after the other"""
    import threading
    import rlap_b200
    ei = util.small_cases()[2][1]          # Cora-shaped SBM
    n, t = 2708, 812
    optr, ocol, ow = oracle_port.ingest(ei, None, n)
    want = {}
    for tid, (o_v, o_n) in enumerate((("degree", "asc"), ("random", "desc"))):
        want[tid] = oracle_port.keyed_schur(optr, ocol, ow, t, o_v, o_n, seed=11 + tid, view=0)
    got, errs = {}, []

    def work(tid, o_v, o_n):
        try:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                g = rlap_b200.prepare(torch.from_numpy(ei).cuda(), None, n)
                for rep in range(6):
                    (r, c, w), vp = rlap_b200.schur_views(g, t, o_v, o_n, num_views=3, seed=11 + tid, dtype=None)
                    got[(tid, rep)] = (r[:int(vp[1])].cpu().numpy(), c[:int(vp[1])].cpu().numpy(), w[:int(vp[1])].cpu().numpy())
        except Exception as ex:   # pragma: no cover
            errs.append(ex)

    th = [threading.Thread(target=work, args=(0, "degree", "asc")), threading.Thread(target=work, args=(1, "random", "desc"))]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errs, errs
    for (tid, rep), (r, c, w) in got.items():
        r0, c0, w0 = want[tid]
        assert np.array_equal(r, r0) and np.array_equal(c, c0) and np.array_equal(w.view(np.uint32), w0.view(np.uint32)), (tid, rep)


@pytest.mark.gpu
def test_unweighted_and_colptr_outputs_match_packed_rows():
    """schur_views(weights=False) and (colptr=True, weights=False) are the packed rows minus what they leave out"""
    import rlap_b200
    name, ei, n, gptr, t = util.small_cases()[1]
    g = _gpu_graph(ei, None, n, gptr)
    (r, c, w), vp = rlap_b200.schur_views(g, t, "degree", "asc", num_views=3, seed=5, dtype=None)
    (r2, c2, w2), vp2 = rlap_b200.schur_views(g, t, "degree", "asc", num_views=3, seed=5, dtype=None, weights=False)
    assert w2 is None and torch.equal(r, r2) and torch.equal(c, c2) and torch.equal(vp, vp2)
    (r3, cp3, w3), vp3 = rlap_b200.schur_views(g, t, "degree", "asc", num_views=3, seed=5, dtype=None, weights=False, colptr=True)
    assert w3 is None and torch.equal(r, r3)
    col = rlap_b200.ops.expand_cols(cp3.cpu(), vp3)
    assert torch.equal(col, c.cpu())


@pytest.mark.gpu
def test_second_device_in_one_process(oracle_port):
    """function attributes, streams and the host mailbox are per device: the same process runs the path on
    cuda:1 after cuda:0 when the box has two GPUs"""
    import rlap_b200
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one box")
    name, ei, n, gptr, t = util.small_cases()[2]
    optr, ocol, ow = oracle_port.ingest(ei, None, n)
    r0, c0, w0 = oracle_port.keyed_schur(optr, ocol, ow, t, "coarsen", "asc", seed=9, view=0)
    for d in (0, 1, 0):
        g = rlap_b200.prepare(torch.from_numpy(ei).to(f"cuda:{d}"), None, n)
        (r, c, w), vp = rlap_b200.schur_views(g, t, "coarsen", "asc", num_views=2, seed=9, dtype=None)
        assert r.device.index == d
        e = int(vp[1])
        assert np.array_equal(r[:e].cpu().numpy(), r0) and np.array_equal(w[:e].cpu().numpy().view(np.uint32), w0.view(np.uint32))


@pytest.mark.gpu
def test_degree_initial_bucket_and_first_pop_match_reference(oracle_port):
    """a7: the initial keys / buckets are the reference's (preconditioner.cc:125-165: key = stored entries of the column,
    bucket head = highest id). With num_remove = 1 the reference pops exactly the head of the lowest non-empty bucket,
    and so must the CUDA path (the last round keeps the highest ids of its independent set); with a larger budget the
    first round of the keyed schedule is the independent set of that bucket, computed here from the degrees alone."""
    import rlap_b200
    for name, ei, n, gptr, t in util.small_cases()[:3]:
        info = util.edge_info(ei)
        deg = np.bincount(ei[1], minlength=n)
        m = deg.min()
        head = int(np.nonzero(deg == m)[0].max())
        g = _gpu_graph(ei, None, n, None)
        for o_v in ("degree", "coarsen"):
            ref = oracle_port.ref_approximate_cholesky(info, n, 1, o_v, "asc")
            gone_ref = set(range(n)) - set(ref[:, 1].astype(np.int64).tolist())
            (row, col, w), vp = rlap_b200.schur_views(g, 1, o_v, "asc", num_views=1, seed=1, dtype=None)
            gone_gpu = set(range(n)) - set(col.cpu().numpy().tolist())
            iso = set(np.nonzero(deg == 0)[0].tolist())
            assert gone_ref - iso == {head} - iso and gone_gpu - iso == gone_ref - iso, (name, o_v, head, gone_ref, gone_gpu)
        # first round of the keyed schedule = bucket members without a higher-id bucket neighbour
        optr, ocol, ow = oracle_port.ingest(ei, None, n)
        order = oracle_port.keyed_schur(optr, ocol, ow, n // 2, "degree", "asc", seed=1, return_order=True)[-1]
        bucket = deg == m
        blocked = np.zeros(n, dtype=bool)
        src, dst = ei[0], ei[1]
        sel = bucket[src] & bucket[dst] & (src > dst)
        blocked[dst[sel]] = True
        first = np.nonzero(bucket & ~blocked)[0]
        if first.size <= n // 2:
            assert np.array_equal(np.nonzero(order == 0)[0], first), name


@pytest.mark.parametrize("o_n", ["asc", "desc"])
def test_tie_order_of_big_stars_matches_std_sort(oracle_port, o_n):
    """o_n = asc / desc ties among more than 16 neighbours stand where libstdc++'s std::sort leaves them (DESIGN.md
    §3.3): one star per tier - 24 neighbours (register tile), 100 (warp, shared memory), 600 (block, shared memory),
    3 000 (global scratch slot, partition loop on a shared-memory copy of the keys) and 6 000 (scratch slot, in place)
    - eliminated first under o_v = random (the centre is the vertex of rank 0), with unit weights (every key equal),
    two-valued weights (long runs of ties) and distinct weights. Bit-exact against the oracle, whose order IS the
    std::sort call of the reference (preconditioner.cc:295-303)."""
    import rlap_b200
    seed = 17
    for leaves in (24, 100, 600, 3000, 6000):
        n = leaves + 1
        c = int(np.argmin(oracle_port.rank_perm(seed, 0, 0, n)))
        others = np.array([v for v in range(n) if v != c], dtype=np.int64)
        # a ring among the leaves as well, so that the output is more than the sampled tree
        ring_a, ring_b = others, np.roll(others, 1)
        src = np.concatenate([np.full(leaves, c), others, ring_a, ring_b])
        dst = np.concatenate([others, np.full(leaves, c), ring_b, ring_a])
        ei = np.stack([src, dst])
        for kind in ("unit", "two", "distinct"):
            if kind == "unit":
                w = None
            else:
                a, b = np.minimum(src, dst), np.maximum(src, dst)
                h = (a * 7919 + b * 104729) % (2 if kind == "two" else 1000003)
                w = (1.0 + h.astype(np.float64) * (1.0 if kind == "two" else 1e-6)).astype(np.float32)
            optr, ocol, ow = oracle_port.ingest(ei, w, n)
            g = _gpu_graph(ei, w, n, None)
            (row, col, wt), vp, st = rlap_b200.schur_views(g, 1, "random", o_n, seed=seed, dtype=None, return_stats=True)
            r0, c0, w0, so = oracle_port.keyed_schur(optr, ocol, ow, 1, "random", o_n, seed=seed, return_stats=True)
            assert so["maxlen"] == leaves, (so, leaves)               # the centre went first
            assert np.array_equal(row.cpu().numpy(), r0) and np.array_equal(col.cpu().numpy(), c0), (leaves, kind)
            assert np.array_equal(wt.cpu().numpy().view(np.uint32), w0.view(np.uint32)), (leaves, kind)
