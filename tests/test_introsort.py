"""rlap_b200/csrc/introsort.cuh (the tie order of o_n = asc / desc among more than 16 neighbours) is
plain C++: compiled for the host here and pinned against libstdc++'s std::sort itself with the reference's comparators
(preconditioner.cc:295-303), on tie-heavy inputs of 1 .. 70 000 elements and on median-of-three killer sequences
that drive the loop into its heap-sort branch."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_introsort_restatement_matches_std_sort(tmp_path):
    exe = str(tmp_path / "introsort_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "native", "introsort_check.cc")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    tag, cases, heap = out.stdout.split()
    assert tag == "ok" and int(cases) > 10000
    assert int(heap) > 0, "the depth-limit (heap sort) branch was never exercised"


def test_keyed_oracle_tie_order_is_std_sort(oracle_port):
    """known answers of libstdc++'s std::sort on equal keys in id order (what the reference's compressColumn leaves for
    unit weights, preconditioner.cc:295-303): identity up to 16 elements, median to the front + pairwise reversal
    of the rest above - a toolchain whose std::sort differs would show up here"""
    import numpy as np
    one = lambda n: np.full(n, 1 << 40, dtype=np.uint64)
    assert oracle_port.star_order(one(16), "asc").tolist() == list(range(16))
    want17 = [8, 16, 15, 14, 13, 12, 11, 10, 9, 0, 7, 6, 5, 4, 3, 2, 1]
    want40 = [30, 21, 22, 23, 24, 25, 26, 27, 28, 29, 20, 31, 32, 33, 34, 35, 36, 37, 38, 39,
              10, 1, 2, 3, 4, 5, 6, 7, 8, 9, 0, 11, 12, 13, 14, 15, 16, 17, 18, 19]
    for o_n in ("asc", "desc"):
        assert oracle_port.star_order(one(17), o_n).tolist() == want17
        assert oracle_port.star_order(one(40), o_n).tolist() == want40
    # distinct keys: plain sorted order, whatever the size
    rng = np.random.default_rng(0)
    q = rng.permutation(300).astype(np.uint64) + 1
    assert np.array_equal(q[oracle_port.star_order(q, "asc")], np.sort(q))
    assert np.array_equal(q[oracle_port.star_order(q, "desc")], np.sort(q)[::-1])
    # o_n = random is untouched by the rule: shuffle keys (all equal here) then id
    assert oracle_port.star_order(one(40), "random").tolist() == list(range(40))


@pytest.mark.parametrize("o_n", ["asc", "desc"])
@pytest.mark.parametrize("leaves", [16, 17, 40, 200])
def test_tie_order_is_the_one_the_reference_samples_in(oracle_port, o_n, leaves):
    """The claimed order, read back from the UNMODIFIED reference's own output. Star K_{1,leaves} with unit weights, the
    centre eliminated first (o_v = random with an injected random_device stream that pops it first): the reference
    gives neighbour j (in its std::sort order) exactly one fill edge, to a LATER neighbour (preconditioner.cc:379-417).
    So in the resulting tree every vertex but the last of the order has exactly one neighbour that stands later in the
    order, and the last has none - for every sampling seed. A wrong order (identity above 16 elements, say) breaks
    that within a few seeds."""
    from oracle import ref
    n = leaves + 1
    centre = None
    for rd in range(4000):
        if int(oracle_port.ref_random_order(n, rd)[0]) == 0:
            centre, rd_seed = 0, rd
            break
    assert centre == 0, "no injected stream pops vertex 0 first"
    others = np.arange(1, n)
    ei = np.stack([np.concatenate([np.zeros(leaves, dtype=np.int64), others]),
                   np.concatenate([others, np.zeros(leaves, dtype=np.int64)])])
    _check_readback(oracle_port, ei, None, leaves, o_n, rd_seed, expect_id_order_refuted=leaves > 16)
    # partial ties: two- and three-valued weights (the order is by weight, the arrangement decides inside a class)
    rng = np.random.default_rng(leaves)
    for nvals in (2, 3):
        wl = rng.integers(1, nvals + 1, size=leaves).astype(np.float64)
        _check_readback(oracle_port, ei, np.concatenate([wl, wl]), leaves, o_n, rd_seed, expect_id_order_refuted=False)


def _check_readback(oracle_port, ei, w, leaves, o_n, rd_seed, expect_id_order_refuted):
    from oracle import ref
    n = leaves + 1
    info = util.edge_info(ei, w)
    q = np.full(leaves, 1 << 40, dtype=np.uint64) if w is None else (w[:leaves] * float(1 << 40)).astype(np.uint64)
    order = oracle_port.star_order(q, o_n) + 1                                          # vertex ids 1..leaves
    pos = np.empty(n, dtype=np.int64)
    pos[order] = np.arange(leaves)
    ident = np.arange(n) - 1                                                            # the id-order hypothesis
    run = ref.approximate_cholesky if ref.available() else oracle_port.ref_approximate_cholesky
    wrong_hypothesis_survives = True
    for sample_seed in range(20):
        out = run(info, n, 1, "random", o_n, sample_seed=1 + sample_seed, rd_seed=rd_seed)
        r, c = out[:, 0].astype(np.int64), out[:, 1].astype(np.int64)
        assert r.shape[0] == 2 * (leaves - 1) and 0 not in set(r.tolist()) | set(c.tolist())
        later = np.bincount(r[pos[c] > pos[r]], minlength=n)[1:]       # per vertex: neighbours later in the order
        want = np.ones(leaves, dtype=np.int64)
        want[order[-1] - 1] = 0
        assert np.array_equal(later, want), (o_n, leaves, sample_seed)
        if expect_id_order_refuted:
            later_id = np.bincount(r[ident[c] > ident[r]], minlength=n)[1:]
            want_id = np.ones(leaves, dtype=np.int64)
            want_id[-1] = 0
            wrong_hypothesis_survives &= bool(np.array_equal(later_id, want_id))
    if expect_id_order_refuted:
        assert not wrong_hypothesis_survives            # the test has teeth: id order is NOT what the reference does
