"""Pins the oracle (oracle/rlap_oracle.cc, ref mode) to the reference:
 * against the committed golden vectors generated from the UNMODIFIED reference C++
   (tests/golden/make_golden.py) - bit for bit, all 9 o_v x o_n combinations;
 * against oracle/_ref itself when that library is present (this container / shipped .so)."""
import glob
import os

import numpy as np
import pytest

from tests import util

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def test_golden_files_present():
    assert len(GOLD) == 27


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_ref_mode_matches_golden(oracle_port, path):
    z = np.load(path)
    name = os.path.basename(path)[:-4]
    _, o_v, o_n = name.rsplit("_", 2)
    out = oracle_port.ref_approximate_cholesky(z["edge_info"], int(z["n"]), int(z["t"]), o_v, o_n,
                                               sample_seed=int(z["sample_seed"]), rd_seed=int(z["rd_seed"]))
    assert out.dtype == np.float64 and out.shape == z["out"].shape
    assert np.array_equal(out, z["out"])  # bit exact, including the float64 weights


@pytest.mark.parametrize("o_v,o_n", util.COMBOS)
def test_ref_mode_matches_reference_build(oracle_port, o_v, o_n):
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    from rlap_b200 import graphs
    for n, ei, t in ((2708, graphs.sbm(2708, 7, 5278, seed=0), 812), (500, graphs.barabasi_albert(500, 4, seed=9), 250)):
        for w in (None, util.sym_weights(ei).astype(np.float64)):
            info = util.edge_info(ei, w)
            a = ref.approximate_cholesky(info, n, t, o_v, o_n, sample_seed=77, rd_seed=5)
            b = oracle_port.ref_approximate_cholesky(info, n, t, o_v, o_n, sample_seed=77, rd_seed=5)
            assert np.array_equal(a, b)


def test_reference_determinism_degree_order(oracle_port):
    """o_v=degree with o_n in {asc, desc} is fully deterministic in the reference (default-seeded
    mt19937_64): the default call reproduces itself"""
    z = np.load([p for p in GOLD if p.endswith("ba100_degree_asc.npz")][0])
    a = oracle_port.ref_approximate_cholesky(z["edge_info"], 100, 50, "degree", "asc", rd_seed=1)
    b = oracle_port.ref_approximate_cholesky(z["edge_info"], 100, 50, "degree", "asc", rd_seed=2)
    assert np.array_equal(a, b) and np.array_equal(a, z["out"])


def test_reference_test_suite_contract(oracle_port):
    """the reference's own test (tests/test_rlap.py:23-65) restated on the oracle: float64 output and a
    symmetric unweighted adjacency; additionally exactly n - num_remove surviving nodes"""
    from rlap_b200 import graphs
    for seed in range(5):
        ei = graphs.barabasi_albert(100, 50, seed=seed)
        out = oracle_port.ref_approximate_cholesky(util.edge_info(ei), 100, 50, "random", "asc", rd_seed=seed)
        assert out.dtype == np.float64
        A = np.zeros((100, 100))
        A[out[:, 0].astype(int), out[:, 1].astype(int)] = 1
        assert (A == A.T).all()
        assert np.unique(out[:, :2]).shape[0] == 50


def test_asymmetric_input_is_an_error(oracle_port):
    info = np.array([[0, 1, 1.0], [1, 0, 1.0], [2, 0, 1.0]])
    with pytest.raises(ValueError):
        oracle_port.ref_approximate_cholesky(info, 3, 1, "random", "asc")
