"""rlap_b200/csrc/introsort.cuh (the tie order of o_n = asc / desc among more than 16 neighbours) is
plain C++: compiled for the host here and pinned against libstdc++'s std::sort itself with the reference's comparators
(preconditioner.cc:295-303), on tie-heavy inputs of 1 .. 70 000 elements and on median-of-three killer sequences
that drive the loop into its heap-sort branch."""
import os
import subprocess
import sys

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
