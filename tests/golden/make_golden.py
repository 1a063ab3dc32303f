"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built from
/root/reference by `make -C oracle ref`). Run in the build container only:

    python tests/golden/make_golden.py

Each file holds the input edge list and the reference's [E',3] float64 output for one
(graph, o_v, o_n) with the injected seeds recorded (sample_seed = the reference's default
mt19937_64 seed 5489; rd_seed feeds the stand-in for std::random_device)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from rlap_b200 import graphs  # noqa: E402
from tests import util  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    assert ref.available(), "build oracle/_ref first (make -C oracle ref)"
    cases = [("ba100", graphs.barabasi_albert(100, 50, seed=1), 100, 50, None),
             ("ba100w", graphs.barabasi_albert(100, 50, seed=2), 100, 50, "w"),
             ("sbm300", graphs.sbm(300, 4, 900, seed=5), 300, 90, None)]
    for name, ei, n, t, wmode in cases:
        w = util.sym_weights(ei).astype(np.float64) if wmode else None
        info = util.edge_info(ei, w)
        for o_v, o_n in util.COMBOS:
            out = ref.approximate_cholesky(info, n, t, o_v, o_n, sample_seed=5489, rd_seed=17)
            np.savez_compressed(os.path.join(HERE, f"{name}_{o_v}_{o_n}.npz"), edge_info=info, n=n, t=t,
                                sample_seed=5489, rd_seed=17, out=out)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
