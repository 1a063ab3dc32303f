"""Per-round phase times of one view group (debug build: RLAP_DEBUG_BUILD=1 python -m rlap_b200._build).

    python tools/probe_rounds.py degree 64
"""
import sys, os; sys.path.insert(0, '/root/repo')
import numpy as np, torch
import rlap_b200
from rlap_b200 import graphs, _native
n = 169343
ei = graphs.barabasi_albert(n, 7, seed=0)
g = rlap_b200.prepare(torch.from_numpy(ei).cuda(), None, n)
ov = sys.argv[1] if len(sys.argv) > 1 else "degree"
V = int(sys.argv[2]) if len(sys.argv) > 2 else 64
import rlap_b200.ops as ops
for rep in range(2):
    ops._DEBUG_FLAGS = 512 if rep == 1 else 0
    out, vp, s = rlap_b200.schur_views(g, n // 2, ov, "asc", num_views=V, seed=1, dtype=None, return_stats=True)
    torch.cuda.synchronize()
print(s)
