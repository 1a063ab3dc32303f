"""One prepare + two schur_views calls on the bench graph (arxiv-shaped BA): the command the ncu captures in profiles/ wrap.

    python tools/profile_one_call.py [o_v] [views]          # RLAP_GROUPS=1: all views in one cooperative launch
"""
import sys, os; sys.path.insert(0, '/root/repo')
import numpy as np, torch
import rlap_b200
from rlap_b200 import graphs
n=169343
ei = np.load('/tmp/ba.npy') if os.path.exists('/tmp/ba.npy') else graphs.barabasi_albert(n,7,seed=0)
np.save('/tmp/ba.npy', ei)
ov = sys.argv[1] if len(sys.argv)>1 else "degree"
V = int(sys.argv[2]) if len(sys.argv)>2 else 64
g = rlap_b200.prepare(torch.from_numpy(ei).cuda(), None, n)
for _ in range(2):
    out,vp,s = rlap_b200.schur_views(g, n//2, ov, "asc", num_views=V, seed=1, dtype=None, return_stats=True)
torch.cuda.synchronize()
print(s)
