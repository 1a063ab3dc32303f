import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch
import rlap_b200
from rlap_b200 import graphs
n=169343
ei = graphs.barabasi_albert(n,7,seed=0)
g = rlap_b200.prepare(torch.from_numpy(ei).cuda(), None, n)
for ov in ["degree","random"]:
  for V in [1,16,64]:
    for _ in range(2):
        out,vp,s = rlap_b200.schur_views(g, n//2, ov, "asc", num_views=V, seed=1, dtype=None, return_stats=True)
    print(ov, V, {k:v for k,v in s.items() if k.startswith('t_') or k in ('elim_us','rounds')})
