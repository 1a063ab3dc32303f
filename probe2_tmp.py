import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch
import rlap_b200
from rlap_b200 import graphs
for rep in range(2):
  for n in [3000, 50000, 169343]:
    ei = graphs.barabasi_albert(n,7,seed=rep)
    g = rlap_b200.prepare(torch.from_numpy(ei).cuda(), None, n)
    for ov in ["degree","coarsen","random"]:
        for V in [1,4,16]:
            try:
                out,vp,s = rlap_b200.schur_views(g, n//2, ov, "asc", num_views=V, seed=1, dtype=None, return_stats=True)
                torch.cuda.synchronize()
            except Exception as ex:
                print(n, ov, V, "FAIL", str(ex)[:150], flush=True); sys.exit(1)
    print(rep, n, "ok", flush=True)
