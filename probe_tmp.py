import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch
import rlap_b200
from rlap_b200 import graphs
n=169343
ei = graphs.barabasi_albert(n,7,seed=0)
g = rlap_b200.prepare(torch.from_numpy(ei).cuda(), None, n)
for ov in (sys.argv[1:] or ["degree","coarsen","random"]):
  for V in [1,16,64]:
    ts=[]
    for _ in range(3):
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True); a.record()
        out,vp,s = rlap_b200.schur_views(g, n//2, ov, "asc", num_views=V, seed=1, dtype=None, return_stats=True)
        b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print(ov, V, f"{min(ts):.2f} ms -> {V/min(ts)*1e3:.0f} views/s", {k:v for k,v in s.items() if k.startswith('t_') or k in ('elim_us','emit_count_us','rounds')})
