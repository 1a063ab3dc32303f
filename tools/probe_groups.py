"""Elimination time against the number of views and view groups (debug build reads RLAP_GROUPS per call).

    python tools/probe_groups.py degree 64:32,64:64,128:128,296:296
"""
import sys, os; sys.path.insert(0, '/root/repo')
import numpy as np, torch
import rlap_b200
from rlap_b200 import graphs
n = 169343
ei = graphs.barabasi_albert(n, 7, seed=0)
g = rlap_b200.prepare(torch.from_numpy(ei).cuda(), None, n)
ov = sys.argv[1] if len(sys.argv) > 1 else "degree"
combos = [tuple(int(x) for x in c.split(':')) for c in (sys.argv[2] if len(sys.argv) > 2 else "64:32,64:64").split(',')]
for V, K in combos:
    if K > 0:
        os.environ['RLAP_GROUPS'] = str(K)
    else:
        os.environ.pop('RLAP_GROUPS', None)
    best = None
    for rep in range(3):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); a.record()
        out, vp, s = rlap_b200.schur_views(g, n // 2, ov, "asc", num_views=V, seed=1, dtype=None, return_stats=True)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        if best is None or s['elim_us'] < best[0]:
            best = (s['elim_us'], ms, s)
        del out
    s = best[2]
    print(f"{ov} V={V} K={K}: elim {best[0]} us = {best[0]/V:.1f} us/view -> {V/best[0]*1e6:.0f} views/s (elim only); call {best[1]:.2f} ms; "
          f"A {s['t_phaseA_us']} B {s['t_phaseB_us']} C {s['t_phaseC_us']} D {s['t_elim_warp_us']} D2 {s['t_elim_block_us']} emit_count {s['emit_count_us']}", flush=True)
