"""Timing probe: schur_views on the bench graph for lists of o_v, view counts and debug flags.

    python tools/probe_views.py degree,coarsen 16,64 0        # RLAP_GROUPS=K overrides the number of view groups
"""
import sys, os; sys.path.insert(0, '/root/repo')
import numpy as np, torch
import rlap_b200
from rlap_b200 import graphs
n=169343
ei = graphs.barabasi_albert(n,7,seed=0)
g = rlap_b200.prepare(torch.from_numpy(ei).cuda(), None, n)
def run(ov, V, flags, reps=3):
    os.environ['RLAP_DEBUG_FLAGS']=str(flags)
    ts=[]
    for _ in range(reps):
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True); a.record()
        out,vp,s = rlap_b200.schur_views(g, n//2, ov, "asc", num_views=V, seed=1, dtype=None, return_stats=True)
        b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print(ov, V, "flags", flags, f"{min(ts):.2f} ms -> {V/min(ts)*1e3:.0f} views/s", {k:v for k,v in s.items() if k.startswith('t_') or k in ('elim_us','emit_count_us','rounds')}, flush=True)
args = sys.argv[1:]
ovs = args[0].split(',') if args else ["degree"]
Vs = [int(x) for x in args[1].split(',')] if len(args)>1 else [64]
fl = [int(x) for x in args[2].split(',')] if len(args)>2 else [0]
for ov in ovs:
    for V in Vs:
        for f in fl:
            run(ov, V, f)
# ingest timing
eid = torch.from_numpy(ei).cuda()
for _ in range(3):
    a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True); a.record()
    g2 = rlap_b200.prepare(eid, None, n)
    b.record(); torch.cuda.synchronize(); print("prepare ms", a.elapsed_time(b))
import time
for _ in range(3):
    torch.cuda.synchronize(); t0=time.perf_counter()
    g2 = rlap_b200.prepare(eid, None, n)
    out,vp,s = rlap_b200.schur_views(g2, n//2, "degree", "asc", num_views=64, seed=1, dtype=None, return_stats=True)
    torch.cuda.synchronize(); print("prepare+views wall ms", (time.perf_counter()-t0)*1e3, s['elim_us'], s['emit_count_us'])
