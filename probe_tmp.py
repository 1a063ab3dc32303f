import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np, torch
import rlap_b200
from rlap_b200 import graphs
n=169343
ei = graphs.barabasi_albert(n,7,seed=0)
eit = torch.from_numpy(ei).cuda()
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts=[]
    for _ in range(reps):
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
if len(sys.argv) > 1:
    ov, V = sys.argv[1], int(sys.argv[2])
    g = rlap_b200.prepare(eit, None, n)
    for _ in range(2):
        out,vp,s = rlap_b200.schur_views(g, n//2, ov, "asc", num_views=V, seed=1, dtype=None, return_stats=True)
    torch.cuda.synchronize(); print(s); sys.exit(0)
print("ingest ms", timeit(lambda: rlap_b200.prepare(eit, None, n)))
g = rlap_b200.prepare(eit, None, n)
for ov,on in [("degree","asc"),("coarsen","asc"),("random","asc")]:
    for V in [1,4,16,64]:
        st={}
        def f():
            out,vp,s = rlap_b200.schur_views(g, n//2, ov, on, num_views=V, seed=1, dtype=None, return_stats=True)
            st.update(s)
        ms = timeit(f, reps=2)
        print(f"{ov}/{on} V={V}: {ms:.2f} ms -> {V/ms*1e3:.1f} views/s  rounds {st['rounds']} maxstar {st['max_star']} elim {st['elim_us']/1e3:.2f} ms emitcount {st['emit_count_us']/1e3:.2f} ms", flush=True)
