import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np, torch
import rlap_b200
from rlap_b200 import graphs, ops
n=169343; V=64
ei = graphs.barabasi_albert(n,7,seed=0)
dev=torch.device('cuda',0)
ei_pinned = torch.from_numpy(ei).pin_memory()
copy_stream = torch.cuda.Stream(device=dev)
hb=[None,None]; pend=[None,None]
T=time.perf_counter
for step in range(8):
    t0=T()
    slot=step&1
    if pend[slot] is not None: pend[slot].synchronize()
    t1=T()
    d = ei_pinned.to(dev, non_blocking=True)
    g = ops.prepare(d, None, n)
    t2=T()
    (row,col,w),vp = ops.schur_views(g, n//2, "degree","asc", num_views=V, seed=step, dtype=None)
    t3=T()
    total=int(vp[-1])
    if hb[slot] is None:
        cap=int(total*1.05); hb[slot]=[torch.empty(cap,dtype=torch.int32).pin_memory(), torch.empty(cap,dtype=torch.int32).pin_memory(), torch.empty(cap,dtype=torch.float32).pin_memory()]
    t4=T()
    ready=torch.cuda.Event(); ready.record()
    with torch.cuda.stream(copy_stream):
        copy_stream.wait_event(ready)
        for h,x in zip(hb[slot],(row,col,w)):
            h[:total].copy_(x, non_blocking=True); x.record_stream(copy_stream)
        done=torch.cuda.Event(); done.record()
    pend[slot]=done
    t5=T()
    print(f"step {step}: wait {1e3*(t1-t0):.1f} prepare {1e3*(t2-t1):.1f} views {1e3*(t3-t2):.1f} alloc {1e3*(t4-t3):.1f} enqueue {1e3*(t5-t4):.1f} ms", flush=True)
torch.cuda.synchronize()
