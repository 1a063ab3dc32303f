"""aggregate an `ncu --page source --print-source cuda,sass --csv` export by source line"""
import csv, sys, collections
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 50
rows = list(csv.reader(open(path)))
cur = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0, ""])   # (file,line) -> samples, inst, text
stall_cols = {}
stalls = collections.defaultdict(lambda: collections.Counter())
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed")
        stall_cols = {i: h for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
        continue
    if hdr is None or len(r) < len(hdr) - 2 or not r[0].isdigit(): continue
    try:
        s = int(r[iS] or 0); ins = int(r[iI] or 0)
    except ValueError:
        continue
    k = (cur, int(r[0]))
    agg[k][0] += s; agg[k][1] += ins; agg[k][2] = r[1].strip()[:110]
    for i, h in stall_cols.items():
        try: stalls[k][h] += int(r[i] or 0)
        except ValueError: pass
tot = sum(v[0] for v in agg.values()); toti = sum(v[1] for v in agg.values())
print(f"total samples {tot}  total warp instructions {toti}")
allst = collections.Counter()
for k, c in stalls.items(): allst.update(c)
print("stall reasons:", ", ".join(f"{h[6:]} {100*c/max(tot,1):.1f}%" for h, c in allst.most_common(10)))
print("--- by samples")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    top = ",".join(f"{h[6:]}:{c}" for h, c in stalls[k].most_common(2))
    print(f"{v[0]:8d} {100*v[0]/tot:5.1f}%  inst={v[1]:10d} ({100*v[1]/toti:4.1f}%) {k[0]}:{k[1]:4d} [{top}] {v[2]}")
print("--- by instructions")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print(f"inst={v[1]:10d} ({100*v[1]/toti:4.1f}%) samples {100*v[0]/tot:5.1f}%  {k[0]}:{k[1]:4d} {v[2]}")
