"""per-kernel table from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list

    python tools/ncu_launch_table.py launches.csv [first_id last_id]      # only the launches with first_id <= ID <= last_id
"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
lo_id = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi_id = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]; ik = h.index('Kernel Name'); im = h.index('Metric Name'); iv = h.index('Metric Value'); iid = h.index('ID')
d = collections.defaultdict(dict)
for r in rows[hi + 1:]:
    if len(r) <= iv: continue
    if not (lo_id <= int(r[iid]) <= hi_id): continue
    d[(int(r[iid]), r[ik])][r[im]] = float(r[iv].replace(',', ''))
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for (i, k), m in d.items():
    name = k.split('(')[0].replace('void ', '').replace('rlap::', '')
    a = agg[name]; a[0] += 1; a[1] += m.get('gpu__time_duration.sum', 0) / 1e3
    a[2] += m.get('dram__bytes_read.sum', 0) / 1e6; a[3] += m.get('dram__bytes_write.sum', 0) / 1e6
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total us | share | avg us | DRAM read MB/launch | DRAM write MB/launch | DRAM GB/s |")
print("|---|---|---|---|---|---|---|---|")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    gbs = (a[2] + a[3]) / 1e3 / (a[1] / 1e6) if a[1] else 0
    print(f"| `{name}` | {a[0]} | {a[1]:.0f} | {100*a[1]/tot:.1f}% | {a[1]/a[0]:.1f} | {a[2]/a[0]:.1f} | {a[3]/a[0]:.1f} | {gbs:.0f} |")
