"""Markdown table of bench.py lines (profiles/r02_bench_*.json): GPU value, e2e, CPU reference, roofline.

    python tools/bench_table.py profiles/r02_bench_*.json
"""
import json, os, sys

rows = []
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as ex:
        print(f"<!-- {f}: {ex} -->")
        continue
    rows.append((os.path.basename(f), d))
print("| line | config | o_v/o_n | GPUs | views/GPU/step | views/s (device) | ms/step | input edges/s | e2e views/s (full rows) | e2e views/s (unweighted CSC) | single call ms | k_eliminate ms | roofline frac (k_eliminate) | path frac | CPU reference views/s (cores) | CPU single process views/s |")
print("|" + "---|" * 16)
for name, d in rows:
    c = d["config"]; r = d["roofline"]; cb = d.get("cpu_baseline") or {}
    wl = c["workload"]
    ovon = wl.split("o_v=")[1].replace(", o_n=", "/")
    fmt = lambda x, p=0: "-" if x is None else f"{x:,.{p}f}"
    print(f"| `{name}` | {c['config']} | {ovon} | {d['n_gpus']} | {c['views_per_gpu_per_step']} | {fmt(d['value'])} | {d['ms_per_step']:.2f} | "
          f"{d['edges_per_sec']:.3g} | {fmt(d['e2e']['value'])} | {fmt((d.get('e2e_csc_unweighted') or {}).get('value'))} | "
          f"{fmt(d.get('single_call_latency_ms'), 2)} | {r['kernel_ms']:.2f} | {r['frac']:.4f} | {r['path']['frac']:.4f} | "
          f"{fmt(cb.get('value'), 2)} ({cb.get('cores', '-')}) | {fmt(cb.get('single_process_value'), 2)} |")
