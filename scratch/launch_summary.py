"""Aggregates an ncu gpu__time_duration launch list (csv) per kernel: python scratch/launch_summary.py file.csv [first last]"""
import csv, re, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
if len(sys.argv) > 3:
    rows = rows[int(sys.argv[2]):int(sys.argv[3])]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r'\(.*$', '', r[4]).replace('void sd::', '').replace('void ', '')
    name = re.sub(r'at::native::.*', 'torch elementwise / reduce', name)
    agg[name][0] += 1
    agg[name][1] += float(r[-1]) / 1e6
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[1]}: {len(rows)} launches, {tot:.1f} ms (ncu per-launch times: cold cache, serialised -- compare SHARES)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if v[1] / tot < 0.0005: continue
    print(f"{v[1]:9.2f} ms {100 * v[1] / tot:5.1f}%  n={v[0]:4d}  avg {v[1] / v[0] * 1000:8.1f} us  {k[:110]}")
