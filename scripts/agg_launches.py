"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel; print conv launches individually."""
import csv, collections, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
agg, tot, per = collections.OrderedDict(), 0.0, []
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(row['Metric Value'].replace(',', ''))
    v = v / 1e3 if row['Metric Unit'] == 'ns' else (v * 1e3 if row['Metric Unit'] == 'ms' else v)
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    agg.setdefault(name, [0, 0.0]); agg[name][0] += 1; agg[name][1] += v; tot += v
    per.append((name, v, row['Grid Size']))
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:64]:64s} n={n:4d} {v/1e3:9.3f} ms {100*v/tot:5.1f}%")
print(f"total {tot/1e3:.3f} ms")
if len(sys.argv) > 2:
    for name, v, g in per:
        if sys.argv[2] in name:
            print(g, f"{v/1e3:.3f}")
