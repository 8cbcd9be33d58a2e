"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table."""
import collections
import csv
import sys

path, title = sys.argv[1], sys.argv[2]
lines = open(path).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(lines[start:]))
agg = collections.OrderedDict()
for r in rows:
    k, v = r["Kernel Name"], float(r["Metric Value"])
    u = r["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else v
    agg.setdefault(k, [0, 0.0])
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print(f"# {title}\n")
print("| kernel | launches | total us | us/launch | share |\n|---|---:|---:|---:|---:|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k[:72]}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / tot:.1f}% |")
print(f"\nTotal {tot:.1f} us over {len(rows)} launches (cold-cache, serialised under ncu: compare shares, not absolutes).")
