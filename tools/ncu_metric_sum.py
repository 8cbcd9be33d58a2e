"""Sum per-kernel metrics of an ncu --csv log (several --metrics): kernel -> launches, total duration, DRAM read/write bytes."""
import collections, csv, sys
lines = open(sys.argv[1]).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
agg = collections.OrderedDict()
ids = collections.defaultdict(set)
for r in csv.DictReader(lines[start:]):
    k = r["Kernel Name"][:60]; m = r["Metric Name"]; v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    if m.startswith("gpu__time"): v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    if m.startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    agg.setdefault(k, collections.Counter())[m] += v
    ids[k].add(r["ID"])
tot = collections.Counter()
print("| kernel | launches | total us | DRAM read MB | DRAM write MB |\n|---|---:|---:|---:|---:|")
for k, c in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    print(f"| `{k}` | {len(ids[k])} | {c['gpu__time_duration.sum']:.1f} | {c['dram__bytes_read.sum']/1e6:.1f} | {c['dram__bytes_write.sum']/1e6:.1f} |")
    tot.update(c)
print(f"| total | {sum(len(v) for v in ids.values())} | {tot['gpu__time_duration.sum']:.1f} | {tot['dram__bytes_read.sum']/1e6:.1f} | {tot['dram__bytes_write.sum']/1e6:.1f} |")
