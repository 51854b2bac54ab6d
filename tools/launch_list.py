"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (in order of first launch)."""
import csv, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = {}
for r in rows[1:]:
    name = re.sub(r"\(.*$", "", r[k])[:100]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[v]) / 1e6
print(f"{'launches':>8} {'total ms':>12} {'avg us':>12}  kernel")
for name, (n, ms) in agg.items():
    print(f"{n:8d} {ms:12.3f} {ms / n * 1e3:12.1f}  {name}")
