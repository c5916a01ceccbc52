"""Aggregates an ncu --csv launch list (gpu__time_duration.sum) by kernel name."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.DictReader(lines)
agg = collections.defaultdict(lambda: [0, 0.0])
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum": continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    if "dgemm" in name and "--bygrid" in sys.argv: name += " grid=" + row.get("Grid Size", "?")
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    if unit in ("nsecond", "ns"): v /= 1e6
    elif unit in ("usecond", "us"): v /= 1e3
    elif unit in ("msecond", "ms"): pass
    elif unit in ("second", "s"): v *= 1e3
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.2f} ms over {sum(v[0] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:10.3f} ms {100*v[1]/tot:5.1f}%  n={v[0]:5d}  avg {1e3*v[1]/v[0]:9.1f} us  {k}")
