"""Summarise an ncu --csv launch list (gpu__time_duration.sum) per kernel name."""
import csv, sys, re, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = collections.OrderedDict()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1; t[1] += v * scale
total = sum(t[1] for t in tot.values())
print(f"{'kernel':40s} {'launches':>8s} {'total us':>12s} {'avg us':>10s} {'share':>7s}")
for k, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:40s} {c:8d} {us:12.1f} {us / c:10.1f} {100 * us / total:6.1f}%")
print(f"{'TOTAL':40s} {'':8s} {total:12.1f}")
