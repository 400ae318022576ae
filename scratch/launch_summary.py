import csv, sys, re, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
agg = collections.OrderedDict()
tot = 0.0
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum": continue
    name = row["Kernel Name"]
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"<unnamed>::", "", name)
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    if unit == "ns": v /= 1e3
    elif unit == "ms": v *= 1e3
    elif unit == "s": v *= 1e6
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += v; tot += v
half = len(sys.argv) > 2
print(f"{'kernel':70s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {n:8d} {t:12.1f} {t/n:10.2f} {100*t/tot:6.1f}%")
print("total_us", tot)
