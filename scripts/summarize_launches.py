"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST iteration."""
import csv, sys, collections, re
path, per_iter = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        rows.append((re.sub(r"\(.*", "", r["Kernel Name"]), v, r.get("Grid Size", ""), r.get("Block Size", "")))
if per_iter:
    rows = rows[-per_iter:]
tot = sum(v for _, v, _, _ in rows)
agg = collections.OrderedDict()
for n, v, g, b in rows:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1; a[1] += v
print("launches %d  total %.1f us" % (len(rows), tot))
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-48s n=%4d  %9.1f us  %5.1f%%" % (n[:48], c, v, 100 * v / tot))
if "--list" in sys.argv:
    for n, v, g, b in rows:
        print("%-44s %9.1f us  grid %s block %s" % (n[:44], v, g, b))
