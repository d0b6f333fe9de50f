"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (count, total, share)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("gpc::", "")
    a = agg[name]
    a[0] += 1
    a[1] += v
    a[2] = max(a[2], v)
    tot += v
print("%-58s %6s %11s %10s %10s %7s" % ("kernel", "n", "total_ms", "avg_us", "max_us", "share"))
for k, (c, t, m) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-58s %6d %11.3f %10.2f %10.2f %6.1f%%" % (k[:58], c, t / 1e3, t / c, m, t / tot * 100))
print("%-58s %6d %11.3f" % ("TOTAL (serialised, cold-cache)", sum(a[0] for a in agg.values()), tot / 1e3))
