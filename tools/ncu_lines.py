"""Aggregate ncu warp-stall samples per CUDA source line:  python tools/ncu_lines.py <report.ncu-rep> [min_pct]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
agg = collections.OrderedDict()
src = {}
stall_cols = {}
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        ci = hdr.index("# Samples")
        stall_cols = {i: h for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
        continue
    if hdr is None or len(r) <= ci:
        continue
    try:
        ln = int(r[0])
        s = int(r[ci])
    except ValueError:
        continue
    a = agg.setdefault(ln, [0, collections.Counter()])
    a[0] += s
    src[ln] = r[1]
    for i, h in stall_cols.items():
        try:
            a[1][h] += int(r[i])
        except ValueError:
            pass
tot = sum(a[0] for a in agg.values())
print("total samples", tot)
for ln in sorted(agg):
    s, st = agg[ln]
    if s >= tot * minpct / 100:
        top = ", ".join("%s %d" % (k.replace("stall_", ""), v) for k, v in st.most_common(3))
        print("%6d %5.1f%%  L%-4d %-90s | %s" % (s, 100.0 * s / tot, ln, src[ln].strip()[:90], top))
