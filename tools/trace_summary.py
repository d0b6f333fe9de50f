"""tools/trace_summary.py TAG [--timeline] -- summary of a kernel timeline captured by tools/trace_eval.py
(gpurun_out/trace_TAG.json.gz): time per kernel, how long ANY kernel / a LARGE kernel (grid >= 100 CTAs) / the
tensor-core GEMM was running (union over streams), what is left for the serial chain of small kernels, and -- with
--timeline -- the sequence of large kernels with the chain segments between them."""
import collections
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rows = json.load(gzip.open(os.path.join(ROOT, "gpurun_out", "trace_%s.json.gz" % tag)))["rows"]


def gridsz(g):
    p = 1
    for x in (g or [0]):
        p *= x
    return p


def union(iv):
    iv = sorted(iv)
    tot, cs, ce = 0.0, None, None
    for a, b in iv:
        if cs is None:
            cs, ce = a, b
        elif a <= ce:
            ce = max(ce, b)
        else:
            tot += ce - cs
            cs, ce = a, b
    if cs is not None:
        tot += ce - cs
    return tot


def short(n):
    return n.replace("void ", "").replace("gpc::", "").split("(")[0][:44]


span = rows[-1][2] + rows[-1][3]
print("%s: %d kernels / copies, span %.1f us" % (tag, len(rows), span))
agg = collections.defaultdict(lambda: [0, 0.0])
for n, s, t, du, g, b in rows:
    a = agg[short(n)]
    a[0] += 1
    a[1] += du
print("%-46s %5s %10s %7s" % ("kernel", "n", "total us", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print("%-46s %5d %10.1f %6.1f%%" % (k, v[0], v[1], 100.0 * v[1] / span))
allk = [(t, t + du) for n, s, t, du, g, b in rows]
big = [(t, t + du) for n, s, t, du, g, b in rows if gridsz(g) >= 100]
oz = [(t, t + du) for n, s, t, du, g, b in rows if "oz_gemm" in n]
leaf = sum(du for n, s, t, du, g, b in rows if "potrf_leaf" in n)
ub, ua, uo = union(big), union(allk), union(oz)
print("any kernel running            %9.1f us  (%.1f%% of the span)" % (ua, 100 * ua / span))
print("a large kernel (>= 100 CTAs)  %9.1f us  (%.1f%%)" % (ub, 100 * ub / span))
print("oz_gemm_kernel running        %9.1f us  (%.1f%%)" % (uo, 100 * uo / span))
print("only small kernels / idle     %9.1f us  (%.1f%%): the serial chain (diagonal-block kernels %.1f us)" %
      (span - ub, 100 * (span - ub) / span, leaf))
print("streams used: %d" % len(set(r[1] for r in rows)))
if "--timeline" in sys.argv:
    seg = None
    for n, s, t, du, g, b in rows:
        isbig = (gridsz(g) >= 100 or du > 60) and "leaf" not in n
        if isbig:
            if seg:
                print("%9.1f  chain x%-3d span %7.1f (kernel time %7.1f)" % (seg[1], seg[0], seg[2] - seg[1], seg[3]))
                seg = None
            print("%9.1f  %-22s stream %-3d %7.1f us  grid %d" % (t, short(n)[:22], s, du, gridsz(g)))
        else:
            if seg is None:
                seg = [0, t, t + du, 0.0]
            seg[0] += 1
            seg[2] = max(seg[2], t + du)
            seg[3] += du
