"""tools/gemm_profile.py [c2|c3] -- per-GEMM time of one evaluation in profiling mode (serial, CUDA events per call),
grouped by engine and shape: where the N^3 flops of potrf + inverse are actually spent."""
import collections
import csv
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("GPC_PROF_DUMP", os.path.join(ROOT, "gpurun_out", "gemm_dump.csv"))
import gpc_b200 as G  # noqa: E402
from gpc_b200._lib import check, lib  # noqa: E402
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
w = bench.WORKLOADS[name]
X, y, params = bench.make_inputs(name)
kern = G.make_kern(w["types"], w["D"])
kern.setParams(params)
gp = G.CGp(kern, X, y)
for _ in range(2):
    gp.KupToDate = False
    gp.logLikelihoodGradient()
print("normal mode phases:", {k: round(float(v), 2) for k, v in gp.timings().items()})
check(lib().gpc_ctx_set_profile(gp.ctx.handle, 1))
gp.KupToDate = False
gp.logLikelihoodGradient()
print("profile mode (serial) phases:", {k: round(float(v), 2) for k, v in gp.timings().items()})
rows = list(csv.DictReader(open(os.environ["GPC_PROF_DUMP"])))
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in rows:
    key = (r["engine"], int(r["k"]), "lower" if r["lower"] == "1" else "full", int(r["m"]), int(r["n"]))
    a = agg[key]
    a[0] += 1
    a[1] += float(r["ms"])
    a[2] += float(r["flops"])
tot = sum(a[1] for a in agg.values())
print("%-6s %6s %-5s %6s %6s %5s %10s %8s %7s" % ("engine", "k", "tri", "m", "n", "count", "ms", "TF/s", "share"))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%-6s %6d %-5s %6d %6d %5d %10.3f %8.1f %6.1f%%" % (key[0], key[1], key[2], key[3], key[4], a[0], a[1],
                                                           a[2] / a[1] / 1e9 if a[1] else 0, 100 * a[1] / tot))
by_eng = collections.defaultdict(lambda: [0, 0.0, 0.0])
for key, a in agg.items():
    b = by_eng[key[0]]
    b[0] += a[0]; b[1] += a[1]; b[2] += a[2]
for e, b in by_eng.items():
    print("engine %s: %d calls, %.2f ms, %.3e flops, %.1f TFLOP/s" % (e, b[0], b[1], b[2], b[2] / b[1] / 1e9))
print("all GEMM calls: %.2f ms" % tot)
