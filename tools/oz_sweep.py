"""tools/oz_sweep.py [c2|c3] -- evaluation time and parity of the full hot path for several GEMM-engine settings."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpc_b200 as G  # noqa: E402
from gpc_b200._lib import check, lib  # noqa: E402
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
w = bench.WORKLOADS[name]
X, y, params = bench.make_inputs(name)
kern = G.make_kern(w["types"], w["D"])
kern.setParams(params)
gp = G.CGp(kern, X, y)
L = lib()


def run(tag, reps):
    ts = []
    for _ in range(reps):
        gp.KupToDate = False
        t0 = time.time()
        g, ll = gp.logLikelihoodGradient()
        ts.append(time.time() - t0)
    import ctypes as _C
    enq = _C.c_double(0)
    check(L.gpc_last_enqueue_ms(gp.ctx.handle, _C.byref(enq)))
    print("   times ms:", " ".join("%.1f" % (t * 1e3) for t in ts), " host enqueue ms: %.2f" % enq.value, flush=True)
    return min(ts[1:] or ts), g, ll, gp.timings()


check(L.gpc_set_gemm_engine(0, 0, 0, 0))
reps = 6 if name == "c2" else 3
t0, g0, ll0, ph0 = run("dmma", reps)
print("%s DMMA only: %.2f ms  ll=%.10f  potrf %.1f inverse %.1f" % (name, t0 * 1e3, ll0, ph0["potrf"], ph0["inverse"]), flush=True)
if len(sys.argv) > 2:
    settings = [tuple(int(v) for v in a.split(",")) for a in sys.argv[2:]]
elif False:
    pass
settings_default = [(8, 2048, 2048), (8, 1024, 1024), (8, 512, 512), (8, 512, 256), (8, 256, 256), (8, 256, 128), (7, 512, 512)] if name == "c2" else \
    [(8, 2048, 2048), (8, 1024, 1024), (8, 512, 512), (7, 1024, 1024)]
if len(sys.argv) <= 2:
    settings = settings_default
for S, mn, mk in settings:
    check(L.gpc_set_gemm_engine(1, S, mn, mk))
    t, g, ll, ph = run("oz", reps)
    print("%s Ozaki S=%d min_mn=%d min_k=%d: %.2f ms  potrf %.1f inverse %.1f  |dll|/|ll|=%.2e  max rel dgrad=%.2e" % (
        name, S, mn, mk, t * 1e3, ph["potrf"], ph["inverse"], abs(ll - ll0) / abs(ll0),
        float(np.max(np.abs(g - g0) / np.maximum(1.0, np.abs(g0))))), flush=True)
