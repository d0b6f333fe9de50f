"""tools/small_n.py -- per-evaluation time and phases for small N (GP-LVM sized problems)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpc_b200 as G
from gpc_b200._lib import check, lib
for N, d in ((1000, 12), (1000, 1), (2000, 1), (4096, 1)):
    rng = np.random.default_rng(1)
    X = rng.standard_normal((N, 2))
    Y = np.sin(X[:, :1]) @ np.ones((1, d)) + 0.1 * rng.standard_normal((N, d))
    kern = G.make_kern(["rbf", "bias", "white"], 2, [0.0, 0.0, -2.0, -2.0])
    gp = G.CGp(kern, X, Y)
    for mode in ("default", "dmma"):
        check(lib().gpc_set_gemm_engine(1 if mode == "default" else 0, 0, 0, 0))
        ts = []
        for rep in range(8):
            gp.KupToDate = False
            t0 = time.time()
            g, ll = gp.logLikelihoodGradient()
            ts.append(time.time() - t0)
        print("N=%d d=%d %s: %.2f ms/eval (min of %s)  phases %s launches/eval %d" % (
            N, d, mode, min(ts) * 1e3, ["%.1f" % (t * 1e3) for t in ts], {k: round(float(v), 2) for k, v in gp.timings().items()},
            0), flush=True)
    check(lib().gpc_set_gemm_engine(1, 0, 0, 0))
    gp.ctx.close()
