"""tools/run_c4_single.py -- C4 (N=65536, D=32, matern52+white) on ONE GPU: K, L, L^-1, K^-1 = 4 x 34.4 GB."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpc_b200 as G  # noqa: E402

N, D = 65536, 32
rng = np.random.default_rng(20261017)
X = rng.standard_normal((N, D))
y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((N, 1))
y -= y.mean()
kern = G.make_kern(["matern52", "white"], D)
kern.setParams([np.sqrt(D), 1.0, 0.01])
gp = G.CGp(kern, X, y)
for rep in range(2):
    gp.KupToDate = False
    t0 = time.time()
    g, ll = gp.logLikelihoodGradient()
    dt = time.time() - t0
    print("C4 single GPU: %.3f s  ll=%.10f  g=%s  phases=%s  (%.1f TFLOP/s-equivalent)" % (
        dt, ll, np.array2string(g, precision=6), {k: round(float(v), 1) for k, v in gp.timings().items()}, N ** 3 / dt / 1e12), flush=True)
