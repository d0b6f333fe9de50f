"""tools/prof_kern.py -- two evaluations of one configuration without PyTorch in the process, for ncu captures of the
HBM-side kernels:
   ncu --set full --clock-control none --import-source on -k regex:'kbuild_kernel|grad_kernel' -c 4 -o out \
       python tools/prof_kern.py c2|c3|c4"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpc_b200 as G  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
N, D, types = {"c2": (8192, 8, ["rbf", "white"]), "c3": (32768, 16, ["rbfard", "white"]),
               "c4": (65536, 32, ["matern52", "white"])}[cfg]
rng = np.random.default_rng(20261017)
X = rng.standard_normal((N, D))
y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((N, 1))
y -= y.mean()
kern = G.make_kern(types, D)
if cfg == "c2":
    kern.setParams([1.0 / D, 1.0, 0.01])
elif cfg == "c3":
    kern.setParams([1.0 / D, 1.0] + [0.25 + 0.5 * k / (D - 1) for k in range(D)] + [0.01])
else:
    kern.setParams([np.sqrt(D), 1.0, 0.01])
gp = G.CGp(kern, X, y)
for _ in range(2):
    gp.KupToDate = False
    t0 = time.time()
    g, ll = gp.logLikelihoodGradient()
    dt = time.time() - t0
print("%s N=%d D=%d eval %.1f ms ll=%.6f phases=%s" % (cfg, N, D, dt * 1e3, ll, gp.timings()))
