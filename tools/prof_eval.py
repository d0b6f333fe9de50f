"""tools/prof_eval.py -- one C2 evaluation between cudaProfilerStart/Stop, for
   ncu --profile-from-start off --metrics gpu__time_duration.sum --csv ... python tools/prof_eval.py [N D]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402  (only for cudaProfilerStart/Stop)
import gpc_b200 as G  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
D = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rng = np.random.default_rng(20261017)
X = rng.standard_normal((N, D))
y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((N, 1))
y -= y.mean()
kern = G.make_kern(["rbf", "white"], D)
kern.setParams([1.0 / D, 1.0, 0.01])
gp = G.CGp(kern, X, y)
for _ in range(2):
    gp.KupToDate = False
    gp.logLikelihoodGradient()
torch.cuda.cudart().cudaProfilerStart()
gp.KupToDate = False
t0 = time.time()
g, ll = gp.logLikelihoodGradient()
dt = time.time() - t0
torch.cuda.cudart().cudaProfilerStop()
print("N=%d eval %.1f ms ll=%.6f launches/eval=%d phases=%s" % (N, dt * 1e3, ll, gp.ctx.launch_count() // 3, gp.timings()))
