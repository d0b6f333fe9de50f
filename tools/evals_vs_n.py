"""tools/evals_vs_n.py -- BASELINE metric M1: GP logLik+grad evaluations per second (fp64) vs N on one B200, with the
unmodified reference (oracle/_ref, all host cores) beside it where it finishes in seconds.  rbf(1/D, 1) + white(0.01),
D = 8, the C2 recipe at every N."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpc_b200 as G  # noqa: E402

try:
    from oracle import refbind as R
    have_ref = R.available()
except Exception:
    have_ref = False

D = 8
rows = []
for N in (1024, 2048, 4096, 8192, 16384, 32768):
    rng = np.random.default_rng(20261017)
    X = rng.standard_normal((N, D))
    y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((N, 1))
    y -= y.mean()
    kern = G.make_kern(["rbf", "white"], D)
    kern.setParams([1.0 / D, 1.0, 0.01])
    gp = G.CGp(kern, X, y)
    ts = []
    for rep in range(6 if N <= 16384 else 3):
        gp.KupToDate = False
        t0 = time.time()
        g, ll = gp.logLikelihoodGradient()
        ts.append(time.time() - t0)
    t = min(ts[1:])
    row = {"N": N, "ms_per_eval": t * 1e3, "evals_per_s": 1.0 / t, "tflops_equiv": N ** 3 / t / 1e12, "ll": ll,
           "phases_ms": {k: round(float(v), 3) for k, v in gp.timings().items()}}
    gp.ctx.close()
    if have_ref and N <= 4096:
        R.set_threads(os.cpu_count() or 1)
        r = R.gp_eval(["rbf", "white"], kern.getTransParams(), X, y)
        row["reference_ms_per_eval"] = r["t_eval"] * 1e3
        row["reference_cores"] = os.cpu_count()
        row["ll_rel_diff_vs_reference"] = abs(ll - r["ll"]) / max(1.0, abs(r["ll"]))
        row["grad_rel_diff_vs_reference"] = float(np.max(np.abs(g - r["g"]) / np.maximum(1.0, np.abs(r["g"]))))
    rows.append(row)
    print(json.dumps(row), flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "evals_vs_n.json"), "w"), indent=1)
