"""tools/posterior_bench.py -- prediction throughput (CGp::posteriorMeanVar, reference CGp.cpp:535-663): test points/s."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpc_b200 as G
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
Ns = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
w = bench.WORKLOADS[name]
X, y, params = bench.make_inputs(name)
kern = G.make_kern(w["types"], w["D"]); kern.setParams(params)
gp = G.CGp(kern, X, y)
g, ll = gp.logLikelihoodGradient()
rng = np.random.default_rng(5)
Xs = rng.standard_normal((Ns, w["D"]))
for rep in range(3):
    t0 = time.time(); mu, var = gp.posteriorMeanVar(Xs); dt = time.time() - t0
    print("%s N=%d: %d test points in %.1f ms = %.0f points/s  (mu[0]=%.10f var[0]=%.10f min var %.3e)" % (
        name, w["N"], Ns, dt * 1e3, Ns / dt, mu[0, 0], var[0, 0], var.min()), flush=True)
