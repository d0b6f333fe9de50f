import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
import gpc_b200 as G
from gpc_b200.sparse import SparseGp
from oracle import gp_oracle as O, gp_sparse_oracle as S
def rel(a,b): return np.max(np.abs(a-b)/np.maximum(1,np.maximum(np.abs(a),np.abs(b))))
rng = np.random.default_rng(17)
N, M, D, d = 700, 150, 3, 3
X = rng.standard_normal((N, D))
Xu = X[rng.choice(N, M, replace=False)] + 0.05 * rng.standard_normal((M, D))
y = np.sin(X[:, :1]) @ np.ones((1, d)) + 0.1 * rng.standard_normal((N, d))
for types,tp in [(["rbf","white"],[-0.3,0.2,-3.0]),(["matern32","white"],[0.4,-0.5,-3.0]),(["lin","white"],[-1.5,-3.0]),(["rbf", "matern32", "lin", "bias", "white"],[-0.3, 0.2, 0.4, -0.5, -1.5, -2.0, -3.0])]:
    tp=np.array(tp)
    for Mx in (64, 128, 150):
        gp = SparseGp(G.make_kern(types, D, tp), X, y, Xu[:Mx], 25.0, "dtc", bias=y.mean(0))
        g, ll = gp.logLikelihoodGradient()
        r = S.sparse_loglik_grad(O.kern_from_trans(types, tp, D), X, y, Xu[:Mx], 25.0, "dtc", bias=y.mean(0))
        print(types, Mx, "ll", abs(ll-r["ll"]), "gXu", rel(g[:Mx*D], r["g"][:Mx*D]), "gk", rel(g[Mx*D:], r["g"][Mx*D:]), "jit", gp._out[4:], flush=True)
        gp.close()
# large
rng = np.random.default_rng(3)
N, M, D = 100000, 1024, 4
X = rng.standard_normal((N, D))
y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((N, 1))
Xu = X[rng.choice(N, M, replace=False)].copy()
types, tp = ["rbf", "white"], np.array([np.log(0.5), 0.0, np.log(0.1)])
for approx in ("dtc", "fitc"):
    gp = SparseGp(G.make_kern(types, D, tp), X, y, Xu, 10.0, approx, bias=y.mean(0))
    t0=time.time(); g, ll = gp.logLikelihoodGradient(); t1=time.time(); g, ll = gp.logLikelihoodGradient(); t2=time.time()
    print(approx, "ll", ll, "out", gp._out, "time", t1-t0, t2-t1, flush=True)
    t0=time.time()
    r = S.sparse_loglik_grad(O.kern_from_trans(types, tp, D), X, y, Xu, 10.0, approx, bias=y.mean(0))
    print(" oracle ll", r["ll"], "t", time.time()-t0, "ll rel", abs(ll-r["ll"])/abs(r["ll"]), "gXu", rel(g[:M*D], r["g"][:M*D]), "gk", rel(g[M*D:], r["g"][M*D:]), g[M*D:], r["g"][M*D:], flush=True)
    gp.close()
