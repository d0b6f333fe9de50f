import faulthandler, sys, time, os
faulthandler.dump_traceback_later(40, exit=True)
t0=time.time()
import numpy as np
print("numpy", time.time()-t0, flush=True)
what=sys.argv[1]
if what=="scipy":
    import scipy.linalg as sla
    print("scipy import", time.time()-t0, flush=True)
    T=np.tril(np.random.rand(16,16))+np.eye(16); B=np.random.rand(16,30)
    print(sla.solve_triangular(T,B,lower=True).sum(), time.time()-t0, flush=True)
elif what=="ref":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import refbind as R
    print("threads", R.get_threads(), flush=True)
    X=np.random.rand(100,3); y=np.random.rand(100,1)
    print(R.kern_compute(['rbf','white'],[0,0,-2],X).sum(), time.time()-t0, flush=True)
    print(R.gp_eval(['rbf','white'],[0,0,-2],X,y)['ll'], time.time()-t0, flush=True)
elif what=="refafter":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import gpc_b200 as G
    X=np.random.rand(300,3); y=np.random.rand(300,1)
    gp=G.CGp(G.make_kern(['rbf','white'],3,[0,0,-2]),X,y); print(gp.logLikelihood(), time.time()-t0, flush=True)
    from oracle import refbind as R
    print(R.gp_eval(['rbf','white'],[0,0,-2],X,y)['ll'], time.time()-t0, flush=True)
