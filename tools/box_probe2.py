import faulthandler, sys, time, os
faulthandler.dump_traceback_later(int(os.environ.get("PROBE_T","40")), exit=True)
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpc_b200 as G
seed=int(sys.argv[1]) if len(sys.argv)>1 else 0
np.random.seed(seed)
X=np.random.rand(300,3); y=np.random.rand(300,1)
gp=G.CGp(G.make_kern(['rbf','white'],3,[0,0,-2]),X,y)
t0=time.time(); print(gp.logLikelihood(), time.time()-t0, gp._out, flush=True)
