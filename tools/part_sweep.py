"""tools/part_sweep.py -- the single-rank one-sweep evaluation (gpc_dist_*, world = 1) with the SMs partitioned into a
chain set and a bulk set (GPC_SM_PARTITION = number of chain SMs; 0 = ordinary priority streams), over block sizes.
    python tools/part_sweep.py [c2] [reps]
One process (one `import torch`), one JSON line per configuration."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import gpc_b200 as G  # noqa: E402
from gpc_b200.dist import DistGp  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
parts = [int(v) for v in os.environ.get("PARTS", "0,8,16,24,32").split(",")]
nbs = [int(v) for v in os.environ.get("NBS", "512,1024").split(",")]
w = bench.WORKLOADS[name]
X, y, params = bench.make_inputs(name)
kern = G.make_kern(w["types"], w["D"])
kern.setParams(params)
ref = None
for nb in nbs:
    for R in parts:
        os.environ["GPC_SM_PARTITION"] = str(R)
        gp = DistGp(kern, X, y, grid=(1, 1), nb=nb, backend="local", devices=[0])
        ts = []
        for r in range(reps):
            t0 = time.time()
            g, ll = gp.logLikelihoodGradient()
            ts.append(time.time() - t0)
        info = gp.info()
        gp.close()
        if ref is None:
            ref = (ll, g.copy())
        out = {"workload": name, "nb": nb, "chain_sms": R, "ms_min": 1e3 * min(ts), "ms_med": 1e3 * float(np.median(ts)),
               "sweep_ms": info["phases_ms"]["sweep"], "ll": ll, "ll_rel": abs(ll - ref[0]) / max(1.0, abs(ref[0])),
               "g_rel": float(np.max(np.abs(g - ref[1]) / np.maximum(1.0, np.abs(ref[1]))))}
        print(json.dumps(out), flush=True)
