"""BASELINE config 5: GP-LVM on examples/oilTrain.svml (N=1000, d=12, q=2, rbf+bias+white, PCA init, SCG, 100
iterations) end-to-end on the device, against the reference trajectory (gplvm -v 3 learn -# 100)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpc_b200 as G  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
Y = np.load(os.path.join(ROOT, "tests/golden/oil_train.npz"))["Y"]
kern = G.make_kern(["rbf", "bias", "white"], 2, [0.0, 0.0, -2.0, -2.0])
t0 = time.time()
lvm = G.CGplvm.fromData(kern, Y, 2)
t_setup = time.time() - t0
log = []
lvm.optimise(iters, log=log)
dt = time.time() - t0
ref = os.path.join(ROOT, "tests/golden/gplvm_c5_trajectory.json")
out = {"iters": len(log), "seconds": dt, "setup_seconds": t_setup, "evals": getattr(lvm, "nevals", None), "obj_1": log[0], "obj_50": log[min(49, len(log) - 1)], "obj_last": log[-1],
       "kernel": kern.params.tolist(), "launches": lvm.ctx.launch_count()}
if os.path.exists(ref):
    r = json.load(open(ref))
    n = min(len(log), len(r["objective"]))
    out["max_rel_diff_vs_reference"] = max(abs(a - b) / max(1.0, abs(b)) for a, b in zip(log[:n], r["objective"][:n]))
    out["reference_seconds"] = r["seconds"]
    out["reference_final"] = r["objective"][n - 1]
print(json.dumps(out))
