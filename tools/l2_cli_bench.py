"""tools/l2_cli_bench.py -- the reference's own `gp learn` front-end end to end (process start, SVM-light read, model
construction, SCG iterations, model file written): OpenBLAS build (oracle/_ref/gp) vs the same gp.cpp compiled on
CGpB200 (oracle/_ref/gp_l2, INTEGRATION.md level 2).
   python tools/l2_cli_bench.py [N_both=2048] [N_l2_only=8192] [iters=10]"""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
N_both = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
N_l2 = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
D = 8


def write_svml(path, N):
    rng = np.random.default_rng(20261017)
    X = rng.standard_normal((N, D))
    y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(N)
    with open(path, "w") as f:
        for i in range(N):
            f.write("%.17g %s\n" % (y[i], " ".join("%d:%.17g" % (j + 1, X[i, j]) for j in range(D))))


def learn(binary, data, model, cwd):
    t0 = time.time()
    out = subprocess.run([os.path.join(REF, binary), "-v", "2", "learn", "-#", str(iters), data, model], cwd=cwd,
                         capture_output=True, text=True, timeout=3000)
    dt = time.time() - t0
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    ll = float(re.findall(r"Log likelihood:\s*([-+0-9.eE]+)", out.stdout)[-1])
    return {"seconds": round(dt, 3), "ll": ll}


res = {"iters": iters, "D": D}
with tempfile.TemporaryDirectory() as tmp:
    for N, bins in ((N_both, ("gp_l2", "gp")), (N_l2, ("gp_l2",))):
        if N <= 0:
            continue
        data = os.path.join(tmp, "d%d.svml" % N)
        write_svml(data, N)
        for b in bins:
            res["N%d_%s" % (N, b)] = learn(b, data, os.path.join(tmp, "m_%s_%d" % (b, N)), tmp)
print(json.dumps(res))
