"""tests/golden/make_golden_c3.py -- TEST INFRASTRUCTURE.  ONE run of the compiled, unmodified reference (oracle/_ref,
oracle/build_ref.sh) on BASELINE.json configs[2] "C3" at its real size: N=32768, D=16, cmpnd(rbfard, white), the
inputs of bench.make_inputs("c3").  BASELINE.md section 3: "run the reference once for parity (ll, 19
gradients)".  Needs ~45 GB of host memory; writes tests/golden/c3_reference.json.

    python tests/golden/make_golden_c3.py 1

RUN IT WITH ONE BLAS THREAD.  The multi-threaded dpotrf_ of the OpenBLAS this container has (scipy's wheel, 0.3.31.dev)
is broken at n = 32768: on 0.5*ones + I (clearly positive definite) it returns info = 16545 with 7-8 threads and info = 0
with one (checked with scipy.linalg.lapack.dpotrf directly, 24 s vs 167 s).  Through the reference that surfaces as
"Matrix non positive definite error" out of CMatrix::jitChol (CMatrix.cpp:767-804) after the jitter schedule gives up.
Single-threaded the evaluation takes 22 minutes (logLikelihood 828 s, logLikelihoodGradient 510 s).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (make_inputs only: numpy, no device code)
from oracle import gp_oracle as O  # noqa: E402
from oracle import refbind as R  # noqa: E402


def main():
    threads = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)
    w = bench.WORKLOADS["c3"]
    X, y, params = bench.make_inputs("c3")
    D = w["D"]
    kern = [("rbfard", params[:2 + D]), ("white", params[2 + D:])]
    tp0 = O.trans_from_kern(kern, D)
    R.set_threads(threads)
    t0 = time.time()
    r = R.gp_eval(w["types"], tp0, X, y)
    out = {"workload": w["desc"], "N": w["N"], "D": D, "types": w["types"], "tparams": list(map(float, tp0)),
           "params": list(map(float, params)), "ll": float(r["ll"]), "g": list(map(float, r["g"])),
           "x_checksum": float(np.sum(X * np.arange(1, D + 1))), "y_checksum": float(np.sum(y * y)),
           "t_ll_s": float(r["t_ll"]), "t_grad_s": float(r["t_grad"]), "t_eval_s": float(r["t_eval"]),
           "threads": threads, "wall_s": time.time() - t0,
           "how": "oracle/_ref/libgpcref.so (GPc -O3 gnu++98 + scipy OpenBLAS), CGp::logLikelihood + logLikelihoodGradient"}
    with open(os.path.join(ROOT, "tests", "golden", "c3_reference.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
