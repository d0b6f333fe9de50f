"""tools/shim_c5.py [iters] -- config 5 with the UNMODIFIED reference gplvm front-end: OpenBLAS build vs the same objects
linked against the gpc_b200 Fortran shim (level 0 of INTEGRATION.md).  The kernel-matrix loops stay on the host, only
dpotrf_/dpotri_/dtrsm_/dsyrk_/dgemm_ move to the GPU: this is the compatibility path, not gpc_eval."""
import os, re, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
Y = np.load(os.path.join(ROOT, "tests/golden/oil_train.npz"))["Y"]
tmp = tempfile.mkdtemp()
data = os.path.join(tmp, "oil.svml")
with open(data, "w") as f:
    for i in range(Y.shape[0]):
        f.write("0 " + " ".join("%d:%.17g" % (j + 1, Y[i, j]) for j in range(Y.shape[1])) + "\n")
for name in ("gplvm", "gplvm_b200"):
    exe = os.path.join(ROOT, "oracle", "_ref", name)
    t0 = time.time()
    out = subprocess.run([exe, "-v", "3", "learn", "-#", str(iters), data, os.path.join(tmp, name + ".model")], cwd=tmp,
                         capture_output=True, text=True)
    dt = time.time() - t0
    errs = re.findall(r"Iteration:\s*(\d+)\s*Error:\s*([-+0-9.eE]+)", out.stdout)
    print("%s: rc=%d %.2f s for %d iterations; last objective %s" % (name, out.returncode, dt, iters, errs[-1] if errs else out.stdout[-300:]))
