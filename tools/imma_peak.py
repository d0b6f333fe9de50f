import ctypes as C, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpc_b200._lib import lib, check
out = {}
for nw in (1, 2, 4):
    t = C.c_double(0)
    check(lib().gpc_bench_imma_peak(0, nw, C.byref(t)))
    out["N=%d" % (64 * nw)] = t.value
t = C.c_double(0)
check(lib().gpc_bench_imma_peak_sustained(0, 4, 3.0, C.byref(t)))
out["N=256 sustained over 3 s"] = t.value
d = C.c_double(0)
check(lib().gpc_bench_dmma_peak(0, C.byref(d)))
out["dmma_tflops"] = d.value
print(json.dumps(out))
