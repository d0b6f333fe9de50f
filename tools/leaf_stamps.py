"""tools/leaf_stamps.py -- the diagonal-block kernel in isolation: time per launch and clock64 phase stamps."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpc_b200 as G  # noqa: E402
from gpc_b200._lib import check, lib, ptr  # noqa: E402

us = C.c_double(0)
st = np.zeros(32, dtype=np.int64)
check(lib().gpc_bench_leaf(0, 50, C.byref(us), ptr(st)))
print("leaf: %.2f us per launch" % us.value)
names = {0: "start", 1: "loaded", 2: "diag0 factored (warp 0)", 24: "factor stored", 25: "inv level 16", 26: "inv level 32",
         27: "inv level 64", 28: "stores issued"}
prev = 0
for i, v in enumerate(st):
    if i and v == 0:
        continue
    nm = names.get(i)
    if nm is None:
        p, k = divmod(i - 3, 3)
        nm = "panel %d %s" % (p, ["solve done", "warp0 next diag done", "joined"][k])
    print("%2d %-28s %8d clk  (+%d)" % (i, nm, v, v - prev))
    prev = v
