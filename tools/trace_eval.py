"""tools/trace_eval.py -- kernel timeline of one evaluation (CUPTI activity records through torch.profiler; there is no
nsys in the image).  Writes gpurun_out/trace_<tag>.json.gz: [name, stream, start_us, dur_us, grid, block] per kernel.
    python tools/trace_eval.py eval|sweep [c2] [nb] [tag]
Environment knobs of the library (GPC_SM_PARTITION, GPC_POTRF_MODE, ...) are read as usual."""
import gzip
import json
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import gpc_b200 as G  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "eval"
name = sys.argv[2] if len(sys.argv) > 2 else "c2"
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 512
tag = sys.argv[4] if len(sys.argv) > 4 else "%s_%s" % (mode, name)
w = bench.WORKLOADS[name]
X, y, params = bench.make_inputs(name)
kern = G.make_kern(w["types"], w["D"])
kern.setParams(params)
if mode == "sweep":
    from gpc_b200.dist import DistGp
    gp = DistGp(kern, X, y, grid=(1, 1), nb=nb, backend="local", devices=[0])

    def run():
        return gp.logLikelihoodGradient()
else:
    gp = G.CGp(kern, X, y)

    def run():
        gp.KupToDate = False
        return gp.logLikelihoodGradient()

for _ in range(3):
    run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g, ll = run()
    torch.cuda.synchronize()
path = os.path.join(ROOT, "gpurun_out", "trace_%s_raw.json" % tag)
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
rows = []
for e in ev:
    if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e:
        a = e.get("args", {})
        rows.append([e["name"][:60], a.get("stream", -1), e["ts"], e["dur"], a.get("grid", None), a.get("block", None)])
rows.sort(key=lambda r: r[2])
t0 = rows[0][2] if rows else 0
for r in rows:
    r[2] = round(r[2] - t0, 3)
with gzip.open(os.path.join(ROOT, "gpurun_out", "trace_%s.json.gz" % tag), "wt") as f:
    json.dump({"ll": ll, "rows": rows}, f)
os.remove(path)
span = (rows[-1][2] + rows[-1][3]) if rows else 0
print(json.dumps({"tag": tag, "kernels": len(rows), "span_us": span, "ll": ll}))
