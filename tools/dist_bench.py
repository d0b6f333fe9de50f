"""Timing of the sharded evaluation (gpc_dist_*, local back-end) on the GPUs of this box.
    python tools/dist_bench.py --workload c3|c4|c2 --ngpu 1 --nb 1024 [--virtual 4] [--reps 3]
--virtual V: V ranks on ONE device (protocol check, not a performance number)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import gpc_b200 as G  # noqa: E402
from gpc_b200.dist import DistGp, default_grid  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3")
ap.add_argument("--ngpu", type=int, default=1)
ap.add_argument("--virtual", type=int, default=0)
ap.add_argument("--nb", type=int, default=1024)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--n", type=int, default=0, help="override N (first N rows of the workload's inputs)")
ap.add_argument("--single", action="store_true", help="also time the single-GPU path (gpc_eval)")
ap.add_argument("--backend", default="local", help="local | nccl (under torchrun: one process per GPU)")
a = ap.parse_args()
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
if a.backend == "nccl":
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("gloo")
w = bench.WORKLOADS[a.workload]
X, y, params = bench.make_inputs(a.workload)
if a.n:
    X, y = np.asfortranarray(X[:a.n]), np.asfortranarray(y[:a.n])
kern = G.make_kern(w["types"], w["D"])
kern.setParams(params)
devices = [0] * a.virtual if a.virtual else list(range(a.ngpu))
if a.backend == "nccl":
    devices = list(range(world))
    grid = default_grid(world)
    gp = DistGp(kern, X, y, grid=grid, nb=a.nb, backend="nccl", device=int(os.environ.get("LOCAL_RANK", "0")))
else:
    grid = default_grid(len(devices))
    gp = DistGp(kern, X, y, grid=grid, nb=a.nb, backend="local", devices=devices)
ts = []
for r in range(a.reps):
    if a.backend == "nccl" and world > 1:
        dist.barrier()
    t0 = time.time()
    g, ll = gp.logLikelihoodGradient()
    ts.append(time.time() - t0)
N = X.shape[0]
out = {"backend": a.backend, "workload": a.workload, "N": N, "devices": devices, "grid": grid, "nb": a.nb, "seconds": ts, "ll": ll,
       "g": list(map(float, g[:4])), "tflops_equiv": N ** 3 / min(ts) / 1e12, "info": gp.info()}
gp.close()
if a.single and rank == 0:
    gp1 = G.CGp(kern, X, y)
    t1 = []
    for r in range(a.reps):
        gp1.KupToDate = False
        t0 = time.time()
        g1, ll1 = gp1.logLikelihoodGradient()
        t1.append(time.time() - t0)
    out["single_seconds"] = t1
    out["ll_rel_vs_single"] = abs(ll - ll1) / max(1.0, abs(ll1))
    out["g_rel_vs_single"] = float(np.max(np.abs(g - g1) / np.maximum(1.0, np.abs(g1))))
if rank == 0:
    print(json.dumps(out))
if a.backend == "nccl" and world > 1:
    dist.destroy_process_group()
