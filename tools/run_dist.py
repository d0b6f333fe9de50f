"""tools/run_dist.py -- time the sharded evaluation (gpc_b200/dist.py) under torchrun:
   python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/run_dist.py c4 [NB]
workloads: c3 (N=32768 D=16 rbfard+white), c4 (N=65536 D=32 matern52+white), or an integer N (rbf+white, D=8)."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpc_b200 as G  # noqa: E402
from gpc_b200.dist import DeviceOps, DistGp  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
NB = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rng = np.random.default_rng(20261017)
if wl == "c4":
    N, D = 65536, 32
    kern = G.make_kern(["matern52", "white"], D)
    kern.setParams([np.sqrt(D), 1.0, 0.01])
elif wl == "c3":
    N, D = 32768, 16
    kern = G.make_kern(["rbfard", "white"], D)
    kern.setParams(np.concatenate([[1.0 / D, 1.0], 0.25 + 0.5 * np.arange(D) / (D - 1), [0.01]]))
else:
    N, D = int(wl), 8
    kern = G.make_kern(["rbf", "white"], D)
    kern.setParams([1.0 / D, 1.0, 0.01])
X = rng.standard_normal((N, D))
y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((N, 1))
y -= y.mean()
ops = DeviceOps(lr)
gp = DistGp(ops, kern, X, y, NB=NB)
ts = []
for rep in range(reps):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    g, ll = gp.logLikelihoodGradient()
    torch.cuda.synchronize()
    ts.append(time.time() - t0)
if rank == 0:
    print(json.dumps({"workload": wl, "N": N, "D": D, "NB": NB, "world": world, "seconds": ts, "best": min(ts), "ll": ll,
                      "g": list(map(float, g[:4])), "tflops_equiv": N ** 3 / min(ts) / 1e12,
                      "mem_gb": torch.cuda.max_memory_allocated() / 1e9, "phases_s": gp.times}))
if world > 1:
    dist.destroy_process_group()
