"""GPU tests of the multi-GPU path (gpc_b200/dist.py + the gpc_dev_* device-level ABI): the single-rank run exercises
every device primitive; the 2-rank NCCL run is skipped when the box has a single GPU."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpc_b200 as G  # noqa: E402
from conftest import rel_err  # noqa: E402
from oracle import gp_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu


def _problem(N, D, d, seed=3):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, D))
    y = np.sin(X[:, :1]) @ np.ones((1, d)) + 0.1 * rng.standard_normal((N, d))
    types = ["rbf", "matern52", "lin", "white"]
    tp = np.array([-1.0, 0.1, 0.8, -0.4, -2.0, -2.2])
    return X, y, types, tp


@pytest.mark.parametrize("N,NB", [(300, 128), (1000, 256), (1500, 512)])
def test_dist_single_rank_vs_oracle(N, NB):
    from gpc_b200.dist import DeviceOps, DistGp
    X, y, types, tp = _problem(N, 5, 2)
    ops = DeviceOps(0)
    gp = DistGp(ops, G.make_kern(types, 5, tp), X, y, NB=NB)
    g, ll = gp.logLikelihoodGradient()
    r = O.gp_loglik_grad(O.kern_from_trans(types, tp, 5), X, y)
    assert rel_err(ll, r["ll"]) < 1e-8
    assert rel_err(g, r["g"]) < 1e-8
    # and against the single-context path
    g1, ll1 = G.CGp(G.make_kern(types, 5, tp), X, y).logLikelihoodGradient()
    assert rel_err(ll, ll1) < 1e-10 and rel_err(g, g1) < 1e-9
    ops.close()


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from gpc_b200.dist import DeviceOps, DistGp
    X, y, types, tp = _problem(2000, 5, 2)
    gp = DistGp(DeviceOps(rank), G.make_kern(types, 5, tp), X, y, NB=256)
    g, ll = gp.logLikelihoodGradient()
    if rank == 0:
        np.savez(out, g=g, ll=ll)
    dist.destroy_process_group()


def test_dist_two_ranks_nccl(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, 29600 + os.getpid() % 1000, out), nprocs=2, join=True)
    r = np.load(out)
    X, y, types, tp = _problem(2000, 5, 2)
    ref = O.gp_loglik_grad(O.kern_from_trans(types, tp, 5), X, y)
    assert rel_err(float(r["ll"]), ref["ll"]) < 1e-8
    assert rel_err(r["g"], ref["g"]) < 1e-8
