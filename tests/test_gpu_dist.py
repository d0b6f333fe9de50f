"""GPU tests of the multi-GPU path (gpc_dist_* C ABI, gpc_b200/csrc/dist.cu): the fused one-sweep K -> K^-1 on a 2-D
block-cyclic layout.  The "local" back-end with one device listed several times runs the complete multi-rank protocol
(ownership, panel production, slot broadcasts, look-ahead, all-reduces) on a single GPU; the NCCL runs need >= 2 GPUs
and are skipped otherwise."""
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


@pytest.mark.parametrize("N,nb,grid", [(300, 128, (1, 1)), (1000, 256, (1, 1)), (700, 128, (1, 2)), (1000, 128, (2, 2)),
                                       (1500, 256, (2, 4)), (1100, 128, (2, 4)), (900, 128, (4, 2)), (640, 128, (3, 1)),
                                       (300, 128, (4, 2))])   # the last one: more process rows than block rows
def test_one_sweep_virtual_ranks_vs_oracle(N, nb, grid):
    """all ranks of a P x Q grid as worker threads on device 0: ll, gradient and K^-1 itself against the oracle"""
    from gpc_b200.dist import DistGp
    X, y, types, tp = _problem(N, 5, 2)
    P, Q = grid
    gp = DistGp(G.make_kern(types, 5, tp), X, y, grid=grid, nb=nb, backend="local", devices=[0] * (P * Q))
    g, ll = gp.logLikelihoodGradient()
    kern = O.kern_from_trans(types, tp, 5)
    r = O.gp_loglik_grad(kern, X, y)
    assert rel_err(ll, r["ll"]) < 1e-8
    assert rel_err(g, r["g"]) < 1e-8
    Kinv = gp.download_kinv()
    Kref = np.linalg.inv(O.kern_compute(kern, X))
    assert np.max(np.abs(Kinv - Kref)) < 1e-9 * max(1.0, np.max(np.abs(Kref)))
    # and against the single-context path
    g1, ll1 = G.CGp(G.make_kern(types, 5, tp), X, y).logLikelihoodGradient()
    assert rel_err(ll, ll1) < 1e-10 and rel_err(g, g1) < 1e-9
    info = gp.info()
    assert info["ranks"] == P * Q and info["steps"] == (N + nb - 1) // nb
    # a second evaluation on the same context (events, slots and scalars are reused) gives the same answer
    g2, ll2 = gp.logLikelihoodGradient()
    assert rel_err(ll2, ll) < 1e-12 and rel_err(g2, g) < 1e-10
    gp.close()


def test_one_sweep_jitter_schedule_matches_single_gpu():
    """a kernel matrix that needs jitter: the sharded path follows the same jitChol schedule (CMatrix.cpp:767-804)"""
    from gpc_b200.dist import DistGp
    rng = np.random.default_rng(5)
    X = rng.standard_normal((400, 2))
    X[200:] = X[:200]  # duplicated inputs and no noise term: singular to working precision
    y = rng.standard_normal((400, 1))
    types, tp = ["rbf", "bias"], np.array([0.0, 0.0, -1.0])
    gp1 = G.CGp(G.make_kern(types, 2, tp), X, y)
    g1, ll1 = gp1.logLikelihoodGradient()
    gp = DistGp(G.make_kern(types, 2, tp), X, y, grid=(2, 2), nb=128, backend="local", devices=[0] * 4)
    g, ll = gp.logLikelihoodGradient()
    assert gp.jitter > 0.0
    assert gp.jitter == pytest.approx(float(gp1._out[2]), rel=1e-12)
    assert rel_err(ll, ll1) < 1e-5 and rel_err(g, g1) < 1e-4   # conditioning ~1e10: as test_gp_jitter_retry_matches_oracle
    gp.close()


def test_one_sweep_medium_size_properties():
    """N = 8192 in 1024-blocks on a virtual 2 x 2 grid: agreement with the single-GPU evaluation"""
    from gpc_b200.dist import DistGp
    rng = np.random.default_rng(11)
    N, D = 8192, 8
    X = rng.standard_normal((N, D))
    y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((N, 1))
    types, tp = ["matern52", "white"], np.array([np.log(np.sqrt(D)), 0.0, np.log(0.01)])
    g1, ll1 = G.CGp(G.make_kern(types, D, tp), X, y).logLikelihoodGradient()
    gp = DistGp(G.make_kern(types, D, tp), X, y, grid=(2, 2), nb=1024, backend="local", devices=[0] * 4)
    g, ll = gp.logLikelihoodGradient()
    assert rel_err(ll, ll1) < 1e-9 and rel_err(g, g1) < 1e-8
    gp.close()


def _worker(rank, world, port, out, N, nb):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)   # only carries the 128-byte NCCL id
    from gpc_b200.dist import DistGp
    X, y, types, tp = _problem(N, 5, 2)
    gp = DistGp(G.make_kern(types, 5, tp), X, y, nb=nb, backend="nccl", device=rank)
    g, ll = gp.logLikelihoodGradient()
    Kinv = gp.download_kinv()
    np.savez(out + ".%d.npz" % rank, g=g, ll=ll, Kinv=Kinv)
    gp.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_one_sweep_nccl_ranks(tmp_path, world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    out = str(tmp_path / "res")
    N, nb = 2000, 256
    mp.spawn(_worker, args=(world, 29600 + os.getpid() % 1000, out, N, nb), nprocs=world, join=True)
    X, y, types, tp = _problem(N, 5, 2)
    kern = O.kern_from_trans(types, tp, 5)
    ref = O.gp_loglik_grad(kern, X, y)
    Kinv = np.zeros((N, N))
    for r in range(world):
        z = np.load(out + ".%d.npz" % r)
        assert rel_err(float(z["ll"]), ref["ll"]) < 1e-8          # every rank returns the same reduced result
        assert rel_err(z["g"], ref["g"]) < 1e-8
        Kinv += z["Kinv"]
    # each block of K^-1 is held by exactly one rank (diagonal blocks mirrored by their owner)
    Kref = np.linalg.inv(O.kern_compute(kern, X))
    assert np.max(np.abs(Kinv - Kref)) < 1e-9 * max(1.0, np.max(np.abs(Kref)))
