"""Link-sharded data parallelism over NCCL (needs >= 2 GPUs; skipped otherwise): two ranks, each with half of the
links, one allreduce of the statistics per iteration, must reproduce the single-GPU iteration."""
import os
import sys

import numpy as np
import pytest

from tests.conftest import ROOT

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    from trigenicinteractionpredictor_b200 import dist as tdist
    from trigenicinteractionpredictor_b200.engine import EMEngine
    rk, w, local = tdist.init_from_env(backend="nccl")
    dev = torch.device("cuda", local)
    rng = np.random.default_rng(3)
    P, L, K = 500, 20000, 10
    g = rng.integers(0, P, size=(L, 3)).astype(np.int32)
    g[:P, 0] = np.arange(P)
    lab = (rng.random(L) < 0.2).astype(np.int32)
    theta = rng.dirichlet(np.ones(K), size=P)
    pr = rng.random((K, K, K, 2))
    pr /= pr.sum(axis=3, keepdims=True)
    lo, hi = tdist.shard_bounds(L, rk, w)
    eng = EMEngine(P, K, device=dev, group=torch.distributed.group.WORLD)
    eng.set_train_links(g[lo:hi, 0], g[lo:hi, 1], g[lo:hi, 2], 1 - lab[lo:hi], lab[lo:hi])   # deg is allreduced inside
    eng.set_params(theta, pr)
    eng.em_iterations(5, use_graph=True)
    th, p = eng.get_params()
    ll = eng.loglik("train")
    np.savez(os.path.join(out_dir, "r%d.npz" % rk), th=th, p=p, ll=ll)
    if rk == 0:
        ref = EMEngine(P, K, device=dev)
        ref.set_train_links(g[:, 0], g[:, 1], g[:, 2], 1 - lab, lab)
        ref.set_params(theta, pr)
        for _ in range(5):
            ref.em_iteration()
        th1, p1 = ref.get_params()
        np.savez(os.path.join(out_dir, "single.npz"), th=th1, p=p1, ll=ref.loglik("train"))
    tdist.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_link_shards_match_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world, port = 2, 29500 + os.getpid() % 1000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    one = np.load(tmp_path / "single.npz")
    for r in range(world):
        got = np.load(tmp_path / ("r%d.npz" % r))
        np.testing.assert_allclose(got["th"], one["th"], rtol=1e-10)
        np.testing.assert_allclose(got["p"], one["p"], rtol=1e-10)
        assert got["ll"] == pytest.approx(float(one["ll"]), rel=1e-11)
