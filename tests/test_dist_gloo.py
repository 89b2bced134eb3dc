"""world_size-2 gloo test of the link-sharded data path: shard bounds + allreduce of the statistics
buffer + M-step reproduce the single-process EM step.  The per-shard statistics come from the CPU
oracle here (the CUDA kernels need a GPU; the same plumbing runs over NCCL in tests/test_gpu_*)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from tests.conftest import GOLDEN, ROOT


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from oracle import mmsbm_oracle as orc
    from trigenicinteractionpredictor_b200 import dist as tdist
    rk, w, _ = tdist.init_from_env(backend="gloo")
    assert (rk, w) == (rank, world) and tdist.world_size() == world and tdist.rank() == rank
    base = os.path.join(GOLDEN, "base")
    dg = orc.digest_traintest(open(os.path.join(base, "train1.dat")).readlines(),
                              open(os.path.join(base, "test1.dat")).readlines())
    ids, cnt = orc.links_to_arrays(dg.links)
    tr = np.load(os.path.join(base, "trace_K3.npz"))
    lo, hi = tdist.shard_bounds(len(ids), rank, world)
    nt, npr, _ = orc.em_step_np(tr["theta0"], tr["pr0"], ids[lo:hi], cnt[lo:hi], return_stats=True)
    ll = orc.loglik_np(tr["theta0"], tr["pr0"], ids[lo:hi], cnt[lo:hi])
    stats = torch.from_numpy(np.concatenate([nt.ravel(), npr.ravel(), [ll]]))
    tdist.allreduce_sum_(stats)
    deg = torch.from_numpy(np.bincount(ids[lo:hi].ravel(), minlength=dg.P))
    tdist.allreduce_sum_(deg)
    tdist.barrier()
    # replicas must start from rank 0's parameters / seed whatever each rank drew (link shards)
    mine = torch.full((5,), float(rank + 1), dtype=torch.float64)
    tdist.broadcast_from_first_(mine)
    assert torch.equal(mine, torch.ones(5, dtype=torch.float64))
    assert tdist.broadcast_int(1234 + rank) == 1234
    mx = tdist.max_over_ranks(float(rank))
    s = stats.numpy()
    th, pr = orc.normalise_np(s[: nt.size].reshape(nt.shape), s[nt.size: nt.size + npr.size].reshape(npr.shape),
                              deg.numpy())
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), th=th, pr=pr, ll=s[-1], mx=mx)
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(300)
def test_link_sharded_step_matches_single_process(tmp_path):
    world, port = 2, 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    tr = np.load(os.path.join(GOLDEN, "base", "trace_K3.npz"))
    for r in range(world):
        got = np.load(tmp_path / ("r%d.npz" % r))
        np.testing.assert_allclose(got["th"], tr["theta1"], rtol=1e-12)
        np.testing.assert_allclose(got["pr"], tr["pr1"], rtol=1e-12)
        assert got["ll"] == pytest.approx(tr["loglik"][0], rel=1e-12)
        assert got["mx"] == world - 1
