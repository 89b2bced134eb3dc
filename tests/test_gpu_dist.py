"""Link-sharded data parallelism (needs >= 2 GPUs; skipped otherwise): two ranks, each with half of the links, the
statistics summed every iteration - once with an NCCL allreduce, once through NVLink peer memory fused into the
M-step kernel - must reproduce the single-GPU iteration."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_rank_link_shards_match_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29500 + os.getpid() % 1000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "_nccl_worker.py"), str(tmp_path)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    one = np.load(tmp_path / "single.npz")
    for exchange in ("nccl", "peer", "peer_rs", "peer_gather"):
        for r in range(2):
            got = np.load(tmp_path / ("r%d_%s.npz" % (r, exchange)))
            np.testing.assert_allclose(got["th"], one["th"], rtol=1e-10)
            np.testing.assert_allclose(got["p"], one["p"], rtol=1e-10)
            assert got["ll"] == pytest.approx(float(one["ll"]), rel=1e-11)
    # the peer-memory exchanges add the shards in rank order (and "peer" computes every value on one rank only):
    # replicas are bit-identical
    for mode in ("peer", "peer_rs", "peer_gather"):
        a, b = np.load(tmp_path / ("r0_%s.npz" % mode)), np.load(tmp_path / ("r1_%s.npz" % mode))
        assert np.array_equal(a["th"], b["th"]) and np.array_equal(a["p"], b["p"]), mode
