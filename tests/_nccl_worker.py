"""torchrun worker for tests/test_gpu_dist.py: link-sharded EM over NCCL vs the single-GPU iteration."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(out_dir):
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
    # what each rank would draw from its own pid-seeded stream (the reference seeds with os.getpid(), TIP.py:1149): ranks
    # other than 0 hand DIFFERENT parameters to set_params, which must make rank 0's copy the one everybody starts from
    mine = np.random.default_rng(50 + rk)
    theta_rk = theta if rk == 0 else mine.dirichlet(np.ones(K), size=P)
    pr_rk = pr if rk == 0 else mine.random((K, K, K, 2))
    lo, hi = tdist.shard_bounds(L, rk, w)
    for exchange in ("nccl", "peer", "peer_rs", "peer_gather"):
        eng = EMEngine(P, K, device=dev, group=torch.distributed.group.WORLD, exchange=exchange)
        eng.set_train_links(g[lo:hi, 0], g[lo:hi, 1], g[lo:hi, 2], 1 - lab[lo:hi], lab[lo:hi])   # deg is allreduced inside
        eng.set_params(theta_rk, pr_rk)
        print(rk, exchange, "links set", flush=True)
        eng.em_iterations(3, use_graph=False)            # eager, graph replays and eager again: one parity counter
        if rk == 1:
            torch.cuda._sleep(int(1e8))                  # rank skew (~50 ms): the other rank must wait in the exchange,
        eng.em_iterations(5, use_graph=True)             # not read half-written statistics (compute-sanitizer is closed here)
        if rk == 0:
            torch.cuda._sleep(int(6e7))
        eng.em_iteration()
        th, p = eng.get_params()
        ll = eng.loglik("train")
        if eng.peer is not None:
            eng.peer.check()
        print(rk, exchange, "sharded iterations done", ll, flush=True)
        np.savez(os.path.join(out_dir, "r%d_%s.npz" % (rk, exchange)), th=th, p=p, ll=ll)
    if rk == 0:
        ref = EMEngine(P, K, device=dev)
        ref.set_train_links(g[:, 0], g[:, 1], g[:, 2], 1 - lab, lab)
        ref.set_params(theta, pr)
        for _ in range(9):
            ref.em_iteration()
        th1, p1 = ref.get_params()
        np.savez(os.path.join(out_dir, "single.npz"), th=th1, p=p1, ll=ref.loglik("train"))
        print(rk, "single-GPU reference done", flush=True)
    tdist.barrier()
    torch.cuda.synchronize(dev)
    sys.stdout.flush()
    os._exit(0)      # skip ProcessGroupNCCL teardown, which was seen to hang on the GPU boxes


if __name__ == "__main__":
    main(sys.argv[1])
