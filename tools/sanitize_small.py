#!/usr/bin/env python3
"""Small end-to-end pass over every kernel family once on a tiny problem (written for compute-sanitizer memcheck /
racecheck; the sanitizer is closed on this GPU pool, so it serves as a quick all-kernels smoke pass) - K=10 fused + finalize, K=4 private-S, K=7 odd-K scatter, gene-segmented, K=20 large, streamed host rows
(both formats), likelihood (fused + segmented), scoring, metrics, reducer.
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trigenicinteractionpredictor_b200 import _cabi, testResultsReducer as trr  # noqa: E402
from trigenicinteractionpredictor_b200.engine import EMEngine  # noqa: E402

lib = _cabi.load()
rng = np.random.default_rng(0)
P, L, T = 120, 1500, 200
g = rng.integers(0, P, size=(L, 3)).astype(np.int32)
g[:P, 0] = np.arange(P)
lab = (rng.random(L) < 0.2).astype(np.int32)
gt = rng.integers(0, P, size=(T, 3)).astype(np.int32)
labt = (rng.random(T) < 0.2).astype(np.int32)
for K, flags in ((10, 0), (4, 0), (7, 0), (10, 8), (10, 2), (10, 4), (20, 0), (10, 1)):
    eng = EMEngine(P, K, flags=flags)
    eng.set_train_links(g[:, 0], g[:, 1], g[:, 2], 1 - lab, lab)
    eng.set_test_links(gt[:, 0], gt[:, 1], gt[:, 2], 1 - labt, labt)
    theta = rng.dirichlet(np.ones(K), size=P)
    pr = rng.random((K, K, K, 2))
    pr /= pr.sum(axis=3, keepdims=True)
    eng.set_params(theta, pr)
    eng.em_iteration()
    ll = eng.loglik("train")
    sc = eng.scores()
    c = eng.metric_counts(sc, 40)
    if K <= 10 and flags == 0:
        rows_h = eng.train.rows.cpu().pin_memory()
        rows8_h = torch.empty(eng.train.n_rows, dtype=torch.int64).pin_memory()
        assert lib.tip_rows_compact_host(rows_h.data_ptr(), eng.train.n_rows, rows8_h.data_ptr()) == 0
        eng.em_iteration_host_rows(rows_h, False)
        eng.em_iteration_host_rows(rows8_h, True)
        torch.cuda.synchronize()
        assert eng.host_rows_arrived()
    torch.cuda.synchronize()
    print("K", K, "flags", flags, "loglik", ll, "wins", c["wins"], flush=True)
cols = [rng.random(1 + t % 9).tolist() for t in range(300)]
mean, med, std, c = trr.reduce_cell_on_device(cols, (rng.random(300) < 0.3).astype(int).tolist(), 0.3)
print("reducer ok", float(mean[0]), c["wins"])
