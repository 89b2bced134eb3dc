#!/usr/bin/env python3
"""Where does the host-buffer entry spend its time?  n_iter sweep of tip_em_iterations_host."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trigenicinteractionpredictor_b200 import synth, _cabi
from trigenicinteractionpredictor_b200.engine import EMEngine
P, K, L = 6000, 10, 800_000
dev = torch.device("cuda:0")
eng = EMEngine(P, K, device=dev)
g1, g2, g3, lab = synth.planted_links_soa(P, L, seed=100, device=dev)
g1[:P] = torch.arange(P, dtype=torch.int32, device=dev)
eng.set_train_links(g1, g2, g3, 1 - lab, lab)
rng = np.random.default_rng(0)
theta = rng.dirichlet(np.ones(K), size=P); pr = rng.random((K, K, K, 2)); pr /= pr.sum(axis=3, keepdims=True)
lib = _cabi.load()
rows_h = eng.train.rows.cpu().pin_memory(); deg_h = eng.train.deg.cpu().pin_memory()
th_h = torch.from_numpy(theta.copy()).pin_memory(); p_h = torch.from_numpy(pr.copy()).pin_memory()
rows8_h = torch.empty(eng.train.n_rows, dtype=torch.int64).pin_memory()
assert lib.tip_rows_compact_host(rows_h.data_ptr(), eng.train.n_rows, rows8_h.data_ptr()) == 0
# raw H2D copy times for reference
d16 = torch.empty_like(eng.train.rows); d8 = torch.empty(eng.train.n_rows, dtype=torch.int64, device=dev)
for name, src, dst in (("rows16", rows_h, d16), ("rows8", rows8_h, d8)):
    ts = []
    for rep in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter(); dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    print("H2D %s %.1f MB: min %.3f ms" % (name, src.numel() * src.element_size() / 1e6, 1e3 * min(ts)))
for compact in (0, 1):
    src = rows8_h if compact else rows_h
    for n_iter in (0, 1, 2, 5):
        ts = []
        for rep in range(6):
            t0 = time.perf_counter()
            rc = lib.tip_em_iterations_host(P, K, src.data_ptr(), eng.train.n_rows, eng.train.n_rows_r0, deg_h.data_ptr(),
                                            th_h.data_ptr(), p_h.data_ptr(), n_iter, 16 if compact else 0)
            ts.append(time.perf_counter() - t0)
            assert rc == 0, lib.tip_last_error()
        print("compact %d n_iter %2d: min %.3f ms  median %.3f ms" % (compact, n_iter, 1e3 * min(ts), 1e3 * sorted(ts)[3]), flush=True)
