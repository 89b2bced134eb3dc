#!/usr/bin/env python3
"""Short driver for ncu: cfg2-shaped workload (6000 genes, 800k links, K=10), a few E-steps + M-steps."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trigenicinteractionpredictor_b200 import synth  # noqa: E402
from trigenicinteractionpredictor_b200.engine import EMEngine  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 10
L = int(sys.argv[2]) if len(sys.argv) > 2 else 800_000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
shape = sys.argv[5] if len(sys.argv) > 5 else "uniform"
P = 6000
dev = torch.device("cuda:0")
eng = EMEngine(P, K, device=dev, flags=flags)
if shape == "kuzmin":
    g1, g2, g3, lab = synth.kuzmin_links_soa(P, L, seed=100, device=dev)
else:
    g1, g2, g3, lab = synth.planted_links_soa(P, L, seed=100, device=dev)
g1[:P] = torch.arange(P, dtype=torch.int32, device=dev)
eng.set_train_links(g1, g2, g3, 1 - lab, lab)
rng = np.random.default_rng(0)
theta = rng.dirichlet(np.ones(K), size=P)
pr = rng.random((K, K, K, 2))
pr /= pr.sum(axis=3, keepdims=True)
eng.set_params(theta, pr)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for i in range(iters):
    a.record()
    eng.em_step()
    b.record()
    eng.normalise()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ts_s = sorted(ts[1:]) if len(ts) > 1 else ts
med = ts_s[len(ts_s) // 2]
print("em_step over %d iters: min %.3f ms  median %.3f ms  (%.3e link-updates/s at median)" % (
    len(ts_s), ts_s[0], med, L / med * 1e3))
ll = eng.loglik("train")
a.record()
for _ in range(5):
    ll = eng.loglik("train")
b.record()
torch.cuda.synchronize()
print("loglik %.6f   %.3f ms per call (incl. readback)" % (ll, a.elapsed_time(b) / 5))
