#!/usr/bin/env python3
"""BASELINE config 5: K sweep at 10M triplets on one B200 (E-step + M-step time per iteration), every E-step formulation
that exists for the K: slot-segmented (the default for K >= 4), K^3 per link (K <= 16), gene-segmented (K = 5..32).
    python tools/k_sweep.py [links] [K,K,...] [uniform|kuzmin] > profiles/rN_k_sweep_10M.jsonl"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trigenicinteractionpredictor_b200 import synth  # noqa: E402
from trigenicinteractionpredictor_b200.engine import EMEngine  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
Ks = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2, 3, 4, 6, 8, 10, 12, 16, 20, 24, 32]
shape = sys.argv[3] if len(sys.argv) > 3 else "uniform"
P = 6000
dev = torch.device("cuda:0")
g1, g2, g3, lab = (synth.kuzmin_links_soa if shape == "kuzmin" else synth.planted_links_soa)(P, L, seed=7, device=dev)
g1[:P] = torch.arange(P, dtype=torch.int32, device=dev)
out = []
runs = []
NAMES = {0: "K^3 per link", 8: "gene-segmented (2K^2 per link + scatter)", 32: "slot-segmented (4K^2 per link, no per-link atomics)",
         96: "slot-segmented, gather through L1"}
for K in Ks:
    runs.append((K, 96 if shape == "kuzmin" else 32))
    if K <= 16:
        runs.append((K, 0))
    if K >= 5:
        runs.append((K, 8))
for K, flags in runs:
    eng = EMEngine(P, K, device=dev, flags=flags)
    eng.set_train_links(g1, g2, g3, 1 - lab, lab)
    rng = np.random.default_rng(K)
    theta = rng.dirichlet(np.ones(K), size=P)
    pr = rng.random((K, K, K, 2))
    pr /= pr.sum(axis=3, keepdims=True)
    eng.set_params(theta, pr)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.em_iteration()
    torch.cuda.synchronize()
    n = 5 if (K <= 16 or flags & 32) else 2
    a.record()
    for _ in range(n):
        eng.em_iteration()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    rec = {"K": K, "links": L, "ms_per_iteration": ms, "link_updates_per_s": L / ms * 1e3,
           "algorithmic_tflops": 6.0 * K ** 3 * L / ms * 1e3 / 1e12, "flags": flags, "shape": shape,
           "kernel": NAMES[flags] if not (flags == 0 and K > 16) else NAMES[8]}
    out.append(rec)
    print(json.dumps(rec), flush=True)
    del eng
    torch.cuda.empty_cache()
