#!/usr/bin/env python3
"""Uniform vs hub-shaped (Kuzmin) links at cfg2 size: ms per EM iteration for each E-step formulation.

    python tools/hub_probe.py [--links 800000] [--K 10] [--flags 0,8,32] > gpurun_out/hub_probe.jsonl
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trigenicinteractionpredictor_b200 import synth  # noqa: E402
from trigenicinteractionpredictor_b200.engine import EMEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--links", type=int, default=800_000)
    ap.add_argument("--P", type=int, default=6000)
    ap.add_argument("--K", type=int, default=10)
    ap.add_argument("--flags", default="0,8")
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    P, K, L = args.P, args.K, args.links
    rng = np.random.default_rng(0)
    theta0 = rng.dirichlet(np.ones(K), size=P)
    pr0 = rng.random((K, K, K, 2))
    pr0 /= pr0.sum(axis=3, keepdims=True)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ref = {}
    for shape in ("uniform", "kuzmin"):
        if shape == "uniform":
            g1, g2, g3, lab = synth.planted_links_soa(P, L, seed=100, device=dev)
        else:
            g1, g2, g3, lab = synth.kuzmin_links_soa(P, L, seed=100, device=dev)
        g1[:P] = torch.arange(P, dtype=torch.int32, device=dev)
        for fl in [int(x) for x in args.flags.split(",")]:
            eng = EMEngine(P, K, device=dev, flags=fl)
            eng.set_train_links(g1, g2, g3, 1 - lab, lab)
            eng.set_params(theta0, pr0)
            eng.em_iteration()
            th, p = eng.get_params()
            key = shape
            if key not in ref:
                ref[key] = (th, p)
            err = max(float(np.max(np.abs(th - ref[key][0]) / np.maximum(np.abs(ref[key][0]), 1e-300))),
                      float(np.max(np.abs(p - ref[key][1]) / np.maximum(np.abs(ref[key][1]), 1e-300))))
            eng.set_params(theta0, pr0)
            eng.capture_graphs()
            for _ in range(3):
                flush.zero_()
                eng.graph_step()
            ts = []
            for _ in range(args.steps):
                flush.zero_()
                a.record()
                eng.graph_step()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            ms = float(np.mean(ts))
            print(json.dumps({"shape": shape, "flags": fl, "K": K, "links": L, "ms_per_iteration": ms,
                              "link_updates_per_s": L / (ms * 1e-3), "rel_err_vs_first_flags": err}), flush=True)
            del eng


if __name__ == "__main__":
    main()
