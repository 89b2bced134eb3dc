#!/usr/bin/env python3
"""Throughput of the cross-sample reduction (tip_reduce_samples + tip_metrics) at cfg3 size: S = 50 samples,
T = 200,000 test triplets (the fold-1 test split of 1M triplets).  HBM-bound: algorithmic bytes = 8*S read + 8*S
written (sorted copy) + 24 per triplet."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trigenicinteractionpredictor_b200 import _cabi  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 50
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
lib = _cabi.load()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
scores = torch.rand((S, T), dtype=torch.float64, device=dev, generator=g)
labels = (torch.rand(T, device=dev, generator=g) < 0.1).to(torch.int32)
srt = torch.empty_like(scores)
out = torch.empty((3, T), dtype=torch.float64, device=dev)
nb = ctypes.c_size_t(0)
assert lib.tip_metrics_workspace_bytes(T, ctypes.byref(nb)) == 0
ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
cnt = torch.zeros(8, dtype=torch.int64, device=dev)
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731


def run():
    rc = lib.tip_reduce_samples(S, T, p(scores), None, p(srt), p(out[0]), p(out[1]), p(out[2]), None)
    assert rc == 0, lib.tip_last_error()


def run_metrics():
    cnt.zero_()
    rc = lib.tip_metrics(p(out[0]), p(labels), T, int(0.1 * T), p(ws), nb.value, p(cnt), None)
    assert rc == 0, lib.tip_last_error()


res = {}
for name, fn in (("reduce", run), ("metrics", run_metrics)):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    res[name + "_ms"] = sorted(ts)[len(ts) // 2]
# check against numpy (mean is a sequential sum in sample order: cumulative sum reproduces it exactly)
ref_mean = (np.cumsum(scores.cpu().numpy(), axis=0)[-1]) / S
assert np.array_equal(out[0].cpu().numpy(), ref_mean)
bytes_alg = T * (16 * S + 24)
res.update({"S": S, "T": T, "algorithmic_bytes": bytes_alg, "reduce_gbs": bytes_alg / (res["reduce_ms"] * 1e-3) / 1e9,
            "triplets_per_s": T / (res["reduce_ms"] * 1e-3)})
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    res["hbm_peak_gbs"] = peak["hbm_gbs"]
    res["frac_of_hbm_peak"] = res["reduce_gbs"] / peak["hbm_gbs"]
except Exception:  # noqa: BLE001
    pass
print(json.dumps(res))
