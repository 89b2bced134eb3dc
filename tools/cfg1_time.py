#!/usr/bin/env python3
"""BASELINE configs[0] through the drop-in Model: 1,000 genes x 100,000 triplets, fold 1 of 5, K=2, 100 EM iterations
with the likelihood after every one (what oracle/gen_golden_cfg1.py timed for the reference: 187 s of CPython)."""
import json
import os
import random
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trigenicinteractionpredictor_b200 import Model, synth  # noqa: E402

tmp = tempfile.mkdtemp()
g, lab = synth.planted_triplets(1000, 100_000, seed=1, shape="uniform")
raw = os.path.join(tmp, "input_s2.tsv")
synth.write_raw_s2(raw, g, lab, synth.gene_names(1000))
t0 = time.perf_counter()
m = Model()
m.get_input(raw)
os.chdir(tmp)
np.random.seed(2)
m.fold()
t_fold = time.perf_counter() - t0
t0 = time.perf_counter()
mm = Model()
mm.get_traintest("train1.dat", "test1.dat")
t_digest = time.perf_counter() - t0
random.seed(1000)
mm.initialize_parameters(2)
mm.compute_likelihood()                      # first call builds the device state
import torch  # noqa: E402
torch.cuda.synchronize()
t0 = time.perf_counter()
like = []
for _ in range(100):
    mm.make_iteration()
    like.append(mm.compute_likelihood())
torch.cuda.synchronize()
t_em = time.perf_counter() - t0
mm.make_iterations(100)                     # (captures the CUDA graph of one iteration the first time)
torch.cuda.synchronize()
t0 = time.perf_counter()
mm.make_iterations(100)
torch.cuda.synchronize()
t_em_only = time.perf_counter() - t0
t0 = time.perf_counter()
mm.calculate_test_set_results()
met = mm.calculate_metrics()
t_score = time.perf_counter() - t0
print(json.dumps({"config": "cfg1: 1000 genes x 100k triplets, K=2, fold 1, 100 iterations", "train_links": len(mm.links),
                  "get_input_and_fold_s": t_fold, "get_traintest_s": t_digest,
                  "em_100_iterations_with_likelihood_s": t_em, "em_100_iterations_graph_replay_s": t_em_only,
                  "link_updates_per_s_with_likelihood": len(mm.links) * 100 / t_em,
                  "link_updates_per_s": len(mm.links) * 100 / t_em_only, "scoring_and_metrics_s": t_score,
                  "final_loglik": like[-1], "auc": met[3],
                  "reference_cpython_s_for_the_same_loop": 187.0}))
