#!/usr/bin/env python3
"""Golden record of BASELINE.json configs[0] produced by RUNNING THE UNMODIFIED REFERENCE (build container only):
synthetic 1,000 genes x 100,000 labelled triplets, K=2, 1 sample, 100 EM iterations, fold 1 of 5 train/test.

    python oracle/gen_golden_cfg1.py        # ~5 min of CPython; writes tests/golden/cfg1/

The inputs are regenerated at test time (synth is deterministic; the fold files are ~2.5 MB each and are not
committed): the record holds their sha256, the log-likelihood after every iteration, the final p, sampled rows
of the final theta, the held-out likelihood, the sorted-table head and the metrics.
Seeds: data_seed=1, np.random.seed(2) before fold(), random.seed(1000) before initialize_parameters(2)."""
from __future__ import annotations

import contextlib
import hashlib
import io
import json
import os
import random
import sys
import tempfile
import time

import numpy as np

REF_SRC = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden", "cfg1")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF_SRC)

from trigenicinteractionpredictor_b200 import synth  # noqa: E402

P, N, K, ITERS, FOLD = 1000, 100_000, 2, 100, 1


def sha(path):
    with open(path, "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def main():
    import TrigenicInteractionPredictor as ref
    tmp = tempfile.mkdtemp(prefix="tip_cfg1_")
    g, lab = synth.planted_triplets(P, N, seed=1, shape="uniform")
    raw = os.path.join(tmp, "input_s2.tsv")
    synth.write_raw_s2(raw, g, lab, synth.gene_names(P))
    m = ref.Model()
    with contextlib.redirect_stdout(io.StringIO()):
        m.get_input(raw)
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        np.random.seed(2)
        m.fold()
    finally:
        os.chdir(cwd)
    train, test = os.path.join(tmp, "train%d.dat" % FOLD), os.path.join(tmp, "test%d.dat" % FOLD)
    mm = ref.Model()
    with contextlib.redirect_stdout(io.StringIO()):
        mm.get_traintest(train, test)
    random.seed(1000)
    mm.initialize_parameters(K)
    like = [mm.compute_likelihood()]
    t0 = time.time()
    for it in range(ITERS):
        mm.make_iteration()
        like.append(mm.compute_likelihood())
    em_s = time.time() - t0
    heldout = mm.compute_likelihood("test")
    mm.calculate_test_set_results()
    metrics = mm.calculate_metrics()
    theta = np.array(mm.theta, dtype=np.float64)
    rows = list(range(0, mm.P, 37))
    os.makedirs(OUT, exist_ok=True)
    rec = {
        "config": "BASELINE.json configs[0]: 1000 genes x 100000 triplets, K=2, 100 iterations, fold 1 of 5",
        "P": mm.P, "train_links": len(mm.links), "test_links": len(mm.test_links),
        "sha256": {"raw": sha(raw), "train": sha(train), "test": sha(test)},
        "loglik": like, "heldout": heldout, "metrics": metrics,
        "pr_final": np.array(mm.pr, dtype=np.float64).reshape(-1).tolist(),
        "theta_rows": rows, "theta_final_rows": theta[rows].tolist(),
        "theta_final_sum": float(theta.sum()), "theta_final_sq": float((theta ** 2).sum()),
        "results_head": [[r[0], r[1], r[2]] for r in mm.results[:50]],
        "reference_seconds_for_100_iterations_with_likelihood": em_s,
    }
    with open(os.path.join(OUT, "record.json"), "w") as fh:
        json.dump(rec, fh, indent=1)
    print("cfg1 golden written:", OUT, "reference EM loop took %.1f s" % em_s)


if __name__ == "__main__":
    main()
