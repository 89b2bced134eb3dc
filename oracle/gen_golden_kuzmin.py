#!/usr/bin/env python3
"""Golden vectors on HUB-SHAPED data, frozen by RUNNING THE UNMODIFIED REFERENCE (build container only).

The Kuzmin-2018 screen is (query pair) x (array gene) (TIP.py:272-273: `fields[1].split('+')` + `fields[3]`):
a few query genes take part in thousands of triplets.  This case pins parity there: P = 300 genes,
20 query genes, 3,000 triplets -> get_input -> fold -> train1/test1, EM traces at K = 3 (5 iterations)
and K = 10 (3 iterations), held-out scores, sorted table and metrics.  Only tests/golden/kuzmin/ is
rewritten (oracle/gen_golden.py owns the other cases).

    python oracle/gen_golden_kuzmin.py
"""
from __future__ import annotations

import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import gen_golden as gg  # noqa: E402  (reuses em_trace / digest_record; importing does not run it)
from trigenicinteractionpredictor_b200 import synth  # noqa: E402

OUT = os.path.join(gg.GOLD, "kuzmin")


def main():
    assert gg.sha256_file(os.path.join(gg.REF_SRC, "TrigenicInteractionPredictor.py")) == gg.REF_SHA256
    import TrigenicInteractionPredictor as ref

    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    P, n, nq = 300, 3000, 20
    names = synth.gene_names(P)
    g, lab = synth.planted_triplets(P, n, seed=11, shape="kuzmin", n_query=nq)
    raw = os.path.join(OUT, "input_s2.tsv")
    synth.write_raw_s2(raw, g, lab, names)
    m = ref.Model()
    m.get_input(raw)
    cwd = os.getcwd()
    os.chdir(OUT)
    try:
        np.random.seed(2)
        m.fold()
    finally:
        os.chdir(cwd)
    for f in os.listdir(OUT):                       # keep fold 1 only
        if f.endswith(".dat") and f not in ("train1.dat", "test1.dat"):
            os.remove(os.path.join(OUT, f))
    os.remove(raw)
    train, test = os.path.join(OUT, "train1.dat"), os.path.join(OUT, "test1.dat")
    md = ref.Model()
    gg.quiet(md.get_traintest, train, test)
    with open(os.path.join(OUT, "digest.json"), "w") as fh:
        json.dump(gg.digest_record(md), fh)
    # hub statistics of the training split, per key slot (the decimal-string slot order decides where hubs land)
    deg = np.zeros((3, md.P), dtype=np.int64)
    for key in md.links:
        for s, t in enumerate(key.split("_")):
            deg[s, int(t)] += 1
    info = {"P": md.P, "train_links": len(md.links), "test_links": len(md.test_links), "n_query": nq,
            "max_degree_per_slot": deg.max(axis=1).tolist(),
            "genes_with_degree_over_100_per_slot": (deg > 100).sum(axis=1).tolist(),
            "reference_sha256": gg.REF_SHA256,
            "fold_sha256": {f: gg.sha256_file(os.path.join(OUT, f)) for f in ("train1.dat", "test1.dat")}}
    for K, iters in ((3, 5), (10, 3)):
        gg.em_trace(ref, train, test, K, 1000, iters, os.path.join(OUT, "trace_K%d.npz" % K))
    with open(os.path.join(OUT, "info.json"), "w") as fh:
        json.dump(info, fh, indent=1)
    print(json.dumps(info))


if __name__ == "__main__":
    main()
