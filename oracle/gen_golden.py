#!/usr/bin/env python3
"""Freeze golden vectors by RUNNING THE UNMODIFIED REFERENCE (build container only).

The reference (`/root/reference/src/TrigenicInteractionPredictor.py`) ships no tests and no
readable data, so parity is pinned by executing it here on seeded synthetic inputs and
committing its outputs under `tests/golden/`.  `/root/reference` does not exist on the GPU
box; nothing at test/bench time imports it - only this script does.

    python oracle/gen_golden.py            # rewrites tests/golden/

Seeds: data_seed=1 (numpy Generator, private), np.random.seed(2) before fold(),
random.seed(1000 + sample) before every initialize_parameters().
"""
from __future__ import annotations

import contextlib
import hashlib
import io
import json
import math
import os
import random
import shutil
import sys
import tempfile

import numpy as np

REF_SRC = "/root/reference/src"
REF_SHA256 = "898269709cbfcb2dd5ec01d501224ff63c2a6676406f8d15779af0f531f04290"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")

sys.path.insert(0, ROOT)
sys.path.insert(0, REF_SRC)

from trigenicinteractionpredictor_b200 import synth  # noqa: E402


def sha256_file(path):
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def flat_theta(m):
    return np.array(m.theta, dtype=np.float64).reshape(m.P, m.K)


def flat_pr(m):
    return np.array(m.pr, dtype=np.float64).reshape(m.K, m.K, m.K, m.R)


def digest_record(m):
    return {
        "genes_in_id_order": [m.id_gene[i] for i in range(m.P)],
        "uniqueg": [m.uniqueg[i] for i in range(m.P)],
        "link_keys": list(m.links.keys()),
        "link_counts": [list(v) for v in m.links.values()],
        "nlink_keys": list(m.nlinks.keys()),
        "test_keys": list(m.test_links.keys()),
        "test_counts": [list(v) for v in m.test_links.values()],
        "P": m.P,
    }


def em_trace(ref, train, test, K, seed, iters, out_npz):
    m = ref.Model()
    quiet(m.get_traintest, train, test)
    random.seed(seed)
    m.initialize_parameters(K)
    rec = {"theta0": flat_theta(m), "pr0": flat_pr(m)}
    rec["after_init_random"] = np.array([random.random()])  # stream position check
    random.seed(seed)
    m.initialize_parameters(K)
    like = [m.compute_likelihood()]
    for it in range(iters):
        m.make_iteration()
        rec["theta%d" % (it + 1)] = flat_theta(m)
        rec["pr%d" % (it + 1)] = flat_pr(m)
        like.append(m.compute_likelihood())
    rec["loglik"] = np.array(like)
    rec["heldout"] = np.array([m.compute_likelihood("test")])
    m.calculate_test_set_results()
    rec["result_scores"] = np.array([r[0] for r in m.results], dtype=np.float64)
    rec["result_keys"] = np.array([r[1] for r in m.results])
    rec["result_labels"] = np.array([r[2] for r in m.results], dtype=np.int64)
    rec["scores_test_order"] = np.array(
        [m.do_prediction(*k.split("_")) for k in m.test_links.keys()], dtype=np.float64)
    rec["metrics"] = np.array(m.calculate_metrics(), dtype=np.float64)
    # name-based prediction path (TIP.py:544-545)
    k0 = next(iter(m.test_links.keys())).split("_")
    names = [m.id_gene[int(t)] for t in k0]
    rec["predict_by_name"] = np.array([m.do_prediction(*names), m.do_prediction(*k0)])
    np.savez_compressed(out_npz, **rec)
    return m


def training_loop(ref, train, test, K, sample, iterations, fcheck, bcheck):
    """TIP.py:1253-1279 driven through the reference Model (the __main__ block cannot be imported,
    and it seeds with os.getpid(), so the loop is replayed here with random.seed(1000+sample))."""
    m = ref.Model()
    quiet(m.get_traintest, train, test)
    random.seed(1000 + sample)
    m.initialize_parameters(K)
    like0 = m.compute_likelihood()
    checks = [like0]
    for it in range(iterations):
        m.make_iteration()
        if it % fcheck == 0 and it > bcheck:
            like = m.compute_likelihood()
            checks.append(like)
            if math.fabs((like - like0) / like0) < 0.01:
                return m, True, it, checks, m.to_string()
            like0 = like
    return m, False, iterations, checks, None


def main():
    assert sha256_file(os.path.join(REF_SRC, "TrigenicInteractionPredictor.py")) == REF_SHA256
    import TrigenicInteractionPredictor as ref

    if os.path.isdir(GOLD):
        shutil.rmtree(GOLD)
    os.makedirs(GOLD)
    manifest = {"reference_sha256": REF_SHA256, "python": sys.version.split()[0], "numpy": np.__version__}

    # ---------------------------------------------------------------- base: P=100, 2000 triplets
    base = os.path.join(GOLD, "base")
    os.makedirs(base)
    P, n = 100, 2000
    names = synth.gene_names(P)
    g, lab = synth.planted_triplets(P, n, seed=1, shape="uniform")
    raw = os.path.join(base, "input_s2.tsv")
    synth.write_raw_s2(raw, g, lab, names)
    m = ref.Model()
    m.get_input(raw)
    cwd = os.getcwd()
    os.chdir(base)
    try:
        np.random.seed(2)
        m.fold()
    finally:
        os.chdir(cwd)
    manifest["base"] = {
        "P": P, "triplets": n,
        "fold_sha256": {f: sha256_file(os.path.join(base, f))
                        for f in sorted(os.listdir(base)) if f.endswith(".dat")},
        "get_input": {"genes_in_id_order": [m.id_gene[i] for i in range(m.P)],
                      "link_keys_head": list(m.links.keys())[:50],
                      "n_links": len(m.links),
                      "n_pos": sum(1 for v in m.links.values() if v[1])},
    }
    train, test = os.path.join(base, "train1.dat"), os.path.join(base, "test1.dat")
    md = ref.Model()
    quiet(md.get_traintest, train, test)
    with open(os.path.join(base, "digest.json"), "w") as fh:
        json.dump(digest_record(md), fh)
    for K in (1, 2, 3, 10):
        em_trace(ref, train, test, K, 1000, 5, os.path.join(base, "trace_K%d.npz" % K))

    # training loop + to_string report (K=2, short check cadence so it converges quickly)
    mm, conv, it, checks, text = training_loop(ref, train, test, 2, 0, 300, 5, 10)
    assert conv, "golden training loop did not converge"
    with open(os.path.join(base, "Sample_0_K2.csv"), "w", encoding="utf-8") as fh:
        fh.write(text)
    manifest["base"]["loop"] = {"K": 2, "iterations": 300, "fcheck": 5, "bcheck": 10,
                                "converged_at_iteration": it, "checks": checks,
                                "likelihood": mm.likelihood}

    # ---------------------------------------------------------------- dups: duplicates, conflicts, string-sort trap
    dups = os.path.join(GOLD, "dups")
    os.makedirs(dups)
    rng = np.random.default_rng(7)
    P2 = 14
    nm = synth.gene_names(P2)
    lines = []
    for _ in range(120):
        a, b, c = rng.choice(P2, size=3, replace=False).tolist()
        tri = [nm[a], nm[b], nm[c]]
        rng.shuffle(tri)                      # unsorted names in the file: digestion must sort
        lines.append("_".join(tri) + "\t" + str(int(rng.random() < 0.3)) + "\n")
    lines += lines[:25]                        # exact duplicates (counts of 2)
    flip = [l.rsplit("\t", 1)[0] + "\t" + ("1" if l.strip().endswith("0") else "0") + "\n" for l in lines[25:45]]
    lines += flip                              # conflicting labels (both ratings seen)
    tl = []
    for _ in range(40):
        a, b, c = rng.choice(P2, size=3, replace=False).tolist()
        tl.append("_".join(sorted([nm[a], nm[b], nm[c]])) + "\t" + str(int(rng.random() < 0.3)) + "\n")
    tl += tl[:5]
    with open(os.path.join(dups, "train.dat"), "w") as fh:
        fh.writelines(lines)
    with open(os.path.join(dups, "test.dat"), "w") as fh:
        fh.writelines(tl)
    md = em_trace(ref, os.path.join(dups, "train.dat"), os.path.join(dups, "test.dat"), 3, 1001, 3,
                  os.path.join(dups, "trace_K3.npz"))
    with open(os.path.join(dups, "digest.json"), "w") as fh:
        json.dump(digest_record(md), fh)

    # ---------------------------------------------------------------- testonly: gene seen only in the test file
    to = os.path.join(GOLD, "testonly")
    os.makedirs(to)
    with open(os.path.join(to, "train.dat"), "w") as fh:
        fh.write("A_B_C\t1\nA_B_D\t0\nB_C_D\t0\nA_C_D\t1\n")
    with open(os.path.join(to, "test.dat"), "w") as fh:
        fh.write("A_B_E\t1\nB_C_D\t0\n")
    mt = ref.Model()
    quiet(mt.get_traintest, os.path.join(to, "train.dat"), os.path.join(to, "test.dat"))
    random.seed(5)
    mt.initialize_parameters(2)
    try:
        mt.make_iteration()
        exc = None
    except Exception as e:  # noqa: BLE001
        exc = type(e).__name__
    manifest["testonly"] = {"P": mt.P, "exception": exc, "digest": digest_record(mt)}

    with open(os.path.join(GOLD, "manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    total = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(GOLD) for f in fs)
    print("golden written to", GOLD, "bytes:", total)


if __name__ == "__main__":
    main()
