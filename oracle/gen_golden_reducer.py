#!/usr/bin/env python3
"""Freeze golden vectors for the cross-sample reducer by RUNNING THE UNMODIFIED REFERENCE SCRIPT
`/root/reference/src/testResultsReducer.py` (build container only; nothing at test/bench time reads
/root/reference).

    python oracle/gen_golden_reducer.py        # rewrites tests/golden/reducer/

The script is a `__main__` block with hard-coded locations, so it is executed with `runpy` under these
(environment-only) arrangements, none of which touches its arithmetic:
  * cwd = <tmp>/src, so that its relative `../data/DATA_FOLDS/train{fold}.dat` resolves;
  * `open()` of its absolute output directory `/home/aleixmt/.../REDUCED_TRAIN_TEST_RESULTS/` is redirected to <tmp>/out;
  * `import matplotlib` (never used; not installed here) is satisfied with an empty module;
  * `os.walk` yields directories and files in sorted order (the script's sums follow the visiting order; the
    filesystem's raw order does not travel to another machine, sorted order does).
It needs every (K in 2..5) x (fold in 0..4) cell to hold at least one sample (it divides by the cell sizes), and
sample files that carry the `LIST OF REGISTERED GENES` block - the block `Model.to_string` used to emit and that is
commented out in the current TIP.py:863-867; the sample files here are the reference's own `to_string()` output plus
that block, written exactly as the commented code would.

Seeds: data_seed=11, np.random.seed(3) before fold(), random.seed(5000 + 100*K + 10*fold + sample) before init.
"""
from __future__ import annotations

import builtins
import contextlib
import io
import os
import random
import runpy
import shutil
import sys
import tempfile
import types

import numpy as np

REF_SRC = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden", "reducer")
REF_OUT_PREFIX = "/home/aleixmt/Escritorio/TrigenicInteractionPredictor/data/REDUCED_TRAIN_TEST_RESULTS/"
SAMPLES_PER_K = {2: 1, 3: 2, 4: 3, 5: 4}        # n = 1, 2, 3, 4: both parities, and n = 3 hits the round(n/2) quirk

sys.path.insert(0, ROOT)
sys.path.insert(0, REF_SRC)

from trigenicinteractionpredictor_b200 import synth  # noqa: E402


def gene_list_block(m) -> str:
    """The block of TIP.py:863-867 (commented out there) that testResultsReducer.py:101-113 parses."""
    text = "\nLIST OF REGISTERED GENES\n"
    text += "Gene_ID\tGene_name\tnumAparitions\n"
    for gid in m.id_gene:
        text += str(gid) + "\t" + m.id_gene[gid] + "\t" + str(m.uniqueg[gid]) + "\n"
    text += "\nLIST OF LINKS BETWEEN GENE IDS\n"
    return text


def main():
    import TrigenicInteractionPredictor as ref
    tmp = tempfile.mkdtemp(prefix="tip_reducer_")
    src, folds, results, out = (os.path.join(tmp, d) for d in ("src", "data/DATA_FOLDS", "results", "out"))
    for d in (src, folds, results, out):
        os.makedirs(d)
    # ---- data: 30 genes, 250 distinct triplets, folded by the reference
    P, n = 30, 250
    names = synth.gene_names(P)
    g, lab = synth.planted_triplets(P, n, seed=11, shape="uniform")
    raw = os.path.join(tmp, "input_s2.tsv")
    synth.write_raw_s2(raw, g, lab, names)
    m = ref.Model()
    with contextlib.redirect_stdout(io.StringIO()):
        m.get_input(raw)
    cwd = os.getcwd()
    os.chdir(folds)
    try:
        np.random.seed(3)
        m.fold()
    finally:
        os.chdir(cwd)
    # ---- sample files: reference Model, two EM iterations each, report + gene list block
    for K, ns in SAMPLES_PER_K.items():
        for fold in range(5):
            d = os.path.join(results, "K%d" % K, "fold%d" % fold)
            os.makedirs(d)
            for s in range(ns):
                mm = ref.Model()
                with contextlib.redirect_stdout(io.StringIO()):
                    mm.get_traintest(os.path.join(folds, "train%d.dat" % fold), os.path.join(folds, "test%d.dat" % fold))
                random.seed(5000 + 100 * K + 10 * fold + s)
                mm.initialize_parameters(K)
                mm.make_iteration()
                mm.make_iteration()
                mm.likelihood = mm.compute_likelihood()
                with open(os.path.join(d, "Sample_%d_K%d.csv" % (s, K)), "w", encoding="utf-8") as fh:
                    fh.write(mm.to_string() + gene_list_block(mm))
    # ---- run the reference reducer
    real_open, real_walk = builtins.open, os.walk

    def open_redirect(file, *a, **k):
        if isinstance(file, str) and file.startswith(REF_OUT_PREFIX):
            file = os.path.join(out, file[len(REF_OUT_PREFIX):])
        return real_open(file, *a, **k)

    def walk_sorted(top, *a, **k):
        for dp, dn, fn in real_walk(top, *a, **k):
            dn.sort()
            yield dp, dn, sorted(fn)

    sys.modules.setdefault("matplotlib", types.ModuleType("matplotlib"))
    argv = sys.argv
    os.chdir(src)
    try:
        builtins.open, os.walk = open_redirect, walk_sorted
        sys.argv = ["testResultsReducer.py", "-f", "../results/"]
        with contextlib.redirect_stdout(io.StringIO()):
            runpy.run_path(os.path.join(REF_SRC, "testResultsReducer.py"), run_name="__main__")
    finally:
        builtins.open, os.walk = real_open, real_walk
        sys.argv = argv
        os.chdir(cwd)
    produced = sorted(os.listdir(out))
    assert produced == sorted("K%d_fold%d.csv" % (K, f) for K in SAMPLES_PER_K for f in range(5)), produced
    # ---- freeze
    if os.path.isdir(GOLD):
        shutil.rmtree(GOLD)
    shutil.copytree(results, os.path.join(GOLD, "results"))
    os.makedirs(os.path.join(GOLD, "DATA_FOLDS"))
    for f in range(5):
        shutil.copy(os.path.join(folds, "train%d.dat" % f), os.path.join(GOLD, "DATA_FOLDS"))
    shutil.copytree(out, os.path.join(GOLD, "expected"))
    total = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(GOLD) for f in fs)
    print("reducer golden written to", GOLD, "bytes:", total)
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
