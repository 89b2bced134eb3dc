#!/usr/bin/env python3
"""Golden vectors for the digenic extension (SURVEY f-4): the reference's src/TrigenicInteractionPredictor_23.py, which adds
pair links `dlinks` with their own rating tensor qr[K][K][R] to the trigenic model (make_iteration _23.py:1572-1687,
compute_likelihood _23.py:1534-1562, initialize_parameters _23.py:123-217, get_train_test _23.py:393-560).

    python oracle/gen_golden_23.py            (build container only: needs /root/reference)

THE PATCH RECIPE.  The author's file does not run as shipped: `DataType.ALL` (line 105) and `self.dataType.ALL` (lines
197, 213) name an enum member that is spelled `all` (lines 31-34).  The recipe is exactly two textual substitutions,
    DataType.ALL  ->  DataType.all          dataType.ALL  ->  dataType.all
applied to a temporary copy (never committed; the sha256 of the original is recorded in the golden file).  Nothing else
is touched; every number below is produced by the author's own loops.

Input: the triplets of tests/golden/base/train1.dat plus pair lines derived from them (the two lexicographically first
genes of every third triplet, label = the triplet's label, a few of them repeated so that counts above one occur), in the
mixed train-file format get_train_test reads (a line with two names is a pair)."""
import hashlib
import io
import os
import random
import sys
import tempfile
import contextlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = "/root/reference/src/TrigenicInteractionPredictor_23.py"
OUT = os.path.join(ROOT, "tests", "golden", "digenic")


def patched_module():
    text = open(SRC, encoding="utf-8").read()
    sha = hashlib.sha256(text.encode("utf-8")).hexdigest()
    text = text.replace("DataType.ALL", "DataType.all").replace("dataType.ALL", "dataType.all")
    tmp = tempfile.mkdtemp(prefix="tip23_")
    path = os.path.join(tmp, "tip23_patched.py")
    with open(path, "w", encoding="utf-8") as fh:
        fh.write(text)
    import importlib.util
    spec = importlib.util.spec_from_file_location("tip23_patched", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, sha


def mixed_train_lines():
    lines = open(os.path.join(ROOT, "tests", "golden", "base", "train1.dat"), encoding="utf-8").read().splitlines()
    out = []
    for n, line in enumerate(lines):
        out.append(line)
        if n % 3 == 0:
            names, lab = line.split("\t")
            a, b, _ = names.split("_")
            out.append("%s_%s\t%s" % (a, b, lab))
            if n % 21 == 0:
                out.append("%s_%s\t%s" % (b, a, lab))          # the same pair again, other order: count 2
            if n % 33 == 0:
                out.append("%s_%s\t%d" % (a, b, 1 - int(lab)))  # a conflicting sighting
    return out


def dict_arrays(table, width):
    ids = np.array([[int(t) for t in k.split("_")] for k in table], dtype=np.int64).reshape(len(table), width)
    cnt = np.array(list(table.values()), dtype=np.int64).reshape(len(table), 2)
    return ids, cnt


def main():
    mod, sha = patched_module()
    os.makedirs(OUT, exist_ok=True)
    train = os.path.join(OUT, "train_mixed.dat")
    test = os.path.join(OUT, "test_mixed.dat")
    with open(train, "w", encoding="utf-8") as fh:
        fh.write("\n".join(mixed_train_lines()) + "\n")
    with open(test, "w", encoding="utf-8") as fh:
        fh.write("\n".join(mixed_train_lines()[:40]) + "\n")
    for K in (2, 3, 10):
        m = mod.Model()
        with contextlib.redirect_stdout(io.StringIO()):
            m.get_train_test(train, test)
        random.seed(2300 + K)
        m.initialize_parameters(K)
        rec = {"sha256_reference_23": sha, "P": m.P, "K": K}
        rec["ids3"], rec["cnt3"] = dict_arrays(m.links, 3)
        rec["ids2"], rec["cnt2"] = dict_arrays(m.dlinks, 2)
        rec["theta0"], rec["pr0"], rec["qr0"] = np.array(m.theta), np.array(m.pr), np.array(m.qr)
        ll = [m.compute_likelihood()]
        for it in range(3 if K == 10 else 5):
            m.make_iteration()
            rec["theta%d" % (it + 1)] = np.array(m.theta)
            rec["pr%d" % (it + 1)] = np.array(m.pr)
            rec["qr%d" % (it + 1)] = np.array(m.qr)
            ll.append(m.compute_likelihood())
        rec["loglik"] = np.array(ll)
        rec["rng_next"] = random.random()
        np.savez_compressed(os.path.join(OUT, "trace23_K%d.npz" % K), **rec)
        print("K=%d: P=%d triplets=%d pairs=%d loglik %s" % (K, m.P, len(m.links), len(m.dlinks), ll))


if __name__ == "__main__":
    sys.exit(main())
