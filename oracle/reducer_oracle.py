"""CPU restatement of the reference's cross-sample reducer - TEST INFRASTRUCTURE ONLY.

Follows `/root/reference/src/testResultsReducer.py` (cited below as TRR.py:line) in plain Python, in the
reference's own operation order.  Parity status: PINNED - `tests/test_reducer_oracle.py` checks it, byte for
byte, against the output files the unmodified reference script wrote for the fixture under
`tests/golden/reducer/` (`oracle/gen_golden_reducer.py` ran it).  Only `tests/` may import this module; the
product (`trigenicinteractionpredictor_b200/testResultsReducer.py`) runs the reduction on the GPU and never
routes through here.
"""
from __future__ import annotations

import math
import os
import re


def walk_sample_files(results_folder):
    """TRR.py:76-87: every file below the folder, skipping empty files and lock files (name ending in '#').
    Directories and files are visited in sorted order (the reference follows the raw os.walk order, which is a
    property of the filesystem, not of the algorithm - the golden run pins it to sorted as well)."""
    for dirpath, dirnames, filenames in os.walk(results_folder):
        dirnames.sort()
        for f in sorted(filenames):
            path = os.path.join(dirpath, f)
            if os.stat(path).st_size == 0 or f[-1] == "#":
                continue
            yield path


def cell_of_path(path):
    """TRR.py:90-97: (K, fold, sample) from '.../K{k}/fold{f}/Sample_{s}_K{k}.csv' (lstrip of character sets)."""
    k_number, fold_number, sample_number = path.split("/")[-3:]
    return int(k_number.lstrip("K")), int(fold_number.lstrip("fold")), sample_number.split("_")[1]


def parse_gene_list(path):
    """TRR.py:101-113: gene names in id order from the 'LIST OF REGISTERED GENES' block (ends at a blank line)."""
    names = []
    with open(path) as fh:
        line = fh.readline()
        while not re.match("LIST OF REGISTERED GENES", line):
            line = fh.readline()
            if line == "":
                raise ValueError("no LIST OF REGISTERED GENES block in " + path)  # the reference loops forever here
        fh.readline()
        line = fh.readline()
        while line:
            names.append(line.split("\t")[1])
            line = fh.readline().rstrip("\n")
    return names


def parse_sample(path):
    """TRR.py:116-135: held-out likelihood and the (score, 'id_id_id', label) rows of the 'Test set:' block."""
    rows = []
    with open(path) as fh:
        line = fh.readline()
        while not re.match("Held-out Likelihood", line):
            line = fh.readline()
        heldout = float(line.split("\t")[1])
        while not re.match("Test set:", line):
            line = fh.readline()
        fh.readline()
        while line:
            line = fh.readline().rstrip("\n")
            if not line:
                break
            prob, triplet, real = line.split("\t")
            rows.append((float(prob), triplet, int(real)))
    return heldout, rows


def reduce_values(values):
    """TRR.py:163-186 for one (triplet, K, fold): returns (mean, median, std).  `values` is sorted in place, as in
    the reference, so the deviation sums over the SORTED values."""
    n = len(values)
    sum_total = 0
    for i in range(n):
        sum_total += values[i]
    mean = sum_total / n
    values.sort()
    if n % 2:
        median = values[round(n / 2)]          # Python 3 rounds halves to even: n=3 -> index 2, n=7 -> index 4
    else:
        half = int(n / 2)
        median = sum(values[half - 1:half + 1]) / 2
    sum_square_diff = 0
    for i in range(n):
        sum_square_diff += (values[i] - mean) ** 2
    return mean, median, math.sqrt(sum_square_diff / n)


def metrics_of_cell(records, density):
    """TRR.py:211-258: records = [key, mean, median, std, label] sorted by mean descending (stable).
    Returns (auc, precision, recall, fallout); raises ZeroDivisionError like the reference."""
    predicted = int(density * len(records))
    cut, counter = 0, 0
    for rec in records:
        if predicted == counter:
            cut = rec[1]
            break
        counter += 1
    positives = [r for r in records if r[4]]
    negatives = [r for r in records if not r[4]]
    wins = 0
    for p in positives:
        for q in negatives:
            if p[1] > q[1]:
                wins += 1
    auc = wins / (len(positives) * len(negatives))
    tp = fp = fn = tn = 0
    for r in records:
        if r[1] >= cut:
            if r[4]:
                tp += 1
            else:
                fp += 1
        else:
            if r[4]:
                fn += 1
            else:
                tn += 1
    return auc, tp / (tp + fp), tp / (tp + fn), fp / (fp + tn)


def training_density(folds_folder, fold):
    """TRR.py:199-208: fraction of lines of train{fold}.dat whose label is non-zero."""
    pos = cnt = 0
    with open(os.path.join(folds_folder, "train" + str(fold) + ".dat")) as fh:
        for line in fh.readlines():
            cnt += 1
            if int(line.split("\t")[1]):
                pos += 1
    return float(pos) / float(cnt)


def format_cell(likelihood_mean, auc, precision, recall, fallout, records):
    """TRR.py:259-264."""
    out = ["\nHeld-OutLikelihoodMean\tAUCmean\tPrecision\tRecall\tFallout\n",
           str(likelihood_mean) + "\t" + str(auc) + "\t" + str(precision) + "\t" + str(recall) + "\t" + str(fallout) + "\n",
           "\nTripleteName\tMean\tMedian\tStdDev\tRealinteraction\n"]
    for s in records:
        out.append(str(s[0]) + "\t" + str(s[1]) + "\t" + str(s[2]) + "\t" + str(s[3]) + "\t" + str(s[4]) + "\n")
    return "".join(out)


def reduce_folder(results_folder, folds_folder):
    """The whole script: {(K, fold): text of K{K}_fold{fold}.csv}.  Cells are the (K, fold) pairs that occur (the
    reference walks a fixed 4 x 5 grid, K = 2..5, and divides by zero on an empty cell)."""
    gene_names = {}                 # fold -> names by id, restored once per fold (TRR.py:101)
    data = {}                       # triplet name key -> {"cells": {(K, fold): [scores]}, "real": label}
    likelihoods = {}
    for path in walk_sample_files(results_folder):
        K, fold, _ = cell_of_path(path)
        if fold not in gene_names:
            gene_names[fold] = parse_gene_list(path)
        heldout, rows = parse_sample(path)
        likelihoods.setdefault((K, fold), []).append(heldout)
        for prob, triplet, real in rows:
            names = sorted(gene_names[fold][int(g)] for g in triplet.split("_"))
            key = "_".join(names)
            if key not in data:
                data[key] = {"cells": {}, "real": real}         # label appended only the first time (TRR.py:149)
            data[key]["cells"].setdefault((K, fold), []).append(prob)
    cells = {}
    for key, value in data.items():                              # insertion order = first appearance
        for cell, vals in value["cells"].items():
            mean, median, std = reduce_values(vals)
            cells.setdefault(cell, []).append([key, mean, median, std, value["real"]])
    out = {}
    for cell in sorted(cells):
        K, fold = cell
        records = cells[cell]
        records.sort(key=lambda tup: tup[1], reverse=True)
        lk = likelihoods[cell]
        auc, precision, recall, fallout = metrics_of_cell(records, training_density(folds_folder, fold))
        out[cell] = format_cell(float(sum(lk) / len(lk)), auc, precision, recall, fallout, records)
    return out
