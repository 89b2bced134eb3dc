#!/usr/bin/env python3
"""Cross-sample reducer: drop-in for the reference's `testResultsReducer.py` (cited as TRR.py:line).

Reads the per-sample reports `.../K{k}/fold{f}/Sample_{s}_K{k}.csv` below a results folder, and for every
(K, fold) cell writes `K{k}_fold{f}.csv`: mean held-out likelihood, the AUC / precision / recall / fallout of the
MEAN score of every test triplet across samples, and the per-triplet mean / median / standard deviation, sorted
by mean (descending).

    python -m trigenicinteractionpredictor_b200.testResultsReducer -f ../data/DEFINITIVE_RESULTS/ \
        [--folds ../data/DATA_FOLDS/] [--out ../data/REDUCED_TRAIN_TEST_RESULTS/] [--device cuda:0]

Host side (this file): the reference's text parsing, key construction and report layout.  Device side
(libtip.so, no CPU path): `tip_reduce_samples` - mean / median / deviation of all triplets of a cell in one
launch, in the reference's operation order (mean and median bit-identical to CPython) - and `tip_metrics`, the
exact integer pair count that replaces the O(pos * neg) loop of TRR.py:236-240.

Differences from the reference, all at the edges: the folds folder and the output folder are arguments (the
reference hard-codes `../data/DATA_FOLDS/` and an absolute path under /home/aleixmt); any set of K values and
folds is accepted (the reference walks a fixed K = 2..5 x fold 0..4 grid and divides by zero on an empty cell);
directories and files are visited in sorted order (the reference follows the filesystem's raw os.walk order);
a sample file without the `LIST OF REGISTERED GENES` block raises ValueError (the reference loops forever on it).
`Model.to_string()` of the current TIP.py no longer writes that block (it is commented out, TIP.py:863-867);
`gene_list_block(model)` reproduces it for callers that want reducible reports.
"""
from __future__ import annotations

import ctypes
import getopt
import os
import re
import sys

import numpy as np

__all__ = ["reduce_results", "gene_list_block", "main"]


def gene_list_block(model) -> str:
    """The report block TRR.py:101-113 parses, as TIP.py:863-867 (commented out there) would write it."""
    out = ["\nLIST OF REGISTERED GENES\n", "Gene_ID\tGene_name\tnumAparitions\n"]
    for gid in model.id_gene:
        out.append(str(gid) + "\t" + model.id_gene[gid] + "\t" + str(model.uniqueg[gid]) + "\n")
    out.append("\nLIST OF LINKS BETWEEN GENE IDS\n")
    return "".join(out)


# ----------------------------------------------------------------------------------------------
# host: parsing (TRR.py:76-154)
# ----------------------------------------------------------------------------------------------
def _sample_files(results_folder, log):
    for dirpath, dirnames, filenames in os.walk(results_folder):
        dirnames.sort()
        for f in sorted(filenames):
            path = os.path.join(dirpath, f)
            if os.stat(path).st_size == 0:
                log("· WARNING! Found empty file ·")
                continue
            if f[-1] == "#":
                log("· WARNING! Found lock file. Skipping... ·")
                continue
            yield path


def _read_gene_names(path):
    names = []
    with open(path) as fh:
        line = fh.readline()
        while not re.match("LIST OF REGISTERED GENES", line):
            if line == "":
                raise ValueError("%s: no 'LIST OF REGISTERED GENES' block (see gene_list_block)" % path)
            line = fh.readline()
        fh.readline()                                   # column names
        line = fh.readline()
        while line:
            names.append(line.split("\t")[1])
            line = fh.readline().rstrip("\n")
    return names


def _read_sample(path):
    with open(path) as fh:
        line = fh.readline()
        while not re.match("Held-out Likelihood", line):
            if line == "":
                raise ValueError("%s: no 'Held-out Likelihood' line" % path)
            line = fh.readline()
        heldout = float(line.split("\t")[1])
        while not re.match("Test set:", line):
            if line == "":
                raise ValueError("%s: no 'Test set:' block" % path)
            line = fh.readline()
        fh.readline()                                   # column names
        rows = []
        while True:
            line = fh.readline().rstrip("\n")
            if not line:
                break
            prob, triplet, real = line.split("\t")
            rows.append((float(prob), triplet, int(real)))
    return heldout, rows


def _training_density(folds_folder, fold):
    """TRR.py:199-208."""
    pos = cnt = 0
    with open(os.path.join(folds_folder, "train" + str(fold) + ".dat"), "r") as fh:
        for line in fh.readlines():
            cnt += 1
            if int(line.split("\t")[1]):
                pos += 1
    return float(pos) / float(cnt)


# ----------------------------------------------------------------------------------------------
# device: one cell = one tip_reduce_samples launch + one tip_metrics call
# ----------------------------------------------------------------------------------------------
def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def reduce_cell_on_device(columns, labels, density, device=None):
    """columns: list over triplets of the score list of that triplet (sample order); labels: 0/1 per triplet.
    Returns (mean, median, std) as float64 numpy arrays and the integer metric counts of the mean scores."""
    import torch
    from . import _cabi
    if not torch.cuda.is_available():
        raise _cabi.TipLibraryError("no CUDA device: trigenicinteractionpredictor_b200 has no CPU path")
    lib = _cabi.load()
    dev = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
    T = len(columns)
    n = np.fromiter((len(c) for c in columns), dtype=np.int32, count=T)
    S = int(n.max()) if T else 1
    mat = np.zeros((S, T), dtype=np.float64)             # sample-major: [j][t]
    for t, col in enumerate(columns):
        mat[: len(col), t] = col
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        d_mat = torch.from_numpy(mat).to(dev)
        d_n = torch.from_numpy(n).to(dev)
        d_sorted = torch.empty_like(d_mat)
        d_out = torch.empty((3, max(T, 1)), dtype=torch.float64, device=dev)
        _cabi.check(lib.tip_reduce_samples(S, T, _ptr(d_mat), _ptr(d_n), _ptr(d_sorted), _ptr(d_out[0]), _ptr(d_out[1]),
                                           _ptr(d_out[2]), st), "tip_reduce_samples")
        counts = {"wins": 0, "n_pos": 0, "n_neg": 0, "tp": 0, "fp": 0, "fn": 0, "tn": 0}
        if T:
            d_lab = torch.from_numpy(np.asarray(labels, dtype=np.int32)).to(dev)
            nb = ctypes.c_size_t(0)
            _cabi.check(lib.tip_metrics_workspace_bytes(T, ctypes.byref(nb)), "tip_metrics_workspace_bytes")
            ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
            cnt = torch.zeros(8, dtype=torch.int64, device=dev)
            # predictedNumberOfPositivesTest = int(density * T); the cut is the mean at that rank (TRR.py:218-227)
            _cabi.check(lib.tip_metrics(_ptr(d_out[0]), _ptr(d_lab), T, int(density * T), _ptr(ws), nb.value, _ptr(cnt),
                                        st), "tip_metrics")
            o = cnt.cpu().numpy()
            counts = {"wins": int(o[0]), "n_pos": int(o[1]), "n_neg": int(o[2]), "tp": int(o[3]), "fp": int(o[4]),
                      "fn": int(o[5]), "tn": int(o[6])}
        res = d_out.cpu().numpy()
    return res[0, :T], res[1, :T], res[2, :T], counts


def _format_cell(likelihood_mean, auc, precision, recall, fallout, records):
    """TRR.py:259-264."""
    out = ["\nHeld-OutLikelihoodMean\tAUCmean\tPrecision\tRecall\tFallout\n",
           str(likelihood_mean) + "\t" + str(auc) + "\t" + str(precision) + "\t" + str(recall) + "\t" + str(fallout) + "\n",
           "\nTripleteName\tMean\tMedian\tStdDev\tRealinteraction\n"]
    for s in records:
        out.append(str(s[0]) + "\t" + str(s[1]) + "\t" + str(s[2]) + "\t" + str(s[3]) + "\t" + str(s[4]) + "\n")
    return "".join(out)


def reduce_results(results_folder, folds_folder="../data/DATA_FOLDS/", out_folder=None, device=None, log=print):
    """Run the whole reduction.  Returns {(K, fold): {"text", "likelihood", "auc", "precision", "recall",
    "fallout", "records"}}; with `out_folder` also writes `K{K}_fold{fold}.csv` there."""
    log("· Start reading files ·")
    gene_names = {}                  # fold -> names by id, restored once per fold (TRR.py:101)
    data = {}                        # name key -> [cells {(K, fold): [scores]}, real label]   (insertion ordered)
    likelihoods = {}
    for path in _sample_files(results_folder, log):
        k_number, fold_number, sample_number = path.split("/")[-3:]              # TRR.py:90-97
        k_number = k_number.lstrip("K")
        fold_number = fold_number.lstrip("fold")
        sample_number = sample_number.split("_")[1]
        log("Dataset K=" + k_number + " fold=" + fold_number + " sample=" + sample_number)
        cell = (int(k_number), int(fold_number))
        if cell[1] not in gene_names:
            gene_names[cell[1]] = _read_gene_names(path)
        heldout, rows = _read_sample(path)
        likelihoods.setdefault(cell, []).append(heldout)
        names_of = gene_names[cell[1]]
        for prob, triplet, real in rows:
            names = [names_of[int(g)] for g in triplet.split("_")]
            names.sort()
            key = "_".join(names)
            entry = data.get(key)
            if entry is None:
                entry = data[key] = [{}, real]                                  # label kept from the first sighting
            entry[0].setdefault(cell, []).append(prob)
    log("· Reducing results ·")
    per_cell = {}
    for key, (cells, real) in data.items():
        for cell, vals in cells.items():
            keys, cols, labs = per_cell.setdefault(cell, ([], [], []))
            keys.append(key)
            cols.append(vals)
            labs.append(real)
    log("· Compute metrics ·")
    out = {}
    for cell in sorted(per_cell):
        K, fold = cell
        keys, cols, labs = per_cell[cell]
        density = _training_density(folds_folder, fold)
        mean, median, std, c = reduce_cell_on_device(cols, labs, density, device)
        records = [[keys[t], float(mean[t]), float(median[t]), float(std[t]), labs[t]] for t in range(len(keys))]
        records.sort(key=lambda tup: tup[1], reverse=True)                      # stable, like TRR.py:216
        auc = c["wins"] / (c["n_pos"] * c["n_neg"])                             # ZeroDivisionError as TRR.py:241
        precision = c["tp"] / (c["tp"] + c["fp"])
        recall = c["tp"] / (c["tp"] + c["fn"])
        fallout = c["fp"] / (c["fp"] + c["tn"])
        lk = likelihoods[cell]
        lk_mean = float(sum(lk) / len(lk))
        text = _format_cell(lk_mean, auc, precision, recall, fallout, records)
        out[cell] = {"text": text, "likelihood": lk_mean, "auc": auc, "precision": precision, "recall": recall,
                     "fallout": fallout, "records": records}
        if out_folder is not None:
            os.makedirs(out_folder, exist_ok=True)
            with open(os.path.join(out_folder, "K" + str(K) + "_fold" + str(fold) + ".csv"), "w+") as fh:
                fh.write(text)
    return out


def main(argv=None) -> int:
    """Same flag as the reference (-f / --folder, TRR.py:28-41; bad flag or missing folder -> exit status 2), plus
    --folds, --out and --device for what the reference hard-codes."""
    argv = sys.argv[1:] if argv is None else list(argv)
    results_folder = "../data/DEFINITIVE_RESULTS/"
    folds_folder = "../data/DATA_FOLDS/"
    out_folder = "../data/REDUCED_TRAIN_TEST_RESULTS/"
    device = None
    try:
        opts, _ = getopt.getopt(argv, "f:", ["folder=", "folds=", "out=", "device="])
        for opt, arg in opts:
            if opt in ("-f", "--folder"):
                if os.path.exists(str(arg)):
                    results_folder = arg
                else:
                    print("\n\nERROR: The selected path does not exist.")
                    raise ValueError
            elif opt == "--folds":
                folds_folder = arg
            elif opt == "--out":
                out_folder = arg
            elif opt == "--device":
                device = arg
    except getopt.GetoptError:
        return 2
    except ValueError:
        return 2
    reduce_results(results_folder, folds_folder, out_folder, device)
    return 0


if __name__ == "__main__":
    sys.exit(main())
