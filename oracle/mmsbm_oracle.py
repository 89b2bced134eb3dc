"""CPU ORACLE - test infrastructure only.  NOT a product path.

A restatement, on the CPU, of the MMSBM hot path of AleixMT/TrigenicInteractionPredictor
(`src/TrigenicInteractionPredictor.py`, sha256 898269709cbf...f04290, called TIP.py below).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module; the shipped package never does (it fails loudly without its
CUDA library instead).

Parity status: PINNED.  `oracle/gen_golden.py` ran the unmodified reference in the build
container on seeded synthetic inputs and froze its outputs under `tests/golden/`;
`tests/test_oracle_golden.py` checks every function below against those vectors
(bit-exact for ids / folds / init, <=1e-12 relative for the floating-point trace).

Three flavours of the numeric path are kept:
  * ``*_loops``  - pure-Python loops in the reference's literal operation order (small cases,
                   and the CPython CPU baseline that bench.py times);
  * ``*_np``     - vectorised NumPy fp64 (fast; sizes up to ~1e6 links);
  * ``oracle/mmsbm_oracle.c`` - the literal loop order again in C (see `c_oracle()`).
"""
from __future__ import annotations

import ctypes
import math
import os
import random as _random
import re
import subprocess

import numpy as np

EPS = 1e-10  # TIP.py:88
R = 2        # TIP.py:79


# ----------------------------------------------------------------------------------------------
# Link digestion                                                            TIP.py:321-423
# ----------------------------------------------------------------------------------------------
class Digest:
    """Result of reading a train file then a test file."""

    def __init__(self):
        self.gene_id: dict[str, int] = {}
        self.id_gene: dict[int, str] = {}
        self.uniqueg: dict[int, int] = {}
        self.links: dict[str, list[int]] = {}
        self.nlinks: dict[str, list[int]] = {}
        self.test_links: dict[str, list[int]] = {}
        self.P = 0


def _register(dg: Digest, names: list[str], next_id: int):
    """ids in order of first appearance; every appearance bumps uniqueg (TIP.py:336-349)."""
    ids = []
    for nm in names:
        gid = dg.gene_id.get(nm)
        if gid is None:
            gid = next_id
            dg.gene_id[nm] = gid
            dg.id_gene[gid] = nm
            dg.uniqueg[gid] = 0
            next_id += 1
        dg.uniqueg[gid] += 1
        ids.append(str(gid))
    return ids, next_id


def digest_traintest(train_lines, test_lines) -> Digest:
    """TIP.py:321-413.  Train lines are split with strip().split('\\t') (329), test lines with
    re.split(r'\\t+') (376).  The link key joins the decimal id STRINGS after a string sort
    (353, 400): ids {9,10,11} give "10_11_9"."""
    dg = Digest()
    nxt = 0
    for line in train_lines:
        fields = line.strip().split("\t")
        names = fields[0].split("_")
        r = int(fields[1])
        ids, nxt = _register(dg, names, nxt)
        names.sort()
        ids.sort()
        nkey, key = "_".join(names), "_".join(ids)
        if key not in dg.links:
            # the reference touches links first and creates both entries on KeyError (360-368)
            dg.nlinks[nkey] = [0, 0]
            dg.links[key] = [0, 0]
            dg.nlinks[nkey][r] += 1
            dg.links[key][r] += 1
        else:
            dg.links[key][r] += 1
            dg.nlinks[nkey][r] += 1
    for line in test_lines:
        fields = re.split(r"\t+", line)
        names = fields[0].split("_")
        r = int(fields[1])
        ids, nxt = _register(dg, names, nxt)
        ids.sort()
        key = "_".join(ids)
        if key not in dg.test_links:
            dg.test_links[key] = [0, 0]
        dg.test_links[key][r] += 1
    dg.P = len(dg.id_gene)
    return dg


def links_to_arrays(links: dict[str, list[int]]):
    """dict order -> (ids[L,3] int64 in key slot order, counts[L,2] int64)."""
    L = len(links)
    ids = np.empty((L, 3), dtype=np.int64)
    cnt = np.empty((L, 2), dtype=np.int64)
    for n, (key, c) in enumerate(links.items()):
        a, b, d = key.split("_")
        ids[n] = (int(a), int(b), int(d))
        cnt[n] = c
    return ids, cnt


# ----------------------------------------------------------------------------------------------
# 5-fold split                                                               TIP.py:447-523
# ----------------------------------------------------------------------------------------------
def fold_texts(links: dict[str, list[int]], id_gene: dict[int, str], fraction: float = 0.2):
    """Returns (test_texts, train_texts): the exact contents the reference writes to
    test{i}.dat / train{i}.dat.  Consumes the GLOBAL numpy legacy stream through
    np.random.shuffle on the list of key strings, like TIP.py:455."""
    size = int(len(links) * fraction)          # 448
    nfold = int(1 / fraction)                  # 449
    keys = [k for k in links]                  # 452-454
    np.random.shuffle(keys)                    # 455
    tests = [keys[size * i: size * (i + 1)] for i in range(nfold)]   # 466-467
    tests[nfold - 1] += keys[size * nfold:]    # 470: remainder into the last fold

    def line(key):
        rating = 0 if links[key][0] else 1     # 475-477 / 512-514
        names = sorted(id_gene[int(t)] for t in key.split("_"))      # 478-486
        return "_".join(names) + "\t" + str(rating) + "\n"

    test_texts = ["".join(line(k) for k in tests[i]) for i in range(nfold)]
    train_texts = []
    for i in range(nfold):
        acc = []
        for other in tests[:i] + tests[i + 1:]:                      # 500-501
            acc.extend(other)
        train_texts.append("".join(line(k) for k in acc))
    return test_texts, train_texts


# ----------------------------------------------------------------------------------------------
# Parameter initialisation                                                   TIP.py:106-170
# ----------------------------------------------------------------------------------------------
def init_params(P: int, K: int, rng=_random):
    """Draw order: P*K theta draws, then K^3*R p draws (117-139); theta rows are normalised by
    the builtin sum() (151), p cells by a left-to-right running sum (162-170).  `rng` defaults
    to the global `random` module so the stream position matches a reference run."""
    theta = [[rng.random() for _ in range(K)] for _ in range(P)]
    pr = [[[[rng.random() for _ in range(R)] for _ in range(K)] for _ in range(K)] for _ in range(K)]
    for i in range(P):
        s = 0.0
        for k in range(K):
            s += theta[i][k]
        if s < EPS:                                    # 147-149
            theta[i] = [rng.random() for _ in range(K)]
        s = sum(theta[i])                              # 151
        for k in range(K):
            try:
                theta[i][k] /= s
            except ZeroDivisionError:
                theta[i][k] /= (s + EPS)
    for i in range(K):
        for j in range(K):
            for k in range(K):
                s = 0.0
                for r in range(R):
                    s += pr[i][j][k][r]
                for r in range(R):
                    try:
                        pr[i][j][k][r] /= s
                    except ZeroDivisionError:
                        pr[i][j][k][r] /= (s + EPS)
    return np.array(theta, dtype=np.float64).reshape(P, K), np.array(pr, dtype=np.float64).reshape(K, K, K, R)


# ----------------------------------------------------------------------------------------------
# EM step / likelihood / prediction: literal loops                 TIP.py:984-1043, 952-974, 530-547
# ----------------------------------------------------------------------------------------------
def em_step_loops(theta, pr, ids, cnt):
    """One make_iteration in the reference's operation order.  theta: list[P][K], pr: list
    [K][K][K][R] (nested Python lists), ids: list of (a,b,c), cnt: list of (n0,n1).
    Returns new (theta, pr) as nested lists.  Raises ZeroDivisionError for a gene with no
    training link, like TIP.py:1018."""
    P = len(theta)
    K = len(theta[0])
    ntheta = [[0.0] * K for _ in range(P)]
    npr = [[[[0.0] * R for _ in range(K)] for _ in range(K)] for _ in range(K)]
    deg = [0] * P
    for (a, b, c), n in zip(ids, cnt):
        ta, tb, tc = theta[a], theta[b], theta[c]
        d = [EPS] * R
        deg[a] += 1
        deg[b] += 1
        deg[c] += 1
        for i in range(K):
            for j in range(K):
                for k in range(K):
                    cell = pr[i][j][k]
                    for r in range(R):
                        d[r] += ta[i] * tb[j] * tc[k] * cell[r]
        for i in range(K):
            for j in range(K):
                for k in range(K):
                    cell = pr[i][j][k]
                    acc = npr[i][j][k]
                    for r in range(R):
                        w = (ta[i] * tb[j] * tc[k] * cell[r]) / d[r]
                        ntheta[a][i] += w * n[r]
                        ntheta[b][j] += w * n[r]
                        ntheta[c][k] += w * n[r]
                        acc[r] += w * n[r]
    for g in range(P):
        for k in range(K):
            ntheta[g][k] /= float(deg[g])
    for i in range(K):
        for j in range(K):
            for k in range(K):
                d = EPS
                for r in range(R):
                    d += npr[i][j][k][r]
                for r in range(R):
                    npr[i][j][k][r] /= d
    return ntheta, npr


def loglik_loops(theta, pr, ids, cnt):
    K = len(theta[0])
    total = 0.0
    for (a, b, c), n in zip(ids, cnt):
        ta, tb, tc = theta[a], theta[b], theta[c]
        d = [EPS] * R
        for i in range(K):
            for j in range(K):
                for k in range(K):
                    cell = pr[i][j][k]
                    for r in range(R):
                        d[r] += ta[i] * tb[j] * tc[k] * cell[r]
        for r in range(R):
            total += n[r] * math.log(d[r])
    return total


def predict_loops(theta, pr, a, b, c):
    """P(r=1) with no eps (TIP.py:531-539); the accumulator starts as int 0 like the reference."""
    K = len(theta[0])
    acc = 0
    for i in range(K):
        for j in range(K):
            for k in range(K):
                acc += theta[a][i] * theta[b][j] * theta[c][k] * pr[i][j][k][1]
    return acc


# ----------------------------------------------------------------------------------------------
# Vectorised NumPy versions (flat form, SURVEY 8a "verified restatements")
# ----------------------------------------------------------------------------------------------
def _chunks(n, step):
    for s in range(0, n, step):
        yield slice(s, min(n, s + step))


def em_step_np(theta, pr, ids, cnt, chunk: int = 8192, return_stats: bool = False):
    """theta[P,K], pr[K,K,K,R], ids[L,3], cnt[L,2] -> new (theta, pr).  fp64 throughout."""
    P, K = theta.shape
    ntheta = np.zeros((P, K))
    npr = np.zeros((K, K, K, R))
    deg = np.bincount(ids.reshape(-1), minlength=P).astype(np.float64)
    for sl in _chunks(ids.shape[0], chunk):
        a, b, c = ids[sl, 0], ids[sl, 1], ids[sl, 2]
        G = np.einsum("la,lb,lc->labc", theta[a], theta[b], theta[c])
        X = G[..., None] * pr[None]                           # [l,K,K,K,R]
        d = EPS + X.sum(axis=(1, 2, 3))                       # [l,R]
        s = cnt[sl] / d                                       # n_r / d_r
        W = X * s[:, None, None, None, :]
        Wr = W.sum(axis=4)
        np.add.at(ntheta, a, Wr.sum(axis=(2, 3)))
        np.add.at(ntheta, b, Wr.sum(axis=(1, 3)))
        np.add.at(ntheta, c, Wr.sum(axis=(1, 2)))
        npr += W.sum(axis=0)
    if return_stats:
        return ntheta, npr, deg
    if (deg == 0).any():
        raise ZeroDivisionError("float division by zero")     # TIP.py:1018
    ntheta /= deg[:, None]
    npr /= (EPS + npr[..., 0] + npr[..., 1])[..., None]
    return ntheta, npr


def normalise_np(ntheta, npr, deg):
    """M-step alone (TIP.py:1016-1028) - used to check shard-and-sum data parallelism."""
    if (np.asarray(deg) == 0).any():
        raise ZeroDivisionError("float division by zero")
    return ntheta / np.asarray(deg, dtype=np.float64)[:, None], npr / (EPS + npr[..., 0] + npr[..., 1])[..., None]


def denominators_np(theta, pr, ids, chunk: int = 65536):
    """sum_{abc} theta theta theta p  (no eps) for every link: [L,R]."""
    out = np.empty((ids.shape[0], R))
    for sl in _chunks(ids.shape[0], chunk):
        a, b, c = ids[sl, 0], ids[sl, 1], ids[sl, 2]
        q = np.einsum("abcr,lc->labr", pr, theta[c])
        out[sl] = np.einsum("la,lb,labr->lr", theta[a], theta[b], q)
    return out


def loglik_np(theta, pr, ids, cnt):
    d = EPS + denominators_np(theta, pr, ids)
    return float((cnt * np.log(d)).sum())


def scores_np(theta, pr, ids):
    """do_prediction for every test triplet (TIP.py:530-547): rating 1, no eps."""
    return denominators_np(theta, pr, ids)[:, 1]


# ----------------------------------------------------------------------------------------------
# Held-out results and metrics                                         TIP.py:557-569, 583-637
# ----------------------------------------------------------------------------------------------
def test_results(scores, test_links: dict[str, list[int]]):
    """[[score, key, label], ...] descending by (score, key string, label) - sort();reverse()."""
    res = []
    for s, (key, n) in zip(scores, test_links.items()):
        res.append([float(s), key, 0 if n[0] else 1])          # 560-563
    res.sort()
    res.reverse()
    return res


def metrics(results, links: dict[str, list[int]], n_test: int):
    """[precision, recall, fallout, auc] with the reference's conventions: positives_fraction
    counts train links with n1 == 1 exactly (588); cut = results[int(frac*T)][0], 0 if that index
    is past the end (595-599); AUC counts strict > pairs (613); predicted positive iff >= cut."""
    pos_train = sum(1 for n in links.values() if n[1] == 1)
    frac = pos_train / len(links)
    npos = int(frac * n_test)
    cut = results[npos][0] if npos < len(results) else 0
    pos = np.array([x[0] for x in results if x[2]], dtype=np.float64)
    neg = np.sort(np.array([x[0] for x in results if not x[2]], dtype=np.float64))
    # strict pairs: for each positive, negatives strictly below it
    wins = int(np.searchsorted(neg, pos, side="left").sum())
    auc = wins / (len(pos) * len(neg))
    tp = sum(1 for x in results if x[0] >= cut and x[2])
    fp = sum(1 for x in results if x[0] >= cut and not x[2])
    fn = sum(1 for x in results if not x[0] >= cut and x[2])
    tn = sum(1 for x in results if not x[0] >= cut and not x[2])
    return [tp / (tp + fp), tp / (tp + fn), fp / (fp + tn), auc]


def metrics_quadratic(results, links, n_test):
    """The O(pos*neg) double loop exactly as TIP.py:611-615 (tiny cases only)."""
    pos_train = sum(1 for n in links.values() if n[1] == 1)
    npos = int(pos_train / len(links) * n_test)
    cut = 0
    for i, row in enumerate(results):
        if i == npos:
            cut = row[0]
            break
    P_ = [x for x in results if x[2]]
    N_ = [x for x in results if not x[2]]
    wins = 0
    for p in P_:
        for n in N_:
            if p[0] > n[0]:
                wins += 1
    auc = wins / (len(P_) * len(N_))
    tp = fp = fn = tn = 0
    for s, _, y in results:
        if s >= cut:
            if y:
                tp += 1
            else:
                fp += 1
        elif y:
            fn += 1
        else:
            tn += 1
    return [tp / (tp + fp), tp / (tp + fn), fp / (fp + tn), auc]


# ----------------------------------------------------------------------------------------------
# Training loop semantics                                                    TIP.py:1253-1279
# ----------------------------------------------------------------------------------------------
def train_sample_np(theta, pr, ids, cnt, iterations, fcheck, bcheck):
    """Returns (theta, pr, converged, iterations_done, last_checked_loglik, trace)."""
    like0 = loglik_np(theta, pr, ids, cnt)
    trace = [like0]
    for it in range(iterations):
        theta, pr = em_step_np(theta, pr, ids, cnt)
        if it % fcheck == 0 and it > bcheck:
            like = loglik_np(theta, pr, ids, cnt)
            trace.append(like)
            if math.fabs((like - like0) / like0) < 0.01:
                return theta, pr, True, it + 1, like, trace
            like0 = like
    return theta, pr, False, iterations, like0, trace


# ----------------------------------------------------------------------------------------------
# C restatement (literal loop order), compiled on demand into oracle/_build/
# ----------------------------------------------------------------------------------------------
_HERE = os.path.dirname(os.path.abspath(__file__))
_C_LIB = None


def build_c_oracle(force: bool = False) -> str:
    src = os.path.join(_HERE, "mmsbm_oracle.c")
    out_dir = os.path.join(_HERE, "_build")
    out = os.path.join(out_dir, "libmmsbm_oracle.so")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        os.makedirs(out_dir, exist_ok=True)
        # -ffp-contract=off: no FMA contraction, so the arithmetic is the reference's op for op
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-fopenmp",
                               src, "-o", out, "-lm"])
    return out


def c_oracle():
    global _C_LIB
    if _C_LIB is None:
        lib = ctypes.CDLL(build_c_oracle())
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int64)
        lib.oracle_em_step.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64, ip, ip, dp, dp, dp, dp]
        lib.oracle_em_step.restype = ctypes.c_int
        lib.oracle_loglik.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64, ip, ip, dp, dp]
        lib.oracle_loglik.restype = ctypes.c_double
        lib.oracle_em_stats_mt.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64, ip, ip, dp, dp, dp, dp,
                                           ctypes.c_int]
        lib.oracle_em_stats_mt.restype = ctypes.c_int
        _C_LIB = lib
    return _C_LIB


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


def em_step_c(theta, pr, ids, cnt):
    """Literal-order C make_iteration; returns (theta, pr) or raises ZeroDivisionError."""
    lib = c_oracle()
    P, K = theta.shape
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    pr = np.ascontiguousarray(pr, dtype=np.float64)
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    cnt = np.ascontiguousarray(cnt, dtype=np.int64)
    nt = np.zeros_like(theta)
    npr = np.zeros_like(pr)
    rc = lib.oracle_em_step(P, K, ids.shape[0], _ip(ids), _ip(cnt), _dp(theta), _dp(pr), _dp(nt), _dp(npr))
    if rc == 1:
        raise ZeroDivisionError("float division by zero")
    return nt, npr


def loglik_c(theta, pr, ids, cnt):
    lib = c_oracle()
    P, K = theta.shape
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    pr = np.ascontiguousarray(pr, dtype=np.float64)
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    cnt = np.ascontiguousarray(cnt, dtype=np.int64)
    return float(lib.oracle_loglik(P, K, ids.shape[0], _ip(ids), _ip(cnt), _dp(theta), _dp(pr)))


def em_stats_c_mt(theta, pr, ids, cnt, threads: int):
    """Multi-threaded E-step (OpenMP over link blocks, per-thread statistics summed at the end).
    Same per-link arithmetic as the literal order; only the cross-link summation order differs."""
    lib = c_oracle()
    P, K = theta.shape
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    pr = np.ascontiguousarray(pr, dtype=np.float64)
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    cnt = np.ascontiguousarray(cnt, dtype=np.int64)
    nt = np.zeros_like(theta)
    npr = np.zeros_like(pr)
    lib.oracle_em_stats_mt(P, K, ids.shape[0], _ip(ids), _ip(cnt), _dp(theta), _dp(pr), _dp(nt), _dp(npr),
                           int(threads))
    return nt, npr
