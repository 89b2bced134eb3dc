"""CPU restatement of the digenic extension (SURVEY f-4) - TEST INFRASTRUCTURE, never imported by the product.

The reference's src/TrigenicInteractionPredictor_23.py (`_23.py` below) extends the trigenic model with PAIR links
`dlinks` that share theta and have their own rating tensor qr[K][K][R].  Parity status: PINNED - the vectors under
tests/golden/digenic/ were produced by the author's own code, made importable by the two-token recipe documented in
oracle/gen_golden_23.py; tests/test_oracle_golden.py holds this restatement to them.

  digest_mixed       _23.py:393-450   mixed train file: a line with two gene names is a pair, with three a triplet
  init_params_23     _23.py:123-217   draw order: theta rows, pr cells, THEN qr cells; normalisations as in the base model
  em_step_23_np      _23.py:1572-1687 triplet term as the base model, pair term d = eps + sum_ij th_a th_b q_ij,r
  loglik_23_np       _23.py:1534-1562
"""
from __future__ import annotations

import random as _random

import numpy as np

from oracle import mmsbm_oracle as base

EPS = 1e-10
R = 2


def digest_mixed(train_lines):
    """ids by first appearance, string-sorted id keys (_23.py:402-431); returns (links3, links2, P) as dicts."""
    gene_id: dict[str, int] = {}
    links3: dict[str, list[int]] = {}
    links2: dict[str, list[int]] = {}
    for line in train_lines:
        fields = line.strip().split("\t")
        names = fields[0].split("_")
        if "hoΔ" in names:
            names.remove("hoΔ")
        rating = int(fields[1])
        ids = []
        for name in names:
            if name not in gene_id:
                gene_id[name] = len(gene_id)
            ids.append(str(gene_id[name]))
        ids.sort()
        table = links3 if len(names) == 3 else links2 if len(names) == 2 else None
        if table is None:
            continue
        table.setdefault("_".join(ids), [0, 0])[rating] += 1
    return links3, links2, len(gene_id)


def init_params_23(P: int, K: int, rng=_random):
    """(theta, pr, qr) with the reference's draw order and normalisations (_23.py:137-217)."""
    theta = [[rng.random() for _ in range(K)] for _ in range(P)]
    pr = [[[[rng.random() for _ in range(R)] for _ in range(K)] for _ in range(K)] for _ in range(K)]
    qr = [[[rng.random() for _ in range(R)] for _ in range(K)] for _ in range(K)]
    for g in range(P):
        acc = 0.0
        for k in range(K):
            acc += theta[g][k]
        if acc < EPS:
            theta[g] = [rng.random() for _ in range(K)]
        total = sum(theta[g])
        theta[g] = [v / total for v in theta[g]]
    th = np.array(theta, dtype=np.float64).reshape(P, K)
    p = np.array(pr, dtype=np.float64).reshape(K, K, K, R)
    q = np.array(qr, dtype=np.float64).reshape(K, K, R)
    p = p / ((0.0 + p[..., 0]) + p[..., 1])[..., None]
    q = q / ((0.0 + q[..., 0]) + q[..., 1])[..., None]
    return th, p, q


def pair_denominators(theta, qr, ids2):
    return EPS + np.einsum("li,lj,ijr->lr", theta[ids2[:, 0]], theta[ids2[:, 1]], qr, optimize=True)


def em_step_23_np(theta, pr, qr, ids3, cnt3, ids2, cnt2):
    """One make_iteration of the digenic model: returns (theta, pr, qr).  The triplet statistics come from the base
    oracle (same loops, _23.py:1575-1606 = TIP.py:987-1012); the pair term follows _23.py:1607-1635; the normalisers
    count every appearance of a gene in a DISTINCT triplet or pair (_23.py:1584-1586, 1615-1616, 1640-1643)."""
    P, K = theta.shape
    ntheta, npr, _ = base.em_step_np(theta, pr, ids3, cnt3, return_stats=True)
    ntheta = ntheta.copy()
    d = pair_denominators(theta, qr, ids2)                      # [L2][R]
    s = cnt2 / d
    ta, tb = theta[ids2[:, 0]], theta[ids2[:, 1]]
    u = np.einsum("lj,ijr,lr->li", tb, qr, s, optimize=True)     # sum_j th_b[j] q_ij,r s_r
    v = np.einsum("li,ijr,lr->lj", ta, qr, s, optimize=True)
    np.add.at(ntheta, ids2[:, 0], ta * u)
    np.add.at(ntheta, ids2[:, 1], tb * v)
    nq = qr * np.einsum("li,lj,lr->ijr", ta, tb, s, optimize=True)
    deg = np.bincount(ids3.ravel(), minlength=P) + np.bincount(ids2.ravel(), minlength=P)
    th_new = ntheta / deg[:, None].astype(np.float64)
    pr_new = npr / ((EPS + npr[..., 0]) + npr[..., 1])[..., None]
    qr_new = nq / ((EPS + nq[..., 0]) + nq[..., 1])[..., None]
    return th_new, pr_new, qr_new


def loglik_23_np(theta, pr, qr, ids3, cnt3, ids2, cnt2):
    return base.loglik_np(theta, pr, ids3, cnt3) + float((cnt2 * np.log(pair_denominators(theta, qr, ids2))).sum())
