#!/usr/bin/env python3
"""Drop-in `Model` for the DIGENIC extension of the reference, src/TrigenicInteractionPredictor_23.py (`_23.py`): the
trigenic model plus links between two genes (`dlinks`) that share theta and have their own rating tensor `qr[K][K][R]`.

Scope (SURVEY f-4): the training path - `get_train_test` (mixed train files: a line with two gene names is a pair,
_23.py:393-450), `initialize_parameters` (_23.py:123-217: theta rows, pr cells, then qr cells from the same
`random.random()` stream), `make_iteration` (_23.py:1572-1687) and `compute_likelihood` (_23.py:1534-1562).  Same method
names, attributes (`theta`, `pr`, `qr`, `links`, `dlinks`, `test_links`, `dtest_links`, `gene_id`, `id_gene`,
`gene_num_aparitions`, `likelihood`, `K`, `P`, `R`, `eps`) and error behaviour as the author's class; the numerics run on
the device: the triplet term through the same E-step kernels as the trigenic `Model`, the pair term through
`tip_pairs_step` / `tip_pairs_normalise` / `tip_pairs_loglik` (csrc/tip_pairs.cu).  The author's file does not run as
shipped (an enum member is misspelled); oracle/gen_golden_23.py documents the two-token recipe that makes it importable
and froze the vectors this class is tested against.  Prediction / test-set scoring of pairs is not part of this slice.

There is no CPU fallback: without libtip.so and a CUDA device the numeric methods raise."""
from __future__ import annotations

import codecs
import random
import re

import numpy as np

__all__ = ["Model"]


class Model:
    def __init__(self, device=None, flags: int | None = None):
        self.id_gene, self.gene_id = {}, {}
        self.links, self.nLinks = {}, {}
        self.dlinks, self.ndlinks = {}, {}
        self.test_links, self.dtest_links = {}, {}
        self.gene_num_aparitions = {}
        self.results = []
        self.likelihood = 0
        self.likelihoodVector = []
        self.R, self.K, self.P = 2, 0, 0
        self.eps = 1e-10
        self._device, self._flags = device, flags
        self._engine = None
        self._theta, self._pr, self._qr = [], [], []
        self._on_device = False

    # ---------------------------------------------------------------- parameters (host mirrors, lists like the reference's)
    def _pull(self):
        if self._on_device:
            th, pr = self._engine.get_params()
            self._theta, self._pr, self._qr = th.tolist(), pr.tolist(), self._engine.get_q().tolist()
            self._on_device = False

    @property
    def theta(self):
        self._pull()
        return self._theta

    @property
    def pr(self):
        self._pull()
        return self._pr

    @property
    def qr(self):
        self._pull()
        return self._qr

    def initialize_parameters(self, k=2, interaction=None):
        try:
            self.K = int(k)
        except ValueError:
            self.K = 2
        self.dataType = interaction
        self.likelihoodVector = []
        K, R, rnd = self.K, self.R, random.random
        # draw order of _23.py:137-172: every theta row, every pr cell (i, j, k, r), then every qr cell (i, j, r)
        theta = [[rnd() for _ in range(K)] for _ in range(self.P)]
        pr = [[[[rnd() for _ in range(R)] for _ in range(K)] for _ in range(K)] for _ in range(K)]
        qr = [[[rnd() for _ in range(R)] for _ in range(K)] for _ in range(K)]
        for g in range(self.P):
            acc = 0.0
            for v in theta[g]:
                acc += v
            if acc < self.eps:                               # _23.py:180-182
                theta[g] = [rnd() for _ in range(K)]
            total = sum(theta[g])
            row = theta[g]
            for kk in range(K):
                try:
                    row[kk] /= total
                except ZeroDivisionError:
                    row[kk] /= (total + self.eps)
        for cells in ([c for plane in pr for line in plane for c in line], [c for line in qr for c in line]):
            for cell in cells:
                acc = 0.0
                for r in range(R):
                    acc += cell[r]
                for r in range(R):
                    try:
                        cell[r] /= acc
                    except ZeroDivisionError:
                        cell[r] /= (acc + self.eps)
        self._theta, self._pr, self._qr = theta, pr, qr
        self._on_device = False
        self._uploaded = False

    # ---------------------------------------------------------------- digestion (_23.py:393-560)
    def _register(self, names, next_id):
        ids = []
        for name in names:
            gid = self.gene_id.get(name)
            if gid is None:
                gid = next_id
                next_id += 1
                self.gene_id[name] = gid
                self.id_gene[gid] = name
                self.gene_num_aparitions[gid] = 0
            self.gene_num_aparitions[gid] += 1
            ids.append(str(gid))
        return ids, next_id

    @staticmethod
    def _bump(table, key, r):
        cell = table.get(key)
        if cell is None:
            cell = table[key] = [0, 0]
        cell[r] += 1

    def get_train_test(self, train_file_path, test_file_path):
        try:
            next_id = 0
            with codecs.open(train_file_path, encoding='utf-8', mode='r') as fh:
                for line in fh.readlines():
                    fields = line.strip().split('\t')
                    names = fields[0].split('_')
                    if 'hoΔ' in names:
                        names.remove('hoΔ')
                    rating = int(fields[1])
                    ids, next_id = self._register(names, next_id)
                    names.sort()
                    ids.sort()                               # STRING sort of the decimal ids (_23.py:427)
                    if len(names) == 3:
                        self._bump(self.links, '_'.join(ids), rating)
                        self._bump(self.nLinks, '_'.join(names), rating)
                    if len(names) == 2:
                        self._bump(self.dlinks, '_'.join(ids), rating)
                        self._bump(self.ndlinks, '_'.join(names), rating)
                self.P = len(self.id_gene)
            print('number of triplets, pairs', len(self.links), len(self.dlinks))
            with codecs.open(test_file_path, encoding='utf-8', mode='r') as fh:
                for line in fh.readlines():
                    fields = re.split(r'\t+', line)
                    names = fields[0].split('_')
                    if 'hoΔ' in names:
                        names.remove('hoΔ')
                    rating = int(fields[1])
                    ids, next_id = self._register(names, next_id)
                    names.sort()
                    ids.sort()
                    key = '_'.join(ids)
                    self._bump(self.test_links, key, rating)          # every line, and once more for a triplet (_23.py:502-516)
                    if len(names) == 3:
                        self._bump(self.test_links, key, rating)
                    if len(names) == 2:
                        self._bump(self.dtest_links, key, rating)
                self.P = len(self.id_gene)
        except ValueError as error:
            print(error)
        except IOError as error:
            print('Error, file does not exist or can\'t be read')
            print(error)
        print('READ DATA train', len(self.links), len(self.nLinks))
        print('READ DATA train', len(self.dlinks), len(self.ndlinks))
        print('READ DATA test', len(self.test_links))
        self._engine = None

    # ---------------------------------------------------------------- device
    @staticmethod
    def _arrays(table, width):
        n = len(table)
        if n == 0:
            return np.empty((0, width), dtype=np.int32), np.empty((0, 2), dtype=np.int32)
        ids = np.array("_".join(table.keys()).split("_"), dtype=np.int64).astype(np.int32).reshape(n, width)
        cnt = np.fromiter((c for pair in table.values() for c in pair), dtype=np.int32, count=2 * n).reshape(n, 2)
        return ids, cnt

    def _ready(self):
        from .engine import EMEngine
        if self.K < 1 or self.P < 1:
            raise ValueError("initialize_parameters() and get_train_test() must run before numeric methods")
        if self._engine is None or self._engine.K != self.K or self._engine.P != self.P:
            eng = EMEngine(self.P, self.K, device=self._device, flags=self._flags)
            ids3, cnt3 = self._arrays(self.links, 3)
            ids2, cnt2 = self._arrays(self.dlinks, 2)
            eng.set_train_links(ids3[:, 0], ids3[:, 1], ids3[:, 2], cnt3[:, 0], cnt3[:, 1])
            eng.set_pair_links(ids2[:, 0], ids2[:, 1], cnt2[:, 0], cnt2[:, 1])
            deg = np.bincount(np.concatenate([ids3.ravel(), ids2.ravel()]), minlength=self.P)
            self._deg_zero = bool((deg[: self.P] == 0).any())
            self._engine = eng
            self._uploaded = False
        if not self._uploaded:
            self._engine.set_params(np.asarray(self._theta, dtype=np.float64), np.asarray(self._pr, dtype=np.float64))
            self._engine.set_q(np.asarray(self._qr, dtype=np.float64))
            self._uploaded = True
        return self._engine

    def make_iteration(self):
        eng = self._ready()
        if self._deg_zero:
            raise ZeroDivisionError("float division by zero")     # _23.py:1643, a gene that has no training link
        eng.em_iteration()
        self._on_device = True

    def compute_likelihood(self):
        eng = self._ready()
        self.likelihood = eng.loglik("train")
        return self.likelihood
