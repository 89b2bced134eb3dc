"""Host-side logic of the drop-in Model (no GPU): digestion, fold, init, CLI parsing - bit-exact
against the vectors the reference produced (tests/golden, see oracle/gen_golden.py)."""
import json
import os
import random

import numpy as np
import pytest

from tests.conftest import GOLDEN
from trigenicinteractionpredictor_b200 import Model, main
from trigenicinteractionpredictor_b200 import dist as tdist

BASE = os.path.join(GOLDEN, "base")
DUPS = os.path.join(GOLDEN, "dups")


def _load(case, train, test):
    m = Model()
    m.get_traintest(os.path.join(case, train), os.path.join(case, test))
    return m


def _check(m, rec):
    assert m.P == rec["P"]
    assert [m.id_gene[i] for i in range(m.P)] == rec["genes_in_id_order"]
    assert [m.uniqueg[i] for i in range(m.P)] == rec["uniqueg"]
    assert list(m.links.keys()) == rec["link_keys"]
    assert [list(v) for v in m.links.values()] == rec["link_counts"]
    assert list(m.nlinks.keys()) == rec["nlink_keys"]
    assert list(m.test_links.keys()) == rec["test_keys"]
    assert [list(v) for v in m.test_links.values()] == rec["test_counts"]
    assert m.gene_id == {v: k for k, v in m.id_gene.items()}


def test_get_traintest_bit_exact(capsys):
    m = _load(BASE, "train1.dat", "test1.dat")
    _check(m, json.load(open(os.path.join(BASE, "digest.json"))))
    out = capsys.readouterr().out
    assert "READ DATA train 1600 1600" in out and "READ DATA test 400" in out


def test_get_traintest_dups():
    _check(_load(DUPS, "train.dat", "test.dat"), json.load(open(os.path.join(DUPS, "digest.json"))))


def test_get_input_and_fold_bit_exact(tmp_path, monkeypatch):
    man = json.load(open(os.path.join(GOLDEN, "manifest.json")))["base"]
    m = Model()
    m.get_input(os.path.join(BASE, "input_s2.tsv"))
    assert [m.id_gene[i] for i in range(m.P)] == man["get_input"]["genes_in_id_order"]
    assert list(m.links)[:50] == man["get_input"]["link_keys_head"]
    assert len(m.links) == man["get_input"]["n_links"]
    assert sum(1 for v in m.links.values() if v[1]) == man["get_input"]["n_pos"]
    monkeypatch.chdir(tmp_path)
    np.random.seed(2)
    m.fold()
    for i in range(5):
        for kind in ("test", "train"):
            name = "%s%d.dat" % (kind, i)
            assert open(name, "rb").read() == open(os.path.join(BASE, name), "rb").read(), name


def test_fold_remainder_goes_to_last_fold(tmp_path, monkeypatch):
    m = Model()
    for i in range(13):
        m.gene_id["g%d" % i] = i
        m.id_gene[i] = "g%d" % i
    for i in range(11):
        m.links["%d_%d_%d" % (i, i + 1, i + 2)] = [1, 0] if i % 3 else [0, 1]
    monkeypatch.chdir(tmp_path)
    np.random.seed(0)
    m.fold(0.2)
    sizes = [len(open("test%d.dat" % i).readlines()) for i in range(5)]
    assert sizes == [2, 2, 2, 2, 3]
    assert [len(open("train%d.dat" % i).readlines()) for i in range(5)] == [9, 9, 9, 9, 8]


@pytest.mark.parametrize("K", [1, 2, 3, 10])
def test_initialize_parameters_bit_exact(K):
    m = _load(BASE, "train1.dat", "test1.dat")
    tr = np.load(os.path.join(BASE, "trace_K%d.npz" % K))
    random.seed(1000)
    m.initialize_parameters(K)
    assert np.array_equal(np.array(m.theta), tr["theta0"])
    assert np.array_equal(np.array(m.pr), tr["pr0"])
    assert random.random() == tr["after_init_random"][0]
    assert m.K == K and m.vlikelihood == []
    assert np.array(m.ntheta).shape == (m.P, K) and not np.array(m.npr).any()


def test_initialize_parameters_bad_k_defaults_to_10():
    m = Model()
    m.P = 3
    m.initialize_parameters("not-a-number")
    assert m.K == 10 and len(m.theta[0]) == 10


def test_numeric_methods_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from trigenicinteractionpredictor_b200._cabi import TipLibraryError
    m = _load(BASE, "train1.dat", "test1.dat")
    m.initialize_parameters(2)
    with pytest.raises(TipLibraryError):
        m.make_iteration()
    with pytest.raises(TipLibraryError):
        m.compute_likelihood()


def test_cli_rejects_bad_arguments(tmp_path, capsys):
    assert main(["--bogus"]) == 2
    assert main(["-k", "0", "-t", os.path.join(BASE, "train1.dat"), "-e", os.path.join(BASE, "test1.dat")]) == 2
    assert main(["-t", "/nonexistent", "-e", os.path.join(BASE, "test1.dat")]) == 2
    assert main(["-o", str(tmp_path / "missing"), "-t", os.path.join(BASE, "train1.dat")]) == 2
    assert main(["-h"]) == 0
    assert main(["--mode", "tf32", "-t", os.path.join(BASE, "train1.dat"), "-e", os.path.join(BASE, "test1.dat")]) == 2


def test_shard_bounds_and_sample_assignment():
    for n, w in [(10, 3), (100_000_000, 8), (5, 8), (0, 2)]:
        spans = [tdist.shard_bounds(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    per = [len(tdist.samples_for_rank(0, 50, r, 8)) for r in range(8)]
    assert per == [7, 7, 6, 6, 6, 6, 6, 6]
    assert sorted(s for r in range(8) for s in tdist.samples_for_rank(10, 50, r, 8)) == list(range(10, 60))


@pytest.mark.parametrize("K", [1, 2, 3, 10, 17, 32])
def test_block_drawn_initialisation_equals_the_reference_loops(K):
    """initialize_parameters draws all P*K + 2K^3 numbers in one block (numpy RandomState on the random module's MT19937
    state) and restates builtin sum(): values, and the position of the `random` stream afterwards, must be bit-identical
    to one random.random() per cell in the reference's order (TIP.py:117-170)."""
    import random
    from trigenicinteractionpredictor_b200.TrigenicInteractionPredictor import Model
    m = Model()
    m.P = 613
    random.seed(4242 + K)
    th1, p1 = m._draw_parameters_loops(K, 2)
    after_loops, next_loops = random.getstate(), random.random()
    random.seed(4242 + K)
    th2, p2 = m._draw_parameters_fast(K, 2)
    assert th2 is not None
    after_fast, next_fast = random.getstate(), random.random()
    assert th2.tolist() == th1 and p2.tolist() == p1
    assert after_fast == after_loops and next_fast == next_loops
    random.seed(4242 + K)
    m.initialize_parameters(K)
    assert m.theta == th1 and m.pr == p1 and isinstance(m.theta, list) and isinstance(m.pr[0][0][0], list)


def test_vectorised_soa_equals_the_per_link_loop():
    """Model._soa parses all keys in one pass; it must give what splitting every key in a Python loop gives."""
    from trigenicinteractionpredictor_b200.TrigenicInteractionPredictor import Model
    rng = np.random.default_rng(4)
    table = {}
    for _ in range(5000):
        a, b, c = rng.integers(0, 1200, 3).tolist()
        table.setdefault("%d_%d_%d" % (a, b, c), [0, 0])[int(rng.integers(0, 2))] += 1
    g1, g2, g3, n0, n1 = Model._soa(table)
    assert all(x.dtype == np.int32 for x in (g1, g2, g3, n0, n1))
    for i, (key, c) in enumerate(table.items()):
        assert [int(g1[i]), int(g2[i]), int(g3[i])] == [int(t) for t in key.split("_")] and [int(n0[i]), int(n1[i])] == c
    assert all(len(x) == 0 for x in Model._soa({}))


def test_links_handed_over_as_arrays_behave_like_the_dict():
    """Model.set_links_soa (how configs 3 and 4 enter the drop-in): links / test_links keep len, key order, values."""
    from trigenicinteractionpredictor_b200.TrigenicInteractionPredictor import Model, SoALinks
    g = np.array([[10, 11, 9], [3, 7, 5], [0, 2, 1]], dtype=np.int32)
    n0, n1 = np.array([1, 0, 2], dtype=np.int32), np.array([0, 1, 1], dtype=np.int32)
    m = Model()
    m.set_links_soa((g[:, 0], g[:, 1], g[:, 2], n0, n1), (g[:1, 0], g[:1, 1], g[:1, 2], n0[:1], n1[:1]))
    assert m.P == 12 and len(m.links) == 3 and len(m.test_links) == 1 and isinstance(m.links, SoALinks)
    assert list(m.links.keys()) == ["10_11_9", "3_7_5", "0_2_1"] and list(m.links) == list(m.links.keys())
    assert list(m.links.values()) == [[1, 0], [0, 1], [2, 1]]
    assert dict(m.links.items())["3_7_5"] == [0, 1] and m.links.count_single_positive() == 2
    got = Model._soa(m.links)
    assert all(np.array_equal(a, b) for a, b in zip(got, (g[:, 0], g[:, 1], g[:, 2], n0, n1)))


def test_product_never_touches_the_oracle_or_the_reference():
    """The oracle is test infrastructure: nothing under the package (nor the GPU arm of bench.py) may import it,
    read /root/reference, or carry a CPU fallback for the numerics."""
    import ast
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "trigenicinteractionpredictor_b200")
    for name in sorted(os.listdir(pkg)):
        if not name.endswith(".py"):
            continue
        src = open(os.path.join(pkg, name), encoding="utf-8").read()
        for node in ast.walk(ast.parse(src)):
            mods = []
            if isinstance(node, ast.Import):
                mods = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                mods = [node.module or ""]
            assert not any(m == "oracle" or m.startswith("oracle.") for m in mods), name
        assert "/root/reference" not in src, name
    for name in sorted(os.listdir(os.path.join(pkg, "csrc"))):
        if name.endswith((".cu", ".cuh")):
            assert "oracle" not in open(os.path.join(pkg, "csrc", name), encoding="utf-8").read().lower(), name
    bench = open(os.path.join(root, "bench.py"), encoding="utf-8").read()
    ours = bench[bench.index("def run_ours("):bench.index("def _measure_peak(") if "def _measure_peak(" in bench else len(bench)]
    # inside the GPU arm the oracle appears only in the cpu_baseline leg (functions defined above run_ours)
    assert not re.search(r"^\s*(from|import)\s+oracle", ours, flags=re.M)
    assert "/root/reference" not in bench


@pytest.mark.parametrize("seed", range(8))
def test_random_files_digest_and_fold_like_the_oracle(tmp_path, monkeypatch, seed):
    """Random train/test files with everything the digestion has to get right - unsorted names, duplicated lines,
    conflicting labels, names whose order differs from the order of their ids, ids with different digit counts
    (the decimal-STRING sort of TIP.py:353), genes seen only in the test file, several tabs before a test label -
    give the same dictionaries, in the same order, as the oracle (which is pinned to the reference), and the same
    fold files from the same numpy seed."""
    from oracle import mmsbm_oracle as orc
    rng = np.random.default_rng(1000 + seed)
    P = int(rng.integers(12, 140))
    names = ["G%d" % v for v in rng.permutation(5000)[:P]]                  # digit counts 1..4, id order != name order
    def triple():
        a, b, c = rng.choice(P, size=3, replace=False).tolist()
        t = [names[a], names[b], names[c]]
        rng.shuffle(t)
        return "_".join(t)
    train = [triple() + "\t" + str(int(rng.random() < 0.3)) + "\n" for _ in range(int(rng.integers(30, 400)))]
    train += [train[i] for i in rng.integers(0, len(train), size=len(train) // 5)]                      # duplicates
    train += [train[i].rsplit("\t", 1)[0] + "\t" + ("0" if train[i].strip().endswith("1") else "1") + "\n"
              for i in rng.integers(0, len(train), size=len(train) // 7)]                               # conflicts
    extra = ["X%d" % i for i in range(3)]                                                               # test-only genes
    test = [triple() + ("\t" * int(rng.integers(1, 4))) + str(int(rng.random() < 0.3)) + "\n"
            for _ in range(int(rng.integers(5, 80)))]
    test.append("_".join([extra[0], extra[1], names[0]]) + "\t1\n")
    test.append("_".join([extra[2], names[1], names[2]]) + "\t0\n")
    tr, te = tmp_path / "train.dat", tmp_path / "test.dat"
    tr.write_text("".join(train))
    te.write_text("".join(test))
    m = Model()
    m.get_traintest(str(tr), str(te))
    dg = orc.digest_traintest(train, test)
    assert m.P == dg.P and m.gene_id == dg.gene_id and m.id_gene == dg.id_gene and m.uniqueg == dg.uniqueg
    for mine, ref in ((m.links, dg.links), (m.nlinks, dg.nlinks), (m.test_links, dg.test_links)):
        assert list(mine.items()) == list(ref.items())                      # same keys, same order, same counts
    np.random.seed(77 + seed)
    test_txt, train_txt = orc.fold_texts(dg.links, dg.id_gene)
    monkeypatch.chdir(tmp_path)
    np.random.seed(77 + seed)
    m.fold()
    for i in range(5):
        assert (tmp_path / ("test%d.dat" % i)).read_text() == test_txt[i]
        assert (tmp_path / ("train%d.dat" % i)).read_text() == train_txt[i]
