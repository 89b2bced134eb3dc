"""Cross-sample reducer (trigenicinteractionpredictor_b200/testResultsReducer.py): host logic on the CPU, the
device reduction against the reference's own output files and against the oracle on the GPU."""
import os

import numpy as np
import pytest

from tests.conftest import GOLDEN

RED = os.path.join(GOLDEN, "reducer")


def test_cli_errors_like_the_reference(capsys):
    from trigenicinteractionpredictor_b200 import testResultsReducer as trr
    assert trr.main(["-f", "/nonexistent/folder/"]) == 2           # TRR.py:33-41
    assert "does not exist" in capsys.readouterr().out
    assert trr.main(["--bogus"]) == 2


def test_gene_list_block_layout():
    from trigenicinteractionpredictor_b200 import testResultsReducer as trr

    class M:
        id_gene = {0: "YAL001C", 1: "YBR002W"}
        uniqueg = {0: 3, 1: 5}
    assert trr.gene_list_block(M()) == ("\nLIST OF REGISTERED GENES\nGene_ID\tGene_name\tnumAparitions\n"
                                        "0\tYAL001C\t3\n1\tYBR002W\t5\n\nLIST OF LINKS BETWEEN GENE IDS\n")


def test_parsing_matches_fixture_and_no_cpu_path():
    """The host parser reads the reference-format sample files; without a CUDA device the reduction refuses to run."""
    import torch
    from trigenicinteractionpredictor_b200 import _cabi, testResultsReducer as trr
    path = os.path.join(RED, "results", "K3", "fold1", "Sample_0_K3.csv")
    names = trr._read_gene_names(path)
    assert len(names) == 30 and all(n.startswith("G000") for n in names)
    heldout, rows = trr._read_sample(path)
    assert len(rows) == 50 and heldout < 0 and rows == sorted(rows, reverse=True)
    assert 0.0 < trr._training_density(os.path.join(RED, "DATA_FOLDS"), 1) < 0.5
    if not torch.cuda.is_available():
        with pytest.raises(_cabi.TipLibraryError):
            trr.reduce_results(os.path.join(RED, "results") + "/", os.path.join(RED, "DATA_FOLDS"), log=lambda *_: None)


@pytest.fixture
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _parse_cell(text):
    lines = text.split("\n")
    assert lines[0] == "" and lines[1] == "Held-OutLikelihoodMean\tAUCmean\tPrecision\tRecall\tFallout"
    assert lines[3] == "" and lines[4] == "TripleteName\tMean\tMedian\tStdDev\tRealinteraction"
    rows = [l.split("\t") for l in lines[5:] if l]
    return lines[2], rows


@pytest.mark.gpu
def test_reducer_reproduces_reference_files(torch_cuda, tmp_path):
    from trigenicinteractionpredictor_b200 import testResultsReducer as trr
    out = trr.reduce_results(os.path.join(RED, "results") + "/", os.path.join(RED, "DATA_FOLDS"), str(tmp_path),
                             log=lambda *_: None)
    expected = sorted(os.listdir(os.path.join(RED, "expected")))
    assert sorted(os.listdir(tmp_path)) == expected and len(out) == 20
    for name in expected:
        with open(os.path.join(RED, "expected", name)) as fh:
            head_e, rows_e = _parse_cell(fh.read())
        with open(os.path.join(tmp_path, name)) as fh:
            head_g, rows_g = _parse_cell(fh.read())
        assert head_g == head_e, name                       # likelihood mean, AUC, precision, recall, fallout: exact text
        assert len(rows_g) == len(rows_e)
        for g, e in zip(rows_g, rows_e):
            assert g[0] == e[0] and g[1] == e[1] and g[2] == e[2] and g[4] == e[4], (name, g, e)   # order, mean, median
            # the reference squares with `** 2` (libm pow), the kernel with a product: at most an ulp or two apart
            assert float(g[3]) == pytest.approx(float(e[3]), rel=1e-13, abs=1e-300), (name, g, e)


@pytest.mark.gpu
@pytest.mark.parametrize("T,S", [(4000, 13), (1500, 64), (300, 200)])     # 200 samples: the global-scratch sort path
def test_reduce_samples_kernel_vs_oracle_ragged(torch_cuda, T, S):
    from oracle import reducer_oracle as ro
    from trigenicinteractionpredictor_b200 import testResultsReducer as trr
    rng = np.random.default_rng(9)
    cols, labs = [], []
    for t in range(T):
        n = 1 + (t % S)
        v = rng.random(n)
        if t % 7 == 0:
            v = np.round(v, 1)                               # exact ties inside a column
        cols.append(v.tolist())
        labs.append(int(rng.random() < 0.2))
    mean, median, std, c = trr.reduce_cell_on_device(cols, labs, 0.2)
    exp = [ro.reduce_values(list(col)) for col in cols]
    assert mean.tolist() == [e[0] for e in exp]              # bit-identical
    assert median.tolist() == [e[1] for e in exp]
    np.testing.assert_allclose(std, [e[2] for e in exp], rtol=1e-13, atol=0)
    recs = [[str(t), exp[t][0], exp[t][1], exp[t][2], labs[t]] for t in range(T)]
    recs.sort(key=lambda r: r[1], reverse=True)
    auc, precision, recall, fallout = ro.metrics_of_cell(recs, 0.2)
    assert c["wins"] / (c["n_pos"] * c["n_neg"]) == auc
    assert [c["tp"] / (c["tp"] + c["fp"]), c["tp"] / (c["tp"] + c["fn"]), c["fp"] / (c["fp"] + c["tn"])] == [precision, recall, fallout]
