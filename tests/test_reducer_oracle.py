"""The reducer oracle (oracle/reducer_oracle.py) against the files the unmodified reference script
testResultsReducer.py wrote for tests/golden/reducer/ (oracle/gen_golden_reducer.py)."""
import os

import pytest

from oracle import reducer_oracle as ro
from tests.conftest import GOLDEN

RED = os.path.join(GOLDEN, "reducer")


def test_oracle_reproduces_reference_files_byte_for_byte():
    out = ro.reduce_folder(os.path.join(RED, "results") + "/", os.path.join(RED, "DATA_FOLDS"))
    expected = sorted(os.listdir(os.path.join(RED, "expected")))
    assert sorted("K%d_fold%d.csv" % c for c in out) == expected and len(expected) == 20
    for (K, fold), text in out.items():
        with open(os.path.join(RED, "expected", "K%d_fold%d.csv" % (K, fold))) as fh:
            assert fh.read() == text, (K, fold)


@pytest.mark.parametrize("n,idx", [(1, 0), (3, 2), (5, 2), (7, 4), (9, 4), (11, 6)])
def test_median_follows_python_round_half_even(n, idx):
    vals = [float(v) for v in range(n, 0, -1)]
    mean, median, std = ro.reduce_values(vals)
    assert median == sorted(vals)[idx]
    assert mean == sum(range(1, n + 1)) / n


def test_even_median_and_std_over_sorted_values():
    mean, median, std = ro.reduce_values([0.4, 0.1, 0.3, 0.2])
    assert median == (0.2 + 0.3) / 2
    acc = 0
    for v in (0.1, 0.2, 0.3, 0.4):
        acc += (v - mean) ** 2
    assert std == (acc / 4) ** 0.5
