"""The C-ABI library builds, loads and exports every symbol include/tip.h declares (no compute calls)."""
import ctypes
import os

import pytest

from trigenicinteractionpredictor_b200 import _cabi, build


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _cabi.load()


def test_header_and_binding_agree(lib):
    declared = _cabi.header_functions()
    assert declared, "no functions parsed from include/tip.h"
    assert sorted(_cabi.SIGNATURES) == declared
    for name in declared:
        assert hasattr(lib, name), name


def test_trivial_host_only_calls(lib):
    assert lib.tip_abi_version() == 1
    assert lib.tip_stats_len(6000, 10) == 6000 * 10 + 2000 + 1
    assert lib.tip_rows_capacity(1000) >= 2 * 1000 + 62
    assert lib.tip_loglik_workspace_bytes(100, 10) > 0
    assert lib.tip_loglik_workspace_bytes(100, 20) >= lib.tip_loglik_workspace_bytes(100, 10) + 2 * 100 * 400 * 8
    nb = ctypes.c_size_t(0)
    assert lib.tip_em_workspace_bytes(100, 10, 3200, 0, ctypes.byref(nb)) == 0 and nb.value == 2 * 100 * 100 * 8
    assert lib.tip_em_workspace_bytes(100, 4, 3200, 0, ctypes.byref(nb)) == 0 and nb.value == 0
    assert lib.tip_em_workspace_bytes(100, 16, 3200, 0, ctypes.byref(nb)) == 0 and nb.value == 2 * 100 * 256 * 8
    assert lib.tip_em_workspace_bytes(100, 20, 3200, 0, ctypes.byref(nb)) == 0 and nb.value == 4 * 100 * 400 * 8
    assert lib.tip_em_workspace_bytes(100, 20, 3200, 1, ctypes.byref(nb)) == 0 and nb.value == 3200 * 8
    assert lib.tip_em_workspace_bytes(100, 33, 3200, 0, ctypes.byref(nb)) != 0
    assert b"bad arguments" in lib.tip_last_error()


def test_sass_is_sm100a_with_fp64_fma_and_cp_async():
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-sass", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "DFMA" in out and "LDGSTS" in out and "REDG.E.ADD.F64" in out
