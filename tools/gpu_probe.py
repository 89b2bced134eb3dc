#!/usr/bin/env python3
"""Measure the roofline denominators on the box: DFMA / FFMA / DMMA peaks and fp64 RED throughput."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trigenicinteractionpredictor_b200 import _cabi  # noqa: E402

lib = _cabi.load()
out = {}
v = ctypes.c_double(0)
for kind, name in [(0, "dfma_tflops"), (1, "ffma_tflops"), (2, "dmma_tflops"), (3, "dfma_plus_dmma_tflops")]:
    rc = lib.tip_measure_fma_peak(kind, ctypes.byref(v))
    out[name] = v.value if rc == 0 else ("error: %s" % lib.tip_last_error())
for n_addr in (60000, 6000000):
    for mode, nm in [(0, "scattered"), (1, "row10")]:
        rc = lib.tip_measure_red_f64(n_addr, mode, ctypes.byref(v))
        out["red_f64_%s_%d_gops" % (nm, n_addr)] = v.value if rc == 0 else ("error: %s" % lib.tip_last_error())
print(json.dumps(out, indent=1))
