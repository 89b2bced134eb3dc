#!/usr/bin/env python3
"""profiles/r2_seg3_pass_metrics.json from an `ncu --set full` capture of the two seg3_pass_kernel launches of ONE E-step
(pass A, then pass B + C); read here, no GPU.  bench.py quotes `traffic_bytes` (DRAM read + write of the two launches) and
the ncu figures beside the roofline it measures in-run.
    python tools/ncu_seg3_metrics.py gpurun_out/x.ncu-rep "<how it was captured>" > profiles/r2_seg3_pass_metrics.json"""
import collections
import csv
import json
import subprocess
import sys

rep, how = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def page(name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


raw = page("raw")
hdr, units, data = raw[0], raw[1], raw[2:]
rows = [r for r in data if "seg3_pass_kernel" in r[hdr.index("Kernel Name")]][:2]


def val(r, name, scale=False):
    i = hdr.index(name)
    v = float(r[i].replace(",", ""))
    return v * UNIT.get(units[i], 1.0) if scale else v


def maybe(name):
    return [val(r, name) for r in rows] if name in hdr else None


out = {
    "capture": how,
    "kernels": [r[hdr.index("Kernel Name")] for r in rows],
    "duration_us": maybe("gpu__time_duration.sum"),
    "traffic_bytes": sum(val(r, "dram__bytes_read.sum", True) + val(r, "dram__bytes_write.sum", True) for r in rows),
    "dram_read_bytes": [val(r, "dram__bytes_read.sum", True) for r in rows],
    "dram_write_bytes": [val(r, "dram__bytes_write.sum", True) for r in rows],
    "lts_t_bytes": [val(r, "lts__t_bytes.sum", True) for r in rows] if "lts__t_bytes.sum" in hdr else None,
    "lts_sectors_from_sm": maybe("lts__t_sectors_srcunit_tex.sum"),
    "lts_sectors_red": maybe("lts__t_sectors_op_red.sum"),
    "l1_hit_rate_pct": maybe("l1tex__t_sector_hit_rate.pct"),
    "issue_active_pct": maybe("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "fp64_pipe_pct": maybe("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    "dmma_pipe_active_pct": maybe("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"),
    "warps_active_pct": maybe("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "sm_cycles_active_over_elapsed": [val(r, "sm__cycles_active.avg") / val(r, "sm__cycles_elapsed.avg") for r in rows],
    "registers": maybe("launch__registers_per_thread"),
    "grid": maybe("launch__grid_size"),
}
src = page("source", ["--print-source", "sass"])
heads = [i for i, r in enumerate(src) if r and r[0] == "Address"]
mix = []
for hi in heads[:2]:
    sh = src[hi]
    st = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    seen = set()
    for r in src[hi + 1:]:
        if len(r) != len(sh) or r[0] in seen:
            if len(r) == len(sh):
                break
            continue
        seen.add(r[0])
        for h in st:
            try:
                tot[h[6:]] += float(r[sh.index(h)])
            except ValueError:
                pass
    s = sum(tot.values()) or 1.0
    mix.append({k: round(100 * v / s, 1) for k, v in tot.most_common(6)})
out["stall_mix"] = mix
print(json.dumps(out, indent=1))
