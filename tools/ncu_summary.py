#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU): headline metrics, per-opcode counts, stall mix, hottest SASS lines."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 0
kname = sys.argv[3] if len(sys.argv) > 3 else None     # substring of the kernel whose SASS table is printed


def page(name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


raw = page("raw")
hdr, units, data = raw[0], raw[1], raw[2:]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "launch__registers_per_thread",
        "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_op_red.sum", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_pipe_uniform.sum",
        "idc__request_cycles_active.avg.pct_of_peak_sustained_active", "smsp__pcsamp_sample_buffers.sum",
        "lts__t_sectors_srcunit_tex.sum", "lts__t_sectors_srcunit_tex.sum.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "derived__lts__lts2xbar_bytes.sum.per_second",
        "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed"]
if "Kernel Name" in hdr:
    print("kernels:", [r[hdr.index("Kernel Name")][:48] for r in data])
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print("%-70s %-12s %s" % (w, units[i], [r[i] for r in data]))
for h in hdr:
    if ("idc" in h or "imc" in h) and "pct" in h:
        i = hdr.index(h)
        print("%-70s %-12s %s" % (h, units[i], [r[i] for r in data]))

src = page("source", ["--print-source", "sass"])
his = [i for i, r in enumerate(src) if r and r[0] == "Address"]
hi = his[0]
if kname:
    for i in his:
        if i > 0 and len(src[i - 1]) > 1 and kname in src[i - 1][1]:
            hi = i
            break
print("SASS table of:", src[hi - 1][1][:80] if hi > 0 and len(src[hi - 1]) > 1 else "?")
sh = src[hi]
col = {h: i for i, h in enumerate(sh)}
rows, seen = [], set()
for r in src[hi + 1:]:
    if len(r) != len(sh):
        continue
    if r[0] in seen:
        break
    seen.add(r[0])
    rows.append(r)


def f(r, h):
    try:
        return float(r[col[h]])
    except Exception:
        return 0.0


byop = collections.defaultdict(lambda: [0, 0, 0, 0])
for r in rows:
    toks = r[col["Source"]].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = ".".join(op.split(".")[:2])
    b = byop[op]
    b[0] += f(r, "Instructions Executed")
    b[1] += f(r, "L1 Wavefronts Shared")
    b[2] += f(r, "# Samples")
    b[3] += 1
print("\n%-22s %14s %14s %9s %7s" % ("opcode", "warp-instr", "smem wavefr", "samples", "static"))
tot_inst = sum(b[0] for b in byop.values())
for op, b in sorted(byop.items(), key=lambda kv: -kv[1][0])[:22]:
    print("%-22s %14.0f %14.0f %9.0f %7d" % (op, b[0], b[1], b[2], b[3]))
print("total warp-instr %.0f" % tot_inst)
st = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: sum(f(r, h) for r in rows) for h in st}
s = sum(tot.values()) or 1
print("\nstall mix (all samples):", ", ".join("%s %.1f%%" % (h[6:], 100 * v / s) for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:9]))
if top:
    print("\nhottest SASS lines:")
    for r in sorted(rows, key=lambda r: -f(r, "# Samples"))[:top]:
        reasons = sorted(((f(r, h), h[6:]) for h in st), reverse=True)[:3]
        print("%6.0f  %-60s %s" % (f(r, "# Samples"), r[col["Source"]][:60], " ".join("%s=%d" % (n, v) for v, n in reasons if v)))
