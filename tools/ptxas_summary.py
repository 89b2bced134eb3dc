#!/usr/bin/env python3
"""One-file register / spill summary of the last build (the per-object ptxas logs are build artefacts, git-ignored).
    python tools/ptxas_summary.py > profiles/rN_ptxas_summary.txt"""
import glob
import os
import re
import subprocess

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = []
for log in sorted(glob.glob(os.path.join(HERE, "trigenicinteractionpredictor_b200", "csrc", "_obj", "*.ptxas.log"))):
    text = open(log).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?Function properties.*?\n\s*(\d+) bytes stack frame, "
                         r"(\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", text, flags=re.S):
        rows.append((os.path.basename(log).replace(".ptxas.log", ""), m.group(1), int(m.group(5)), int(m.group(3)), int(m.group(4))))
names = subprocess.run(["c++filt"] + [r[1] for r in rows], capture_output=True, text=True).stdout.splitlines() if rows else []
print("%-12s %5s %7s %7s  kernel" % ("object", "regs", "spill_st", "spill_ld"))
for (obj, _, regs, st, ld), name in zip(rows, names):
    name = re.sub(r"\(.*", "", name)
    # every kernel that spills, and the K = 10 / 16 / 32 instances of the E-step families
    if st or ld or re.search(r"(seg3_pass_kernel|seg3_finish_kernel|em_fused_kernel|em_finalize_kernel|loglik_\w+)<(10|16|32)[,>]", name):
        print("%-12s %5d %7d %7d  %s" % (obj, regs, st, ld, name[:110]))
print("\n%d kernels in total; listed: every kernel with spills, and the K = 10 / 16 / 32 instances of the E-step families" % len(rows))
