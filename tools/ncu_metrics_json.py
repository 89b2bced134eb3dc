#!/usr/bin/env python3
"""Selected metrics of the first kernel in an .ncu-rep as JSON ({metric: [value, unit]}); read here, no GPU.
    python tools/ncu_metrics_json.py gpurun_out/x.ncu-rep > profiles/x_metrics.json
bench.py reads profiles/r1_em_fused_k10_metrics.json (dram bytes per launch, fp64 pipe utilisation)."""
import csv
import json
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sectors_op_red.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
        "sm__sass_thread_inst_executed_op_dfma_pred_on.sum", "idc__request_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, first = rows[0], rows[1], rows[2]
print(json.dumps({w: [first[hdr.index(w)], units[hdr.index(w)]] for w in WANT if w in hdr}, indent=1))
