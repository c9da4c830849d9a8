#!/usr/bin/env python
"""Summarise .ncu-rep captures (ncu --set full) into the text committed under profiles/.

    python tests/ncu_summary.py gpurun_out/prof_x.ncu-rep [more.ncu-rep ...]  > profiles/rNN_ncu_<what>.txt
Runs on the CPU box (ncu -i ... --page raw --csv).  One block per captured launch."""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__icc_request_hit_rate.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]

for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        print(f"# {path}: no data"); continue
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]] if "Kernel Name" in idx else "?"
        print(f"# {path}\nkernel: {name[:110]}")
        for k in KEYS:
            if k in idx:
                print(f"  {k:86s} {units[idx[k]]:14s} {r[idx[k]]}")
        try:
            rd = float(r[idx["dram__bytes_read.sum"]]); wr = float(r[idx["dram__bytes_write.sum"]])
            ur, uw = units[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_write.sum"]]
            sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            print(f"  {'dram traffic per launch (read + write)':86s} {'Mbyte':14s} {(rd * sc[ur] + wr * sc[uw]) / 1e6:.3f}")
        except Exception:
            pass
        print()
