"""Summarise an .ncu-rep (read here on the CPU box): python scripts/ncu_summary.py gpurun_out/x.ncu-rep [pattern...]"""
import csv, subprocess, sys
rep = sys.argv[1]
pats = sys.argv[2:] or ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak", "achieved_occupancy", "smsp__inst_executed.sum", "pipe_fp64",
    "smsp__issue_active.avg.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "thread_inst_executed_per_inst",
    "warp_issue_stalled", "sm__throughput.avg.pct", "l1tex__t_sector_hit_rate", "lts__t_sector_hit_rate", "sm__cycles_elapsed.max",
    "smsp__cycles_active.avg", "sm__inst_executed_pipe_", "l1tex__data_bank_conflicts", "launch__shared_mem", "dram__throughput"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("=== kernel:", r[hdr.index("Kernel Name")][:80])
    for i, h in enumerate(hdr):
        if any(p in h for p in pats) and r[i] not in ("", "n/a"):
            print(f"  {h:95s} {units[i]:12s} {r[i]}")
