#!/usr/bin/env python
"""profiles/r02_traffic.json from an ncu --csv metrics log of ONE full-span launch of the default bench workload
(scripts/gpu_profile_tx.sh step 3):  python scripts/make_traffic_json.py gpurun_out/r02d_fullspan.csv [n_traj]"""
import csv, json, subprocess, sys
from pathlib import Path

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
m = {r[-3]: float(r[-1].replace(",", "")) for r in rows}
kernel = rows[0][4]
head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
out = {"workload": "c2", "n_traj": int(sys.argv[2]) if len(sys.argv) > 2 else 10000, "span_days": 3.0, "degree": 21,
       "kernel": kernel.split("(")[0], "dram_bytes_read": m["dram__bytes_read.sum"], "dram_bytes_write": m["dram__bytes_write.sum"],
       "dram_bytes_per_launch": m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"], "kernel_ms": m["gpu__time_duration.sum"] / 1e6,
       "fp64_pipe_pct": m.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
       "issue_active_pct": m.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
       "lsu_wavefronts_pct": m.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
       "source": sys.argv[1], "git_head_at_summary": head,
       "note": "one launch under ncu (--clock-control none), preceded by the bench's 256 MiB L2 flush"}
Path("profiles/r02_traffic.json").write_text(json.dumps(out, indent=1) + "\n")
print(json.dumps(out, indent=1))
