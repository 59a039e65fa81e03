#!/bin/bash
# Round 2: pipe / DRAM counters of one full-span launch of the transposed kernel on C4 (GRAIL 70x70, 16 walker positions)
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 170 ncu --clock-control none -k regex:nyxb_k_tx -c 1 --metrics $M --csv --log-file gpurun_out/r02c4_fullspan.csv \
    python bench.py --workload c4 --steps 1 --warmup 0 --no-cpu-baseline --no-strict > gpurun_out/r02c4_fullspan_bench.log 2>&1
grep -E "pipe_fp64|issue_active|lsu_wavefronts|time_duration|dram__bytes|warps_active" gpurun_out/r02c4_fullspan.csv | awk -F'","' '{print $(NF-2), $(NF-1), $(NF)}'
