#!/bin/bash
# Run on the GPU box.  (1) launch list of the bench command (gpu__time_duration of every kernel), (2) one full capture of the
# dominant kernel on a short span (source-level counters), (3) one full capture on the DEFAULT workload for the DRAM traffic.
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nyxb_k_coop -c 1 -o gpurun_out/${TAG}_coop \
    python bench.py --steps 1 --warmup 0 --span-days 0.05 --no-cpu-baseline > gpurun_out/${TAG}_coop_bench.log 2>&1
ncu --clock-control none -k regex:nyxb_k_coop -c 1 --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active \
    --csv --log-file gpurun_out/${TAG}_fullspan.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${TAG}_fullspan_bench.log 2>&1
ls -la gpurun_out/${TAG}_*
