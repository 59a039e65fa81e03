#!/bin/bash
# ncu evidence for the transposed kernel (run on the GPU box):  bash scripts/gpu_profile_tx.sh r02c
#  (1) launch list of the default bench command, (2) one full capture on a short span with source counters,
#  (3) pipe / traffic counters of one full-span launch of the default workload, with and without the L2 flush in front of it.
TAG=${1:-r02c}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-strict > gpurun_out/${TAG}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nyxb_k_tx -c 1 -o gpurun_out/${TAG}_tx \
    python bench.py --steps 1 --warmup 0 --span-days 0.05 --no-cpu-baseline --no-strict --kernel transposed > gpurun_out/${TAG}_tx_bench.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active
ncu --clock-control none -k regex:nyxb_k_tx -c 1 --metrics $M --csv --log-file gpurun_out/${TAG}_fullspan.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-strict --kernel transposed > gpurun_out/${TAG}_fullspan_bench.log 2>&1
ls -la gpurun_out/${TAG}_*
