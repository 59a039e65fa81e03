#!/bin/bash
# Run on the GPU box: one full ncu capture of the warp-cooperative filter kernel on a short C5 arc (the 1000-filter launch, not the
# 64-filter warm-up: -s 1),
# plus the pipe / DRAM metrics of the same launch.
TAG=${1:-r01_od}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:nyxb_k_od_coop -s 1 -c 1 -o gpurun_out/${TAG}_coop \
    python bench.py --workload c5 --span-days 0.01 --no-cpu-baseline > gpurun_out/${TAG}_coop_bench.log 2>&1
ls -la gpurun_out/${TAG}_*
