#!/bin/bash
# Run on the GPU box: one full ncu capture of the STRICT lane-cooperative kernel on a short C2 span.
TAG=${1:-r01_k3}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:nyxb_k_coop_strict -c 1 -o gpurun_out/${TAG}_strict \
    python bench.py --mode strict --steps 1 --warmup 0 --span-days 0.05 --no-cpu-baseline > gpurun_out/${TAG}_strict_bench.log 2>&1
ls -la gpurun_out/${TAG}_*
