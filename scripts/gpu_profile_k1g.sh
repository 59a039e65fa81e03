#!/bin/bash
# Run on the GPU box: one full ncu capture of the per-thread FAST kernel with a gravity field (column walk) on the C2 geometry.
TAG=${1:-r01_k1g}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:nyxb_k_thread_fast -c 1 -o gpurun_out/${TAG}_thread \
    python bench.py --lanes 1 --n-traj ${NTRAJ:-100000} --span-days 0.05 --steps 1 --warmup 0 --no-cpu-baseline --no-strict > gpurun_out/${TAG}_thread_bench.log 2>&1
ls -la gpurun_out/${TAG}_*
