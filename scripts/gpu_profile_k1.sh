#!/bin/bash
# Run on the GPU box: one full ncu capture of the per-thread FAST kernel on the C3 geometry (100 000 JWST-like trajectories).
TAG=${1:-r01_k1}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:nyxb_k_thread_fast -c 1 -o gpurun_out/${TAG}_thread \
    python bench.py --workload c3 --n-traj 100000 --span-days 2 --steps 1 --warmup 0 --no-cpu-baseline --no-strict > gpurun_out/${TAG}_thread_bench.log 2>&1
ls -la gpurun_out/${TAG}_*
