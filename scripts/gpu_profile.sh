#!/bin/bash
# Run on the GPU box: launch list (gpu__time_duration of every kernel of the bench command) + one full capture of the
# dominant kernel.  Outputs land in gpurun_out/; summaries are copied into profiles/ by hand (scripts/ncu_summary.py).
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nyxb_k_coop -c 1 -o gpurun_out/${TAG}_coop \
    python bench.py --steps 1 --warmup 0 --span-days 0.05 --no-cpu-baseline > gpurun_out/${TAG}_coop_bench.log 2>&1
ls -la gpurun_out/${TAG}_*
