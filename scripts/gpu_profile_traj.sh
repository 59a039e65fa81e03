#!/bin/bash
# Run on the GPU box: one full ncu capture each of the resampling kernel and of the event-location kernel (nyxb_traj.cu),
# on a short C2 recording (bench.py --record) and on the event tests.  Summaries: scripts/ncu_summary.py on the .ncu-rep files.
TAG=${1:-r02_traj}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:nyxb_k_traj_resample -c 1 -o gpurun_out/${TAG}_resample \
    python bench.py --steps 1 --warmup 1 --span-days 0.25 --record 512 --no-cpu-baseline --no-strict > gpurun_out/${TAG}_resample_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nyxb_k_event_locate -c 1 -o gpurun_out/${TAG}_locate \
    python -m pytest tests/test_events.py -q -m gpu -k "monte_carlo_api" -p no:cacheprovider > gpurun_out/${TAG}_locate_pytest.log 2>&1
ls -la gpurun_out/${TAG}_*
