#!/bin/bash
# ncu full capture (source counters) of the warp-specialised transposed kernel: 9 472 trajectories (every context holds a set from start
# to end, no parking, no tail), short span.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import nyx_b200.abi as a; a.load_library()" || { echo "libnyxb.so missing or stale"; exit 9; }
timeout 200 ncu --set full --clock-control none --import-source on -k regex:nyxb_k_tx -c 1 -o gpurun_out/r02g_tx \
    python bench.py --steps 1 --warmup 0 --span-days 0.1 --n-traj 9472 --no-cpu-baseline --no-strict --kernel transposed > gpurun_out/r02g_tx_bench.log 2>&1
ls -la gpurun_out/r02g_tx.ncu-rep
timeout 60 python -m pytest tests/test_gpu_tx.py -x -q -p no:cacheprovider -k "fixed_step or ragged" 2>&1 | tail -2
