#!/bin/bash
# Round 2, GPU call X: finer timeline of the stretch between two attempts (trace build only).
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T=${TAG:-r02x}
touch nyx_b200/csrc/nyxb_tx.cu nyx_b200/csrc/nyxb_api.cu
timeout 400 make -C nyx_b200/csrc EXTRA=-DNYXB_TX_TRACE > gpurun_out/${T}_make.log 2>&1; echo "make rc=$?"
NYXB_TX_TRACE_FILE=gpurun_out/${T}_trace.bin timeout 120 python bench.py --steps 1 --warmup 0 --span-days 0.05 --n-traj 9472 --no-cpu-baseline --no-strict --kernel transposed > gpurun_out/${T}_trace_bench.log 2>&1; echo "trace bench rc=$?"
python scripts/tx_trace.py gpurun_out/${T}_trace.bin 8 > gpurun_out/${T}_trace.txt 2>&1; head -5 gpurun_out/${T}_trace.txt
