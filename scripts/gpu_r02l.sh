#!/bin/bash
# Round 2, GPU call L: diagnostic timeline (clock64 stamps of CTA 0) of the transposed kernel, built on the box with -DNYXB_TX_TRACE.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
cp nyx_b200/csrc/libnyxb.so /tmp/libnyxb_product.so
touch nyx_b200/csrc/nyxb_tx.cu nyx_b200/csrc/nyxb_api.cu
timeout 400 make -C nyx_b200/csrc EXTRA=-DNYXB_TX_TRACE > gpurun_out/r02l_make.log 2>&1; echo "make rc=$?"
python -c "import nyx_b200.abi as a; a.load_library()" || { echo "libnyxb.so missing or stale"; exit 9; }
NYXB_TX_TRACE_FILE=gpurun_out/r02l_trace.bin timeout 120 python bench.py --steps 1 --warmup 0 --span-days 0.05 --n-traj 10000 --no-cpu-baseline --no-strict --kernel transposed > gpurun_out/r02l_bench.log 2>&1; echo "bench rc=$?"
ls -la gpurun_out/r02l_trace.bin
python scripts/tx_trace.py gpurun_out/r02l_trace.bin 8
