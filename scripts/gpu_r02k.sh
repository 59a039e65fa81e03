#!/bin/bash
# Round 2, GPU call K: transposed kernel with published powers, one column loop, even columns.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T=${TAG:-r02k}
python -c "import nyx_b200.abi as a; a.load_library()" || { echo "libnyxb.so missing or stale"; exit 9; }
timeout 150 python -m pytest tests/test_gpu_tx.py -x -q -p no:cacheprovider > gpurun_out/${T}_pytest_tx.log 2>&1; echo "pytest tx rc=$?"; tail -6 gpurun_out/${T}_pytest_tx.log
B="python bench.py --no-cpu-baseline --no-strict --steps 3 --warmup 3"
run() { tag=$1; shift; timeout 90 "$@" > gpurun_out/${T}_$tag.json 2> gpurun_out/${T}_$tag.err; echo "$tag rc=$?"; }
run tx_n10000 $B --kernel transposed
run tx_n9472 $B --kernel transposed --n-traj 9472
run tx_100k $B --kernel transposed --n-traj 100000 --steps 2 --warmup 1
run tx10_n10000 $B --kernel transposed --tx-positions 10
run tx10_100k $B --kernel transposed --n-traj 100000 --steps 2 --warmup 1 --tx-positions 10
for f in tx_n10000 tx_n9472 tx_100k tx10_n10000 tx10_100k; do python - "$T" "$f" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/{sys.argv[1]}_{sys.argv[2]}.json"))
    print(sys.argv[2], f"{d['value']:.4g} steps/s  frac {d['roofline']['frac']:.3f}  ms {d['ms_per_step']:.1f} ok {d['config']['ok_trajectories']}")
except Exception as e:
    print(sys.argv[2], "failed:", e)
PY
done
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 120 ncu --clock-control none -k regex:nyxb_k_tx -c 1 --metrics $M --csv --log-file gpurun_out/${T}_fullspan.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-strict --kernel transposed > gpurun_out/${T}_fullspan_bench.log 2>&1
grep -E "pipe_fp64|issue_active|lsu_wavefronts|time_duration|dram__bytes" gpurun_out/${T}_fullspan.csv | awk -F'","' '{print $(NF-2), $(NF)}'
timeout 100 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_case.py tx > gpurun_out/${T}_racecheck_tx.log 2>&1; echo "racecheck tx: $(grep -E 'RACECHECK SUMMARY|ERROR SUMMARY' gpurun_out/${T}_racecheck_tx.log | tail -1)"
if [ "${NCU_FULL:-1}" = 1 ]; then
timeout 200 ncu --set full --clock-control none --import-source on -k regex:nyxb_k_tx -c 1 -o gpurun_out/${T}_tx \
    python bench.py --steps 1 --warmup 0 --span-days 0.1 --n-traj 10000 --no-cpu-baseline --no-strict --kernel transposed > gpurun_out/${T}_tx_bench.log 2>&1
ls -la gpurun_out/${T}_tx.ncu-rep
fi
if [ "${TRACE:-0}" = 1 ]; then   # diagnostic timeline, built on the box (overwrites the box's copy of libnyxb.so: keep this step last)
touch nyx_b200/csrc/nyxb_tx.cu nyx_b200/csrc/nyxb_api.cu
timeout 400 make -C nyx_b200/csrc EXTRA=-DNYXB_TX_TRACE > gpurun_out/${T}_make.log 2>&1; echo "make rc=$?"
NYXB_TX_TRACE_FILE=gpurun_out/${T}_trace.bin timeout 120 python bench.py --steps 1 --warmup 0 --span-days 0.05 --n-traj 10000 --no-cpu-baseline --no-strict --kernel transposed > gpurun_out/${T}_trace_bench.log 2>&1; echo "trace bench rc=$?"
python scripts/tx_trace.py gpurun_out/${T}_trace.bin 8 | grep -E "walk|wait|busy|post|slack|between|context|helper warp|->" | head -44
fi
