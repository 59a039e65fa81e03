#!/bin/bash
# Round 2, GPU call R: sets of 64 trajectories (two per walker lane) against sets of 32; the whole GPU suite.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T=${TAG:-r02r}
python -c "import nyx_b200.abi as a; a.load_library()" || { echo "libnyxb.so missing or stale"; exit 9; }
timeout 300 python -m pytest tests/test_gpu_tx.py -q -p no:cacheprovider > gpurun_out/${T}_pytest_tx.log 2>&1; echo "pytest tx rc=$?"; tail -8 gpurun_out/${T}_pytest_tx.log
B="python bench.py --no-cpu-baseline --no-strict --steps 3 --warmup 3 --kernel transposed"
run() { tag=$1; shift; timeout 90 "$@" > gpurun_out/${T}_$tag.json 2> gpurun_out/${T}_$tag.err; echo "$tag rc=$?"; }
run s64_n10000 $B --tx-set 64
run s64_n9472 $B --tx-set 64 --n-traj 9472
run s64_100k $B --tx-set 64 --n-traj 100000 --steps 2 --warmup 1
run s32_n10000 $B --tx-set 32
for f in s64_n10000 s64_n9472 s64_100k s32_n10000; do python - "$T" "$f" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/{sys.argv[1]}_{sys.argv[2]}.json"))
    print(sys.argv[2], f"{d['value']:.4g} steps/s  frac {d['roofline']['frac']:.3f}  ms {d['ms_per_step']:.1f} ok {d['config']['ok_trajectories']}")
except Exception as e:
    print(sys.argv[2], "failed:", e)
PY
done
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 120 ncu --clock-control none -k regex:nyxb_k_tx -c 1 --metrics $M --csv --log-file gpurun_out/${T}_fullspan_s64.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-strict --kernel transposed --tx-set 64 > gpurun_out/${T}_fullspan_bench.log 2>&1
grep -E "pipe_fp64|issue_active|lsu_wavefronts|time_duration|dram__bytes" gpurun_out/${T}_fullspan_s64.csv | awk -F'","' '{print $(NF-2), $(NF)}'
timeout 100 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_case.py tx64 > gpurun_out/${T}_racecheck_tx64.log 2>&1; echo "racecheck tx64: $(grep -E 'RACECHECK SUMMARY|ERROR SUMMARY' gpurun_out/${T}_racecheck_tx64.log | tail -1)"
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_baseline_spans.py --deselect tests/test_gpu_tx.py > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -6 gpurun_out/${T}_pytest_gpu.log
touch nyx_b200/csrc/nyxb_tx.cu nyx_b200/csrc/nyxb_api.cu
timeout 400 make -C nyx_b200/csrc EXTRA=-DNYXB_TX_TRACE > gpurun_out/${T}_make.log 2>&1; echo "make rc=$?"
NYXB_TX_TRACE_FILE=gpurun_out/${T}_trace.bin timeout 120 python bench.py --steps 1 --warmup 0 --span-days 0.05 --n-traj 9472 --no-cpu-baseline --no-strict --kernel transposed --tx-set 64 > gpurun_out/${T}_trace_bench.log 2>&1; echo "trace bench rc=$?"
python scripts/tx_trace.py gpurun_out/${T}_trace.bin 8 | grep -E "walk|wait|busy|post|slack|between|context|helper warp|->" | head -30
