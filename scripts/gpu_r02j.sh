#!/bin/bash
# Round 2, GPU call J: polled READY mbarriers (walkers take whichever set has a stage ready), trig and candidate state off the serial path.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import nyx_b200.abi as a; a.load_library()" || { echo "libnyxb.so missing or stale"; exit 9; }
timeout 150 python -m pytest tests/test_gpu_tx.py -x -q -p no:cacheprovider > gpurun_out/r02j_pytest_tx.log 2>&1; echo "pytest tx rc=$?"; tail -6 gpurun_out/r02j_pytest_tx.log
B="python bench.py --no-cpu-baseline --no-strict --steps 3 --warmup 3"
run() { tag=$1; shift; timeout 90 "$@" > gpurun_out/r02j_$tag.json 2> gpurun_out/r02j_$tag.err; echo "$tag rc=$?"; }
run tx8_n10000 $B --kernel transposed
run tx12_n10000 $B --kernel transposed --tx-positions 12
run tx8_n9472 $B --kernel transposed --n-traj 9472
run tx12_n9472 $B --kernel transposed --n-traj 9472 --tx-positions 12
run tx8_100k $B --kernel transposed --n-traj 100000 --steps 2 --warmup 1
run tx12_100k $B --kernel transposed --n-traj 100000 --steps 2 --warmup 1 --tx-positions 12
for f in tx8_n10000 tx12_n10000 tx8_n9472 tx12_n9472 tx8_100k tx12_100k; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/r02j_{sys.argv[1]}.json"))
    print(sys.argv[1], f"{d['value']:.4g} steps/s  frac {d['roofline']['frac']:.3f}  ms {d['ms_per_step']:.1f} ok {d['config']['ok_trajectories']} kernel {d['config'].get('kernel')} parity {d.get('parity', {}).get('max_dr_km')}")
except Exception as e:
    print(sys.argv[1], "failed:", e)
PY
done
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active
for P in 8 12; do
timeout 120 ncu --clock-control none -k regex:nyxb_k_tx -c 1 --metrics $M --csv --log-file gpurun_out/r02j_fullspan_p$P.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-strict --kernel transposed --tx-positions $P > gpurun_out/r02j_fullspan_bench_p$P.log 2>&1
echo "P=$P"; grep -E "pipe_fp64|issue_active|lsu_wavefronts|time_duration|dram__bytes" gpurun_out/r02j_fullspan_p$P.csv | awk -F'","' '{print $(NF-2), $(NF)}'
done
timeout 100 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_case.py tx > gpurun_out/r02j_racecheck_tx.log 2>&1; echo "racecheck tx: $(grep -E 'RACECHECK SUMMARY|ERROR SUMMARY' gpurun_out/r02j_racecheck_tx.log | tail -1)"
