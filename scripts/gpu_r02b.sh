#!/bin/bash
# Round 2, GPU call B: first run of the transposed kernel (tests, bench at 9 472 / 10 000 trajectories), FAST-vs-oracle parity at
# the BASELINE spans, sanitizer pass.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tx.py -x -q -p no:cacheprovider > gpurun_out/r02b_pytest_tx.log 2>&1; tail -15 gpurun_out/r02b_pytest_tx.log
B="python bench.py --no-cpu-baseline --no-strict --steps 3 --warmup 3"
run() { tag=$1; shift; timeout 200 "$@" > gpurun_out/r02b_$tag.json 2> gpurun_out/r02b_$tag.err; }
run tx_n9472 $B --kernel transposed --n-traj 9472
run tx_n10000 $B --kernel transposed
run tx_n10000_s16 $B --kernel transposed --tx-slice 16
run tx_n10000_s256 $B --kernel transposed --tx-slice 256
run tx_n20000 $B --kernel transposed --n-traj 20000
run coop_n10000 $B --kernel coop
run tx_100k $B --kernel transposed --n-traj 100000
for f in tx_n9472 tx_n10000 tx_n10000_s16 tx_n10000_s256 tx_n20000 coop_n10000 tx_100k; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/r02b_{sys.argv[1]}.json"))
    print(sys.argv[1], f"{d['value']:.4g} steps/s  frac {d['roofline']['frac']:.3f}  ms {d['ms_per_step']:.1f} ok {d['config']['ok_trajectories']} kernel {d['config'].get('kernel')}")
except Exception as e:
    print(sys.argv[1], "failed:", e)
PY
done
timeout 900 python -m pytest tests/test_gpu_baseline_spans.py -q -s -p no:cacheprovider > gpurun_out/r02b_pytest_spans.log 2>&1; grep -E "^\[|passed|failed|Error|assert" gpurun_out/r02b_pytest_spans.log | head -40
for k in tx coop strict od; do
  timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_case.py $k > gpurun_out/r02b_racecheck_$k.log 2>&1; echo "racecheck $k: $(grep -E 'RACECHECK SUMMARY|ERROR SUMMARY' gpurun_out/r02b_racecheck_$k.log | tail -1)"
done
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_case.py all > gpurun_out/r02b_memcheck.log 2>&1; echo "memcheck: $(grep -E 'ERROR SUMMARY' gpurun_out/r02b_memcheck.log | tail -1)"
