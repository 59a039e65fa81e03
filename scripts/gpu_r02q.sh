#!/bin/bash
# quick: bench (2 sizes) + the transposed kernel's tests + the stretch timeline
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T=${TAG:-r02q2}
python -c "import nyx_b200.abi as a; a.load_library()" || { echo "libnyxb.so missing or stale"; exit 9; }
B="python bench.py --no-cpu-baseline --no-strict --steps 3 --warmup 3 --kernel transposed"
run() { tag=$1; shift; timeout 120 "$@" > gpurun_out/${T}_$tag.json 2> gpurun_out/${T}_$tag.err; echo "$tag rc=$?"; }
for n in 10000 9472; do run n$n $B --n-traj $n; done
for f in n10000 n9472; do python - "$T" "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/{sys.argv[1]}_{sys.argv[2]}.json").read().strip().splitlines()[-1])
    print(sys.argv[2], f"{d['value']:.4g} steps/s  frac {d['roofline']['frac']:.3f}  ms {d['ms_per_step']:.1f} ok {d['config'].get('ok_trajectories')}")
except Exception as e:
    print(sys.argv[2], "failed:", e)
PY
done
timeout 300 python -m pytest -q -p no:cacheprovider -m gpu tests/test_gpu_tx.py tests/test_gpu_parity.py -k "transposed or error_controls" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
touch nyx_b200/csrc/nyxb_tx.cu nyx_b200/csrc/nyxb_api.cu
timeout 400 make -C nyx_b200/csrc EXTRA=-DNYXB_TX_TRACE > gpurun_out/${T}_make.log 2>&1; echo "make rc=$?"
NYXB_TX_TRACE_FILE=gpurun_out/${T}_trace.bin timeout 120 python bench.py --steps 1 --warmup 0 --span-days 0.05 --n-traj 9472 --no-cpu-baseline --no-strict --kernel transposed > gpurun_out/${T}_trace_bench.log 2>&1; echo "trace bench rc=$?"
python scripts/tx_trace.py gpurun_out/${T}_trace.bin 8 > gpurun_out/${T}_trace.txt 2>&1; head -4 gpurun_out/${T}_trace.txt
