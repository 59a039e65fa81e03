#!/bin/bash
# Round 2, GPU call Z: walker column loop unrolled by two + DCM split + set spreading: bench at several ensemble sizes, tests, timeline.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T=${TAG:-r02z}
python -c "import nyx_b200.abi as a; a.load_library()" || { echo "libnyxb.so missing or stale"; exit 9; }
B="python bench.py --no-cpu-baseline --no-strict --steps 3 --warmup 3 --kernel transposed"
run() { tag=$1; shift; timeout 120 "$@" > gpurun_out/${T}_$tag.json 2> gpurun_out/${T}_$tag.err; echo "$tag rc=$?"; }
for n in 10000 9472 5000 2500 1250; do run n$n $B --n-traj $n; done
for f in n10000 n9472 n5000 n2500 n1250; do python - "$T" "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/{sys.argv[1]}_{sys.argv[2]}.json").read().strip().splitlines()[-1])
    print(sys.argv[2], f"{d['value']:.4g} steps/s  frac {d['roofline']['frac']:.3f}  ms {d['ms_per_step']:.1f} ok {d['config'].get('ok_trajectories')}")
except Exception as e:
    print(sys.argv[2], "failed:", e)
PY
done
timeout 600 python -m pytest -q -p no:cacheprovider -m gpu tests/test_gpu_tx.py tests/test_gpu_baseline_spans.py tests/test_gpu_frames_fields.py -k "tx or transposed or c2 or frame or cislunar" -s > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^\[|passed|failed" gpurun_out/${T}_pytest.log | tail -12
touch nyx_b200/csrc/nyxb_tx.cu nyx_b200/csrc/nyxb_api.cu
timeout 400 make -C nyx_b200/csrc EXTRA=-DNYXB_TX_TRACE > gpurun_out/${T}_make.log 2>&1; echo "make rc=$?"
NYXB_TX_TRACE_FILE=gpurun_out/${T}_trace.bin timeout 120 python bench.py --steps 1 --warmup 0 --span-days 0.05 --n-traj 9472 --no-cpu-baseline --no-strict --kernel transposed > gpurun_out/${T}_trace_bench.log 2>&1; echo "trace bench rc=$?"
python scripts/tx_trace.py gpurun_out/${T}_trace.bin 8 > gpurun_out/${T}_trace.txt 2>&1; head -4 gpurun_out/${T}_trace.txt
