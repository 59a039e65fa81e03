#!/bin/bash
# Round 2 validation call at HEAD: the WHOLE GPU suite (BASELINE-span parity included), smoke(), the default bench line + the reference
# arm, ncu launch list / full capture / full-span counters of the headline kernel, racecheck of the transposed kernel.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T=${TAG:-r02f}
python -c "import nyx_b200.abi as a; a.load_library()" || { echo "libnyxb.so missing or stale"; exit 9; }
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -8 gpurun_out/${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -6 gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "bench ref rc=$?"
python - "$T" <<'PY'
import json, sys
T = sys.argv[1]
for f in ("bench_c2", "bench_ref"):
    try:
        d = json.loads(open(f"gpurun_out/{T}_{f}.json").read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "e2e", "roofline", "parity", "cpu_baseline", "gpu_launches", "clocks", "strict_bit_identical")})
    except Exception as e:
        print(f, "failed:", e)
PY
timeout 600 bash scripts/gpu_profile_tx.sh $T > gpurun_out/${T}_profile_tx.log 2>&1; echo "profile tx rc=$?"
[ -f gpurun_out/${T}_tx.ncu-rep ] && timeout 120 python scripts/ncu_summary.py gpurun_out/${T}_tx.ncu-rep > gpurun_out/${T}_tx_ncu_summary.txt 2>&1
timeout 150 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_case.py tx > gpurun_out/${T}_racecheck_tx.log 2>&1; grep -E "RACECHECK SUMMARY|steps" gpurun_out/${T}_racecheck_tx.log | tail -2
ls gpurun_out | grep ${T} | head -40
