#!/bin/bash
# Round 2 validation call: the whole GPU suite, smoke(), the default bench line + the reference arm, parity at the BASELINE spans.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T=${TAG:-r02v}
python -c "import nyx_b200.abi as a; a.load_library()" || { echo "libnyxb.so missing or stale"; exit 9; }
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --deselect tests/test_gpu_baseline_spans.py > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -5 gpurun_out/${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -6 gpurun_out/${T}_smoke.log
timeout 900 python -m pytest tests/test_gpu_baseline_spans.py -q -s -p no:cacheprovider > gpurun_out/${T}_pytest_spans.log 2>&1; echo "spans rc=$?"; grep -E "^\[|passed|failed|Error|assert" gpurun_out/${T}_pytest_spans.log | head -40
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
