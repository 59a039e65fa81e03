#!/bin/bash
# Round 2: the transposed kernel on 70x70 fields — parity at the C4 span with both kernels, its own tests, the C4 bench line with the oracle leg
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T=${TAG:-r02c4b}
timeout 900 python -m pytest -q -s -p no:cacheprovider -m gpu tests/test_gpu_baseline_spans.py tests/test_gpu_tx.py tests/test_gpu_frames_fields.py -k "c4 or 70 or cislunar" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^\[|passed|failed" gpurun_out/${T}_pytest.log | tail -8
timeout 400 python bench.py --workload c4 --no-strict --steps 2 --warmup 1 > gpurun_out/${T}_bench_c4.json 2> gpurun_out/${T}_bench_c4.err; echo "c4 rc=$?"
python - "$T" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/{sys.argv[1]}_bench_c4.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "parity")}, d["roofline"]["frac"], d["config"]["kernel"])
PY
