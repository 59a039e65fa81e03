#!/bin/bash
# Round 2: C3 (BASELINE configs[2] on one GPU: 100 000 JWST-like trajectories, 30 days) at HEAD + the tests with third bodies
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T=${TAG:-r02}
timeout 500 python bench.py --workload c3 --n-traj 100000 --span-days 30 --no-strict > gpurun_out/${T}_bench_c3.json 2> gpurun_out/${T}_bench_c3.err; echo "c3 rc=$?"
python - "$T" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/{sys.argv[1]}_bench_c3.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "parity")})
PY
timeout 600 python -m pytest -q -p no:cacheprovider -m gpu tests/test_gpu_parity.py tests/test_gpu_baseline_spans.py tests/test_gpu_frames_fields.py tests/test_gpu_fuzz.py tests/test_events.py -k "not strict" > gpurun_out/${T}_pytest_bodies.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${T}_pytest_bodies.log
