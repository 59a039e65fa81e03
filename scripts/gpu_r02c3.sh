#!/bin/bash
# Round 2: C3 (BASELINE configs[2] on one GPU: 100 000 JWST-like trajectories, 30 days) at HEAD
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 500 python bench.py --workload c3 --n-traj 100000 --span-days 30 --no-strict > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; echo "c3 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_c3.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "roofline", "parity", "cpu_baseline")})
PY
