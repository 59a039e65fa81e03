#!/bin/bash
# Round 2: C4 (GRAIL 70x70 low lunar orbits) — lane-cooperative kernel at 32 lanes against the transposed kernel with 16 walker positions
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T=${TAG:-r02c4}
run() { tag=$1; shift; timeout 200 "$@" > gpurun_out/${T}_$tag.json 2> gpurun_out/${T}_$tag.err; echo "$tag rc=$?"; }
B="python bench.py --workload c4 --no-strict --no-cpu-baseline --steps 1 --warmup 1"
run coop $B --kernel coop
run tx $B --kernel transposed
run tx_2k $B --kernel transposed --n-traj 2000 --span-days 1
run coop_2k $B --kernel coop --n-traj 2000 --span-days 1
for f in coop tx coop_2k tx_2k; do python - "gpurun_out/${T}_$f.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], f"{d['value']:.4g} steps/s  frac {d['roofline']['frac']:.3f}  ms {d['ms_per_step']:.1f}", d['config'].get('kernel'), d['config'].get('ok_trajectories'))
except Exception as e:
    print(sys.argv[1], "failed:", e, open(sys.argv[1].replace('.json','.err')).read()[-300:])
PY
done
