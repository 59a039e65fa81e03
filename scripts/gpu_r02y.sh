#!/bin/bash
# Round 2, GPU call Y: small-shard lane sweep (strong scaling), context workloads C3 / C4 / C5 at HEAD, C2 at 100 000 trajectories.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T=${TAG:-r02y}
python -c "import nyx_b200.abi as a; a.load_library()" || { echo "libnyxb.so missing or stale"; exit 9; }
run() { tag=$1; shift; timeout 300 "$@" > gpurun_out/${T}_$tag.json 2> gpurun_out/${T}_$tag.err; echo "$tag rc=$?"; }
S2="python bench.py --no-cpu-baseline --no-strict --steps 3 --warmup 3"
for n in 1250 2500 5000; do
  for l in 8 16 32; do run shard_n${n}_l$l $S2 --n-traj $n --kernel coop --lanes $l; done
  run shard_n${n}_tx $S2 --n-traj $n --kernel transposed
done
run c2_100k $S2 --n-traj 100000 --steps 2 --warmup 1
run c3 python bench.py --workload c3 --no-strict
run c4 python bench.py --workload c4 --no-strict
run c5 python bench.py --workload c5 --no-strict
for f in gpurun_out/${T}_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], f"{d['value']:.4g} {d['unit']}  ms {d['ms_per_step']:.1f}  frac {d['roofline']['frac']:.3f}  parity {d.get('parity')}")
except Exception as e:
    print(sys.argv[1], "failed:", e)
PY
done
