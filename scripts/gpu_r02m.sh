#!/bin/bash
# Round 2, 2-GPU call: the multi-device C-ABI call on two real devices, weak and strong scaling of the default bench at N = 2.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T=${TAG:-r02m}
nvidia-smi -L
timeout 300 python -m pytest -q -p no:cacheprovider -m gpu tests/test_gpu_parity.py -k multi_device > gpurun_out/${T}_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/${T}_pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 400 $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-strict > gpurun_out/${T}_bench_2gpu_weak.json 2> gpurun_out/${T}_bench_2gpu_weak.err; echo "weak rc=$?"
timeout 400 $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-strict --no-cpu-baseline --scaling strong > gpurun_out/${T}_bench_2gpu_strong.json 2> gpurun_out/${T}_bench_2gpu_strong.err; echo "strong rc=$?"
for f in weak strong; do python - "gpurun_out/${T}_bench_2gpu_$f.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], f"{d['value']:.4g} {d['unit']} n_gpus {d['n_gpus']} ms {d['ms_per_step']:.1f} scaling {d['scaling']} e2e {d['e2e']['value']:.4g} config {d['config'].get('trajectories_total')}")
except Exception as e:
    print(sys.argv[1], "failed:", e)
PY
done
