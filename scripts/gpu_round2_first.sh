#!/bin/bash
# First GPU call of round 2: the experiments that decide where K2 goes next, in one gpurun call.
#   gpurun --timeout 600 -- 'bash scripts/gpu_round2_first.sh'
# Everything lands under gpurun_out/r02a_*.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-strict --steps 3 --warmup 3"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/smem_probe scripts/smem_probe.cu && timeout 120 /tmp/smem_probe > gpurun_out/r02a_smem_probe.txt 2>&1
timeout 120 $B > gpurun_out/r02a_c2_default.json 2> gpurun_out/r02a_c2_default.err
NYXB_COOP_SCHED=aligned timeout 120 $B > gpurun_out/r02a_c2_aligned.json 2> gpurun_out/r02a_c2_aligned.err
NYXB_COOP_T=2 timeout 120 $B > gpurun_out/r02a_c2_t2.json 2> gpurun_out/r02a_c2_t2.err
NYXB_COOP_T=2 NYXB_COOP_SCHED=aligned timeout 120 $B > gpurun_out/r02a_c2_t2_aligned.json 2> gpurun_out/r02a_c2_t2_aligned.err
timeout 120 $B --lanes 1 --n-traj 100000 > gpurun_out/r02a_c2_100k_k1.json 2> gpurun_out/r02a_c2_100k_k1.err
NYXB_K1_CONST=1 timeout 120 $B --lanes 1 --n-traj 100000 > gpurun_out/r02a_c2_100k_k1const.json 2> gpurun_out/r02a_c2_100k_k1const.err
NYXB_K1_CONST=1 timeout 120 $B --lanes 1 > gpurun_out/r02a_c2_k1const.json 2> gpurun_out/r02a_c2_k1const.err
NYXB_K1_CONST=1 timeout 100 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "thread or lanes1 or per_thread" > gpurun_out/r02a_pytest_k1const.log 2>&1; tail -2 gpurun_out/r02a_pytest_k1const.log
NYXB_COOP_SCHED=aligned timeout 120 $B --workload c4 --n-traj 2000 --span-days 1 > gpurun_out/r02a_c4_aligned.json 2> gpurun_out/r02a_c4_aligned.err
NYXB_COOP_SCHED=aligned timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -q -p no:cacheprovider > gpurun_out/r02a_pytest_aligned.log 2>&1; tail -3 gpurun_out/r02a_pytest_aligned.log
cat gpurun_out/r02a_smem_probe.txt
for f in default aligned t2 t2_aligned 100k_k1 100k_k1const k1const; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/r02a_c2_{sys.argv[1]}.json"))
    print(sys.argv[1], f"{d['value']:.4g} steps/s  frac {d['roofline']['frac']:.3f}")
except Exception as e:
    print(sys.argv[1], "failed:", e)
PY
done
