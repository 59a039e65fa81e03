#!/bin/bash
# Round 2, GPU call A: operand-delivery probes + the K2 variants left unmeasured by round 1 + wave-quantisation sweep.
#   gpurun --timeout 900 -- 'bash scripts/gpu_r02a.sh'
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_gpu.txt 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/smem_probe scripts/smem_probe.cu && timeout 120 /tmp/smem_probe > gpurun_out/r02a_smem_probe.txt 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/opnd_probe scripts/opnd_probe.cu && timeout 200 /tmp/opnd_probe > gpurun_out/r02a_opnd_probe.txt 2>&1
B="python bench.py --no-cpu-baseline --no-strict --steps 3 --warmup 3"
run() { tag=$1; shift; timeout 150 "$@" > gpurun_out/r02a_$tag.json 2> gpurun_out/r02a_$tag.err; }
run c2_default $B
NYXB_COOP_SCHED=aligned run c2_aligned $B
NYXB_COOP_T=2 run c2_t2 $B
NYXB_COOP_T=2 NYXB_COOP_SCHED=aligned run c2_t2_aligned $B
run c2_n9472 $B --n-traj 9472
run c2_n11840 $B --n-traj 11840
run c2_n7104 $B --n-traj 7104
NYXB_COOP_T=2 run c2_t2_n9472 $B --n-traj 9472
NYXB_COOP_T=2 run c2_t2_n14208 $B --n-traj 14208
run c2_100k_k1 $B --lanes 1 --n-traj 100000
NYXB_K1_CONST=1 run c2_100k_k1const $B --lanes 1 --n-traj 100000
cat gpurun_out/r02a_smem_probe.txt gpurun_out/r02a_opnd_probe.txt
for f in c2_default c2_aligned c2_t2 c2_t2_aligned c2_n9472 c2_n11840 c2_n7104 c2_t2_n9472 c2_t2_n14208 c2_100k_k1 c2_100k_k1const; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/r02a_{sys.argv[1]}.json"))
    print(sys.argv[1], f"{d['value']:.4g} steps/s  frac {d['roofline']['frac']:.3f}  ms {d['ms_per_step']:.1f} max_dr {d.get('max_dr_km')}")
except Exception as e:
    print(sys.argv[1], "failed:", e)
PY
done
