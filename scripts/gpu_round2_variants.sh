#!/bin/bash
# Compile-time variants of K2 measured on the GPU box (round 2): rebuilds ONLY nyxb_coop_g8.o with extra -D flags
# (csrc/Makefile: EXTRA), runs the default bench, and restores the default build at the end.
#   gpurun --timeout 900 -- 'bash scripts/gpu_round2_variants.sh'
# Register caps explored without a GPU (ptxas -v, T = 1): 96 registers / 444 B of spill stores at the default (128 threads, 5 CTAs
# per SM); 128 registers / 192 B at (64, 8) or (128, 4) — but 16 resident warps per SM hold only 2 368 of the 2 500 warps of the
# 10 000-trajectory ensemble, so that variant needs the 8 192-trajectory sub-case to be judged fairly (second bench line).
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-strict --steps 3 --warmup 3"
i=0
for extra in "" "-DCOOP_CTA=64 -DCOOP_MINB1=9" "-DCOOP_CTA=64 -DCOOP_MINB1=8" "-DCOOP_CTA=64 -DCOOP_MINB2=5" "-DCOOP_CTA=64 -DCOOP_MINB2=4"; do
  i=$((i + 1))
  touch nyx_b200/csrc/nyxb_coop_g8.cu
  make -C nyx_b200/csrc EXTRA="$extra" > gpurun_out/r02b_build_$i.log 2>&1 || { echo "build failed: $extra"; continue; }
  T=1; case "$extra" in *MINB2*) T=2;; esac
  NYXB_COOP_T=$T timeout 120 $B > gpurun_out/r02b_c2_$i.json 2> gpurun_out/r02b_c2_$i.err
  NYXB_COOP_T=$T timeout 120 $B --n-traj 8192 > gpurun_out/r02b_c2_8192_$i.json 2> gpurun_out/r02b_c2_8192_$i.err
  python - "$i" "$extra" "$T" <<'PY'
import json, sys
for tag in ("c2", "c2_8192"):
    try:
        d = json.load(open(f"gpurun_out/r02b_{tag}_{sys.argv[1]}.json"))
        print(f"[{sys.argv[2] or 'default'}] T={sys.argv[3]} {tag}: {d['value']:.4g} steps/s  frac {d['roofline']['frac']:.3f}")
    except Exception as e:
        print(f"[{sys.argv[2] or 'default'}] {tag} failed: {e}")
PY
done
touch nyx_b200/csrc/nyxb_coop_g8.cu
make -C nyx_b200/csrc > gpurun_out/r02b_build_restore.log 2>&1
