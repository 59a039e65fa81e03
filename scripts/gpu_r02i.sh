#!/bin/bash
# Round 2, GPU call I: where the transposed kernel's time goes after the walk-loop rewrite — slicing on/off at 10 000 / 100 000
# trajectories, and an ncu full capture (source counters) with every context occupied from start to end.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import nyx_b200.abi as a; a.load_library()" || { echo "libnyxb.so missing or stale"; exit 9; }
B="python bench.py --no-cpu-baseline --no-strict --steps 2 --warmup 1 --kernel transposed"
run() { tag=$1; shift; timeout 90 "$@" > gpurun_out/r02i_$tag.json 2> gpurun_out/r02i_$tag.err; echo "$tag rc=$?"; }
run n10000_noslice $B --tx-slice 100000000
run n10000_slice16 $B --tx-slice 16
run n10000_slice256 $B --tx-slice 256
run n100k_noslice $B --n-traj 100000 --tx-slice 100000000
run n100k_slice256 $B --n-traj 100000 --tx-slice 256
run n18944 $B --n-traj 18944
run n18944_noslice $B --n-traj 18944 --tx-slice 100000000
for f in n10000_noslice n10000_slice16 n10000_slice256 n100k_noslice n100k_slice256 n18944 n18944_noslice; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/r02i_{sys.argv[1]}.json"))
    print(sys.argv[1], f"{d['value']:.4g} steps/s  frac {d['roofline']['frac']:.3f}  ms {d['ms_per_step']:.1f} ok {d['config']['ok_trajectories']}")
except Exception as e:
    print(sys.argv[1], "failed:", e)
PY
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:nyxb_k_tx -c 1 -o gpurun_out/r02i_tx \
    python bench.py --steps 1 --warmup 0 --span-days 0.1 --n-traj 9472 --no-cpu-baseline --no-strict --kernel transposed > gpurun_out/r02i_tx_bench.log 2>&1
ls -la gpurun_out/r02i_tx.ncu-rep
