#!/bin/bash
# Round 2, GPU call S (validation of HEAD after the container was re-created): sets of 64 against sets of 32, the WHOLE GPU
# suite (BASELINE-span parity included), smoke(), the default bench line + the reference arm, the ncu launch list / full
# capture / full-span counters of the headline kernel, and one full capture each of the resampling, event-location and
# dispersion kernels.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T=${TAG:-r02s}
python -c "import nyx_b200.abi as a; a.load_library()" || { echo "libnyxb.so missing or stale"; exit 9; }
B="python bench.py --no-cpu-baseline --no-strict --steps 3 --warmup 3 --kernel transposed"
run() { tag=$1; shift; timeout 120 "$@" > gpurun_out/${T}_$tag.json 2> gpurun_out/${T}_$tag.err; echo "$tag rc=$?"; }
run s32_n10000 $B --tx-set 32
run s64_n10000 $B --tx-set 64
run s64_n9472 $B --tx-set 64 --n-traj 9472
run s32_n9472 $B --tx-set 32 --n-traj 9472
for f in s32_n10000 s64_n10000 s64_n9472 s32_n9472; do python - "$T" "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/{sys.argv[1]}_{sys.argv[2]}.json").read().strip().splitlines()[-1])
    print(sys.argv[2], f"{d['value']:.4g} steps/s  frac {d['roofline']['frac']:.3f}  ms {d['ms_per_step']:.1f} ok {d['config'].get('ok_trajectories')} parity {d.get('parity')}")
except Exception as e:
    print(sys.argv[2], "failed:", e)
PY
done
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -8 gpurun_out/${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -6 gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "bench ref rc=$?"
python - "$T" <<'PY'
import json, sys
T = sys.argv[1]
for f in ("bench_c2", "bench_ref"):
    try:
        d = json.loads(open(f"gpurun_out/{T}_{f}.json").read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "e2e", "roofline", "parity", "cpu_baseline", "gpu_launches", "clocks", "strict_bit_identical")})
    except Exception as e:
        print(f, "failed:", e)
PY
timeout 600 bash scripts/gpu_profile_tx.sh $T > gpurun_out/${T}_profile_tx.log 2>&1; echo "profile tx rc=$?"
timeout 300 bash scripts/gpu_profile_traj.sh ${T}_traj > gpurun_out/${T}_profile_traj.log 2>&1; echo "profile traj rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:nyxb_k_mvn -c 1 -o gpurun_out/${T}_mvn \
    python -m pytest tests/test_mvn.py -q -m gpu -p no:cacheprovider > gpurun_out/${T}_mvn_pytest.log 2>&1; echo "profile mvn rc=$?"
for r in tx traj_resample traj_locate mvn; do
  [ -f gpurun_out/${T}_$r.ncu-rep ] && timeout 120 python scripts/ncu_summary.py gpurun_out/${T}_$r.ncu-rep > gpurun_out/${T}_${r}_ncu_summary.txt 2>&1
done
ls -la gpurun_out/ | tail -40
